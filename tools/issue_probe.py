"""Is the accumulate loop bound by the multiplier pipe or by instruction issue?"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk
r = pk.bench_issue_mix()
print(json.dumps(r, indent=1))
print(json.dumps(pk.bench_integer_pipe(), indent=1))
print(json.dumps(pk.bench_fp64_pipe(), indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(r, open(os.path.join(ROOT, "gpurun_out", "issue_probe.json"), "w"), indent=1)
