python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/prove_bench.py 24 3 0 > gpurun_out/s4_prove24.json 2> gpurun_out/s4_prove24.err; tail -c 300 gpurun_out/s4_prove24.err
python tools/prove_bench.py 20 3 1 > gpurun_out/s4_prove20.json 2> gpurun_out/s4_prove20.err; tail -c 300 gpurun_out/s4_prove20.err
python - <<'PY'
import json
for k in (24,20):
    d=json.load(open(f'gpurun_out/s4_prove{k}.json'))
    print(k, d['gpu_ms'], d['gpu_ms_all'], d['phases_ms'][0], d.get('verifier_accepts'), d.get('proof_bytes_identical_to_cpu'), d.get('cpu_ms'))
PY
