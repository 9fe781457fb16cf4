"""Run length of a k_accumulate thread (PLONKISH_CUDA_ACC_WAVES / _ACC_L) against the MSM time, device-resident table-mode
MSMs at 2^18 .. 2^24 -> gpurun_out/acc_run_length.json.  Every result is compared with the default plan's point."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk

sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [18, 19, 20, 21, 22, 23, 24]
waves = sys.argv[2].split(",") if len(sys.argv) > 2 else ["default", "1", "1.5", "2", "2.5", "3", "4", "5", "6", "8"]  # "L<v>": equal runs of v entries, "T<j>": j tiers
res = {}
for lg in sizes:
    n = 1 << lg
    sc = pk.random_scalars(n, 1)
    d = torch.from_numpy(sc.view(np.int64)).cuda()
    b = pk.synth_bases_device(n, 3, 5)
    torch.cuda.synchronize()
    reg = pk.G1Bases(b, mode=pk.G1Bases.TABLE)
    ref = None
    row = {}
    for w in waves:
        os.environ.pop("PLONKISH_CUDA_ACC_WAVES", None)
        os.environ.pop("PLONKISH_CUDA_ACC_L", None)
        os.environ.pop("PLONKISH_CUDA_ACC_TIERS", None)
        if w.startswith("T"):
            os.environ["PLONKISH_CUDA_ACC_TIERS"] = w[1:]
        elif w.startswith("L"):
            os.environ["PLONKISH_CUDA_ACC_L"] = w[1:]
        elif w != "default":
            os.environ["PLONKISH_CUDA_ACC_WAVES"] = w
        plan = pk.msm_plan(n, 0, 0, bases=reg)
        got = pk.variable_base_msm_device(d, reg).cpu().numpy().tobytes()
        ref = ref or got
        assert got == ref, (lg, w)
        st = pk.profile_stages_device(d, reg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            e0.record(); pk.variable_base_msm_device(d, reg); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row[w] = {"L": plan.get("run_length"), "ms": round(min(ts), 4), "accumulate": round(st["accumulate"], 4), "item_levels": round(st["item_levels"], 4)}
        print(lg, w, row[w], flush=True)
    os.environ.pop("PLONKISH_CUDA_ACC_WAVES", None)
    os.environ.pop("PLONKISH_CUDA_ACC_L", None)
    os.environ.pop("PLONKISH_CUDA_ACC_TIERS", None)
    res[lg] = row
    reg.release()
    del d, b
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("ACC_OUT", "acc_run_length.json")), "w"), indent=1)
