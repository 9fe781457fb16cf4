"""Stage times of device-resident table-mode MSMs at 2^12 .. 2^24 (second call; mean of 3) -> gpurun_out/stage_profile.json."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk

res = {}
for lg in (12, 14, 16, 18, 19, 20, 21, 22, 23, 24):
    n = 1 << lg
    sc = pk.random_scalars(n, 1)
    d = torch.from_numpy(sc.view(np.int64)).cuda()
    b = pk.synth_bases_device(n, 3, 5)
    torch.cuda.synchronize()
    reg = pk.G1Bases(b, mode=pk.G1Bases.TABLE)
    pk.profile_stages_device(d, reg)
    runs = [pk.profile_stages_device(d, reg) for _ in range(3)]
    st = {k: round(sum(r[k] for r in runs) / 3, 4) for k in runs[0]}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record(); pk.variable_base_msm_device(d, reg); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    plan = pk.msm_plan(n, 0, 0, bases=reg)
    res[lg] = {"c": plan["window_bits"], "W": plan["windows"], "ms": round(min(ts), 4), "stage_sum": round(sum(st.values()), 4), "stages": st}
    print(lg, res[lg], flush=True)
    reg.release()
    del d, b
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "stage_profile.json"), "w"), indent=1)
