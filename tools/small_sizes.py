"""Stage times (second call) for small MSMs in both bucket layouts."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk
import time
for lg in (4, 8, 10, 12, 14, 16, 18):
    n = 1 << lg
    sc = pk.random_scalars(n, 1)
    d = torch.from_numpy(sc.view(np.int64)).cuda()
    b = pk.synth_bases_device(n, 3, 5)
    torch.cuda.synchronize()
    for mode, name in ((pk.G1Bases.TABLE, "table"), (pk.G1Bases.PLAIN, "plain")):
        reg = pk.G1Bases(b, mode=mode)
        pk.profile_stages_device(d, reg)
        st = pk.profile_stages_device(d, reg)
        host = sc
        pk.variable_base_msm(host, reg)
        t = time.perf_counter()
        for _ in range(10):
            pk.variable_base_msm(host, reg)
        ms = (time.perf_counter() - t) * 100
        plan = pk.msm_plan(n, 0, 0, bases=reg)
        print(lg, name, "c=%d W=%d L=%d thr=%d" % (plan["window_bits"], plan["windows"], plan["run_length"], plan["accumulate_threads"]),
              "host-call %.3f ms" % ms, {k: round(v, 3) for k, v in st.items()}, flush=True)
        reg.release()
