"""MSM sweep 2^16..2^26 on one GPU (BASELINE.json config 5, single-GPU column): device-resident and
host-scalar (e2e) Mpoints/s for resident bases (table of window multiples) and plain bases, each
result checked against the known-discrete-log answer.  Writes gpurun_out/sweep.json."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk
from oracle import pyoracle as po

lo = int(sys.argv[1]) if len(sys.argv) > 1 else 16
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 26
rows = []
for lg in range(lo, hi + 1):
    n = 1 << lg
    host_t = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    sc = host_t.numpy().view(np.uint64)
    sc[:] = pk.random_scalars(n, seed=lg)
    d_sc = host_t.cuda()
    d_bs = pk.synth_bases_device(n, 3, 5)
    torch.cuda.synchronize()
    want = po.known_dlog_answer(3, 5, sc)
    row = {"log_n": lg}
    for mode, name in ((0, "table"), (pk.G1Bases.PLAIN, "plain")):
        t0 = time.perf_counter()
        reg = pk.G1Bases(d_bs, mode=mode)
        torch.cuda.synchronize()
        reg_s = time.perf_counter() - t0
        ok = pk.variable_base_msm_device(d_sc, reg).cpu().numpy().view(np.uint64).tobytes() == want.tobytes()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if lg <= 22 else 3
        ts = []
        for _ in range(reps):
            e0.record(); pk.variable_base_msm_device(d_sc, reg); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ok = ok and pk.variable_base_msm(sc, reg).tobytes() == want.tobytes()
        t0 = time.perf_counter()
        for _ in range(reps):
            pk.variable_base_msm(sc, reg)
        e2e_ms = (time.perf_counter() - t0) / reps * 1e3
        page = np.array(sc)  # pageable copy (what a Rust Vec<Fr> is): goes through the staging ring
        ok = ok and pk.variable_base_msm(page, reg).tobytes() == want.tobytes()
        t0 = time.perf_counter()
        for _ in range(reps):
            pk.variable_base_msm(page, reg)
        e2e_page_ms = (time.perf_counter() - t0) / reps * 1e3
        del page
        plan = pk.msm_plan(n, 0, 0, bases=reg)
        row[name] = {"parity": bool(ok), "device_ms": round(min(ts), 3), "device_mpts": round(n / min(ts) / 1e3, 1),
                     "e2e_ms": round(e2e_ms, 3), "e2e_mpts": round(n / e2e_ms / 1e3, 1),
                     "e2e_pageable_ms": round(e2e_page_ms, 3), "e2e_pageable_mpts": round(n / e2e_page_ms / 1e3, 1), "c": plan["window_bits"], "windows": plan["windows"],
                     "register_s": round(reg_s, 3)}
        reg.release()
    print(json.dumps(row), flush=True)
    rows.append(row)
    del d_sc, d_bs
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
