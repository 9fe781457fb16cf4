"""Round-1 bring-up on a B200: stage-by-stage parity diagnostics + first timings.
Writes gpurun_out/first_light.json.  Usage: python tools/gpu_first_light.py [max_log_n]"""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
report = {"device": torch.cuda.get_device_name(0), "steps": []}


def step(name, fn):
    t = time.time()
    try:
        res = fn()
        report["steps"].append({"name": name, "ok": True, "s": round(time.time() - t, 3), "result": res})
        print(f"[ok]   {name}: {res}", flush=True)
    except Exception as e:  # noqa: BLE001
        report["steps"].append({"name": name, "ok": False, "error": repr(e), "trace": traceback.format_exc()[-1500:]})
        print(f"[FAIL] {name}: {e!r}", flush=True)


def small_parity():
    bad = []
    for n in (1, 2, 33, 1000, 4097, 1 << 14):
        sc = po.random_scalars(n, n)
        bs = po.known_dlog_bases(3, 5, n)
        want = po.known_dlog_answer(3, 5, sc)
        got = pk.variable_base_msm(sc, bs)
        if got.tobytes() != want.tobytes():
            bad.append(n)
    assert not bad, f"mismatch at n={bad}"
    return "host path parity ok"


def synth_parity():
    n = 1 << 12
    d = pk.synth_bases_device(n, 3, 5).cpu().numpy().view(np.uint64)
    assert (d == po.known_dlog_bases(3, 5, n)).all()
    return "synthetic bases match the oracle"


def window_sweep():
    n = 20000
    sc = po.random_scalars(n, 5)
    bs = po.known_dlog_bases(3, 5, n)
    want = po.known_dlog_answer(3, 5, sc)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_bs = torch.from_numpy(bs.view(np.int64)).cuda()
    bad = [c for c in range(8, 17) if pk.variable_base_msm_device(d_sc, d_bs, window_bits=c).cpu().numpy().view(np.uint64).tobytes() != want.tobytes()]
    assert not bad, f"mismatch at c={bad}"
    return "c=8..16 ok"


def timing(log_n, cs=(0,)):
    n = 1 << log_n
    sc = pk.random_scalars(n, seed=log_n)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_bs = pk.synth_bases_device(n, 3, 5)
    torch.cuda.synchronize()
    want = po.known_dlog_answer(3, 5, sc)
    out = {}
    for c in cs:
        got = pk.variable_base_msm_device(d_sc, d_bs, window_bits=c).cpu().numpy().view(np.uint64)
        ok = got.tobytes() == want.tobytes()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for _ in range(3):
            e0.record()
            pk.variable_base_msm_device(d_sc, d_bs, window_bits=c)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        stages = pk.profile_stages_device(d_sc, d_bs, window_bits=c)
        plan = pk.msm_plan(n, c)
        out[f"c={plan['window_bits']}"] = {
            "parity": ok, "ms": round(min(times), 3), "mpts_per_s": round(n / min(times) / 1e3, 1),
            "stages_ms": {k: round(v, 3) for k, v in stages.items()}, "plan": plan,
        }
    # resident bases with the table of window multiples
    t0 = time.time()
    reg = pk.G1Bases(d_bs)
    torch.cuda.synchronize()
    t_reg = time.time() - t0
    got = pk.variable_base_msm_device(d_sc, reg).cpu().numpy().view(np.uint64)
    ok = got.tobytes() == want.tobytes()
    times = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        e0.record()
        pk.variable_base_msm_device(d_sc, reg)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    stages = pk.profile_stages_device(d_sc, reg)
    plan = pk.msm_plan(n, 0, 0, bases=reg)
    out[f"table c={plan['window_bits']}"] = {
        "parity": ok, "ms": round(min(times), 3), "mpts_per_s": round(n / min(times) / 1e3, 1), "register_s": round(t_reg, 3),
        "stages_ms": {k: round(v, 3) for k, v in stages.items()}, "plan": plan,
    }
    reg.release()
    return out


step("integer_pipe", lambda: pk.bench_integer_pipe(0))
step("synth_parity", synth_parity)
step("small_parity", small_parity)
step("window_sweep", window_sweep)
max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for lg in (12, 16, 18, 20, 22, 24, 26):
    if lg > max_log:
        break
    cs = (0,) if lg < 20 else (0, 14, 16) if lg < 24 else (0,)
    step(f"timing_2^{lg}", lambda lg=lg, cs=cs: timing(lg, cs))
report["launches"] = pk.launch_count()
with open(os.path.join(ROOT, "gpurun_out", "first_light.json"), "w") as f:
    json.dump(report, f, indent=1)
print("wrote gpurun_out/first_light.json")
