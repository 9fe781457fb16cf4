"""Device-resident MSM time for the skewed scalar distributions of SURVEY.md §8(d) (table layout)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk
from plonkish_b200 import kzg
from oracle import pyoracle as po

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << log_n
R = kzg.FR_MODULUS
rng = np.random.default_rng(1)


def mont_const(v):
    return np.frombuffer((v % R * (1 << 256) % R).to_bytes(32, "little"), dtype=np.uint64)


def build(kind):
    if kind == "uniform":
        return pk.random_scalars(n, 5)
    out = np.zeros((n, 4), dtype=np.uint64)
    if kind == "selector":  # 0 / 1 / -1, half zeros (backend/hyperplonk/util.rs:133-152)
        pick = rng.integers(0, 4, n)
        out[pick == 2] = mont_const(1)
        out[pick == 3] = mont_const(R - 1)
    elif kind == "all-ones":
        out[:] = mont_const(1)
    elif kind == "same-wide":
        out[:] = mont_const(0x2AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA)
    elif kind == "small-ints":  # permutation polys: values < 3 * 2^k (preprocessor.rs:184-190)
        vals = rng.integers(0, 3 * n, n)
        table = {}
        # small values: Montgomery form via vectorised big-int is slow; use the product's own conversion on a sample
        out = kzg.fr_to_montgomery([int(v) for v in vals[: min(n, 1 << 16)]])
        out = np.tile(out, (n // out.shape[0] + 1, 1))[:n].copy()
    return out


d_bs = pk.synth_bases_device(n, 3, 5)
torch.cuda.synchronize()
reg = pk.G1Bases(d_bs)
res = {}
for kind in ("uniform", "selector", "all-ones", "same-wide", "small-ints"):
    sc = build(kind)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    got = pk.variable_base_msm_device(d_sc, reg).cpu().numpy().view(np.uint64)
    ok = bool((got == po.known_dlog_answer(3, 5, sc)).all())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record(); pk.variable_base_msm_device(d_sc, reg); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    st = pk.profile_stages_device(d_sc, reg)
    res[kind] = {"parity": ok, "ms": round(min(ts), 3), "stages_ms": {k: round(v, 3) for k, v in st.items()}}
    print(kind, res[kind], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"skew_2p{log_n}.json"), "w"), indent=1)
