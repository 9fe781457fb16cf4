"""FP64-pipe arithmetic on the GPU: bit-exactness against the integer pipe, then stream throughputs alone and together."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk
from plonkish_b200 import _lib
from oracle import pyoracle as po
from oracle import bigint_ref as br

lib = _lib.lib()
P = br.P
rng = np.random.default_rng(7)
n = 4096
vals = [0, 1, P - 1, P - 2] + [int.from_bytes(rng.bytes(32), "little") % P for _ in range(n - 4)]
a = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vals), dtype=np.uint64).reshape(-1, 4).copy()
b = np.roll(a, 3, axis=0).copy()
def fop(op, x, y=None):
    out = np.zeros_like(x)
    _lib.check(lib.plonkish_cuda_debug_field_op(0, op, x.ctypes.data, None if y is None else y.ctypes.data, out.ctypes.data, x.shape[0]), "debug_field_op")
    return out
print("dp mul == fq mul:", bool((fop(12, a, b) == fop(0, a, b)).all()))
print("dp sqr == fq sqr:", bool((fop(13, a) == fop(10, a)).all()))
print("word round trip :", bool((fop(14, a) == a).all()))
# points
m = 2048
pts = po.known_dlog_bases(5, 7, m)
acc = np.zeros((m, 16), dtype=np.uint64)
def pop(op, x, y):
    out = np.zeros_like(x)
    _lib.check(lib.plonkish_cuda_debug_point_op(0, op, x.ctypes.data, y.ctypes.data, out.ctypes.data, x.shape[0]), "debug_point_op")
    return out
bb = np.zeros((m, 16), dtype=np.uint64); bb[:, :8] = pts
a1 = pop(0, acc, bb); a2 = pop(4, acc, bb)
ok = bool((a1 == a2).all())
bb2 = np.zeros((m, 16), dtype=np.uint64); bb2[:, :8] = np.roll(pts, 1, axis=0); bb2[::7, :8] = pts[::7]  # every 7th: P + P
bb2[3::11, :8] = 0                                                                                            # identity bases
neg = pts[5::13].copy()
bb2[5::13, :8] = neg; 
a1b = pop(0, a1, bb2); a2b = pop(4, a2, bb2)
ok = ok and bool((a1b == a2b).all())
print("dp madd == int madd (incl. P+P, identity):", ok)
res = {}
for dpb, ib in ((1, 2), (1, 3), (2, 2), (1, 4), (2, 4)):
    r = pk.bench_dp_madd(dpb, ib)
    res[f"dp{dpb}_int{ib}"] = r
    tot_alone = r["int_madd_per_s_alone"]
    print(f"dp blocks/SM {dpb}, int blocks/SM {ib}: dp alone {r['dp_madd_per_s_alone']/1e9:.3f} G/s, int alone {r['int_madd_per_s_alone']/1e9:.3f} G/s, "
          f"together dp {r['dp_madd_per_s_together']/1e9:.3f} + int {r['int_madd_per_s_together']/1e9:.3f} G/s, wall {r['together_ms']:.2f} ms", flush=True)
print("madd streams:", pk.bench_madd())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dp_probe.json"), "w"), indent=1)
