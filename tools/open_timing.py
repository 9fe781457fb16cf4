"""open() on a resident 2^k polynomial for several side-lane thresholds (PLONKISH_CUDA_MANY_SMALL_LOG2)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import plonkish_b200 as pk
from plonkish_b200 import kzg
from bench import g1_generator

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
pp = kzg.setup(g1_generator(np), pk.random_scalars(k, seed=77))
poly = pk.ResidentScalars(pk.random_scalars(1 << k, seed=1))
point = pk.random_scalars(k, seed=79)
ref = None
for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("default", "17", "19", "21")):
    if v == "default":
        os.environ.pop("PLONKISH_CUDA_MANY_SMALL_LOG2", None)  # library rule: a quarter of the largest MSM
    else:
        os.environ["PLONKISH_CUDA_MANY_SMALL_LOG2"] = v
    q, val = kzg.open_resident(pp, poly, point)
    got = np.stack(q).tobytes()
    ref = ref or got
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); kzg.open_resident(pp, poly, point); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"k={k} small<=2^{v}: open_resident {min(ts):.2f} ms (median {sorted(ts)[len(ts)//2]:.2f}) same={got == ref}", flush=True)
