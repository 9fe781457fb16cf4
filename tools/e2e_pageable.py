"""Host-call MSM (C ABI, pageable numpy scalars, registered bases) at 2^k points: min / median ms over 8 calls.
Run once per setting of PLONKISH_CUDA_STAGE_NT / PLONKISH_CUDA_COPY_THREADS (read at library start)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonkish_b200 as pk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
sc = pk.random_scalars(n, 1)
b = pk.synth_bases_device(n, 3, 5)
reg = pk.G1Bases(b, mode=pk.G1Bases.TABLE)
ref = pk.variable_base_msm(sc, reg)
ts = []
for _ in range(8):
    t0 = time.perf_counter(); got = pk.variable_base_msm(sc, reg); ts.append((time.perf_counter() - t0) * 1e3)
    assert (got == ref).all()
print(f"k={k} STAGE_NT={os.environ.get('PLONKISH_CUDA_STAGE_NT', '0')} COPY_THREADS={os.environ.get('PLONKISH_CUDA_COPY_THREADS', 'auto')}: "
      f"e2e pageable {min(ts):.2f} ms (median {sorted(ts)[len(ts) // 2]:.2f})", flush=True)
