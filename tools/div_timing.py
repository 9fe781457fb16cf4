"""Division by (X - z) on resident coefficients (plonkish_cuda_fr_div_linear) at 2^10 .. 2^24 coefficients, wall time per
call (the call synchronises); run once per PLONKISH_CUDA_HORNER_LOG_CHUNK value (read once per process):
PLONKISH_CUDA_HORNER_LOG_CHUNK=5 python tools/div_timing.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import plonkish_b200 as pk  # noqa: E402

if __name__ == "__main__":
    torch.cuda.init()
    z = pk.random_scalars(1, seed=5)[0]
    out = {"log_chunk": os.environ.get("PLONKISH_CUDA_HORNER_LOG_CHUNK", "default")}
    for log_n in (1, 10, 14, 16, 18, 20, 22, 24):
        poly = pk.ResidentScalars(pk.random_scalars(1 << log_n, seed=log_n))
        ms = []
        for _ in range(8):
            t0 = time.perf_counter()
            q, _ = pk.fr_div_linear(poly, z)
            ms.append((time.perf_counter() - t0) * 1e3)
            q.release()
        out[str(log_n)] = round(min(ms[2:]), 3)
        poly.release()
    print(json.dumps(out))
