"""Wall-clock split of the resident prove sequence (bench.py prove_msm_sequence, gpu_resident_ms) at 2^k rows."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import plonkish_b200 as pk
from plonkish_b200 import kzg
from bench import g1_generator

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
pp = kzg.setup(g1_generator(np), pk.random_scalars(k, seed=77))
polys = []
for j in range(4):
    t = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    h = t.numpy().view(np.uint64)
    h[:] = pk.random_scalars(n, seed=4242 + j)
    polys.append(h)
coeffs = pk.random_scalars(4, seed=78)
point = pk.random_scalars(k, seed=79)
for rep in range(3):
    marks = [("start", time.perf_counter())]
    c3, r3 = kzg.batch_commit(pp, polys[:3], keep=True); marks.append(("batch_keep x3", time.perf_counter()))
    c1, r1 = kzg.batch_commit(pp, polys[3:], keep=True); marks.append(("batch_keep x1", time.perf_counter()))
    g = kzg.linear_combination(r3 + r1, coeffs); marks.append(("g_prime merge", time.perf_counter()))
    q, v = kzg.open_resident(pp, g, point); marks.append(("open_resident", time.perf_counter()))
    for r in r3 + r1 + [g]:
        r.release()
    marks.append(("release", time.perf_counter()))
    print("rep", rep, " ".join(f"{name}={1e3 * (t - marks[i][1]):.2f}ms" for i, (name, t) in enumerate(marks[1:])),
          f"total={1e3 * (marks[-1][1] - marks[0][1]):.2f}ms", flush=True)
    marks = [("start", time.perf_counter())]
    pk.variable_base_msm_batch(polys[:3], pp.eq(k)); marks.append(("batch x3", time.perf_counter()))
    pk.variable_base_msm(polys[3], pp.eq(k)); marks.append(("single", time.perf_counter()))
    pk.variable_base_msm_many([polys[0][: 1 << i] for i in reversed(range(k))], [pp.eq(i) for i in reversed(range(k))]); marks.append(("many (host open)", time.perf_counter()))
    print("host", rep, " ".join(f"{name}={1e3 * (t - marks[i][1]):.2f}ms" for i, (name, t) in enumerate(marks[1:])),
          f"total={1e3 * (marks[-1][1] - marks[0][1]):.2f}ms", flush=True)
