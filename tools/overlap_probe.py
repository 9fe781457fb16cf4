"""Do two large MSMs on separate streams (own scratch) finish sooner than back to back?  Probe for overlapping
the HBM-bound sort of one with the IMAD-bound accumulate of the other."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import plonkish_b200 as pk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
d_bases = pk.synth_bases_device(n, 7, 3)
torch.cuda.synchronize()
reg = pk.G1Bases(d_bases)
hosts = []
for j in range(4):
    t = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    h = t.numpy().view(np.uint64); h[:] = pk.random_scalars(n, seed=j); hosts.append(h)
def timeit(f, reps=3):
    f(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)
seq = timeit(lambda: [pk.variable_base_msm(h, reg) for h in hosts])
batch = timeit(lambda: pk.variable_base_msm_batch(hosts, reg))
os.environ["PLONKISH_CUDA_MANY_SMALL_LOG2"] = "24"
lanes = timeit(lambda: pk.variable_base_msm_many(hosts, [reg] * 4))
os.environ["PLONKISH_CUDA_MANY_SMALL_LOG2"] = "0"
main = timeit(lambda: pk.variable_base_msm_many(hosts, [reg] * 4))
print(f"k={k}: 4 MSMs back to back {seq:.1f} ms, batch {batch:.1f} ms, many on side lanes (concurrent streams) {lanes:.1f} ms, many on the main stream {main:.1f} ms")
