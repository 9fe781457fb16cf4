"""Where the time of a device-built SRS goes: table registration of one slice vs the whole setup."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import plonkish_b200 as pk
from plonkish_b200 import kzg
from bench import g1_generator

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
d_bases = pk.synth_bases_device(n, 7, 3)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); reg = pk.G1Bases(d_bases, mode=pk.G1Bases.TABLE); torch.cuda.synchronize()
    print(f"register 2^{k} table: {time.perf_counter() - t0:.3f} s", flush=True)
    reg.release()
del d_bases
torch.cuda.empty_cache()
ss = pk.random_scalars(k, seed=77)
for rep in range(2):
    t0 = time.perf_counter(); pp = kzg.setup(g1_generator(np), ss); torch.cuda.synchronize()
    print(f"kzg.setup k={k}: {time.perf_counter() - t0:.3f} s", flush=True)
    pp.release()
sc = pk.random_scalars(1 << 22, seed=5)
for rep in range(2):
    t0 = time.perf_counter(); pk.fixed_base_msm(g1_generator(np), sc)
    print(f"fixed_base_msm 2^22 host->host: {time.perf_counter() - t0:.3f} s", flush=True)
