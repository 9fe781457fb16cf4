"""Cycles per phase of the staged partition kernels (block 7's view) from a -DPK_STAGE_PROF build:
  nvcc ... -DPK_STAGE_PROF -o plonkish_b200/libplonkish_cuda_prof.so plonkish_b200/csrc/api.cu plonkish_b200/csrc/host_copy.cpp
  PLONKISH_CUDA_LIB=plonkish_b200/libplonkish_cuda_prof.so python tools/stage_phase_probe.py [log_n]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonkish_b200 as pk
from plonkish_b200 import _lib

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
d = torch.from_numpy(pk.random_scalars(n, 1).view(np.int64)).cuda()
reg = pk.G1Bases(pk.synth_bases_device(n, 3, 5), mode=pk.G1Bases.TABLE)
lib = _lib.load()
out = (ctypes.c_ulonglong * 16)()
pk.variable_base_msm_device(d, reg)
lib.plonkish_cuda_debug_stage_prof(0, out, 1)
reps = 3
for _ in range(reps):
    pk.variable_base_msm_device(d, reg)
lib.plonkish_cuda_debug_stage_prof(0, out, 1)
names = ["load + digit split", "rank (shared atomics) + barrier", "scan of the bin counts", "global cursors + stage to shared + barrier", "write-out + barrier", "reset + barrier"]
for base, title in ((0, "level 1 (k_scatter_staged_b)"), (8, "level 2 (k_bucket_scatter_staged_b)")):
    v = [out[base + i] / reps for i in range(6)]
    tot = sum(v)
    print(title, f"block 7: {tot / 1e3:.1f} kcycles per MSM")
    for nm, x in zip(names, v):
        print(f"   {nm:45s} {x / 1e3:9.1f} kcycles  {100 * x / max(tot, 1):5.1f} %")
