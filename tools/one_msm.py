"""Two device-resident table-mode MSMs of 2^k points (the second is the one to capture under ncu:
   ncu --set full --clock-control none --import-source on -k regex:'k_(decompose_b|scatter_staged_b|bucket_hist_b|bucket_scatter_staged_b|accumulate|reduce_items|bucket_reduce|combine_single)' --launch-skip <launches of the first> -c <same> ...)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonkish_b200 as pk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
sc = pk.random_scalars(n, 1)
d = torch.from_numpy(sc.view(np.int64)).cuda()
b = pk.synth_bases_device(n, 3, 5)
reg = pk.G1Bases(b, mode=pk.G1Bases.TABLE)
torch.cuda.synchronize()
l0 = pk.launch_count()
r1 = pk.variable_base_msm_device(d, reg).cpu().numpy()
l1 = pk.launch_count()
r2 = pk.variable_base_msm_device(d, reg).cpu().numpy()
assert (r1 == r2).all()
print(f"k={k} launches per MSM (incl. scans, memsets excluded): {l1 - l0}", flush=True)
