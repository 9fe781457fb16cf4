"""Stage times of one 2^k MSM for several table window widths (PLONKISH_CUDA_TABLE_C is read at registration)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import plonkish_b200 as pk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
dev = torch.device("cuda", 0)
d_bases = pk.synth_bases_device(n, 7, 3, device=dev)
sc = torch.from_numpy(pk.random_scalars(n, seed=1).view(np.int64)).to(dev)
torch.cuda.synchronize()
ref = None
for c in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "20,21,22").split(",")]:
    os.environ["PLONKISH_CUDA_TABLE_C"] = str(c)
    reg = pk.G1Bases(d_bases, mode=pk.G1Bases.TABLE)
    out = pk.variable_base_msm_device(sc, reg)
    torch.cuda.synchronize()
    got = out.cpu().numpy().tobytes()
    ref = ref or got
    best = None
    for _ in range(4):
        st = pk.profile_stages_device(sc, reg)
        tot = sum(st.values())
        if best is None or tot < sum(best.values()):
            best = st
    print(c, "same" if got == ref else "DIFFERENT", "total", round(sum(best.values()), 3), {a: round(b, 3) for a, b in best.items()}, flush=True)
    reg.release()
