"""Small MSMs through every entry point, for compute-sanitizer memcheck."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk
from oracle import pyoracle as po
for n in (1, 33, 1000, 5000):
    sc = po.random_scalars(n, n); bs = po.known_dlog_bases(3, 5, n); want = po.known_dlog_answer(3, 5, sc)
    assert (pk.variable_base_msm(sc, bs) == want).all()
    for mode in (pk.G1Bases.TABLE, pk.G1Bases.PLAIN):
        reg = pk.G1Bases(bs, mode=mode)
        assert (pk.variable_base_msm(sc, reg) == want).all()
        os.environ["PLONKISH_CUDA_HOST_CHUNKS"] = "3"
        assert (pk.variable_base_msm(sc, reg) == want).all()
        os.environ.pop("PLONKISH_CUDA_HOST_CHUNKS")
        assert (pk.variable_base_msm_batch([sc, sc], reg)[1] == want).all()
        d = torch.from_numpy(sc.view(np.int64)).cuda()
        assert (pk.variable_base_msm_device(d, reg).cpu().numpy().view(np.uint64) == want).all()
        reg.release()
    assert (pk.variable_base_msm([s for s in sc], [b for b in bs]) == want).all()
d_b = pk.synth_bases_device(3000, 3, 5)
print("sanitize run ok")
