"""Mid-size table-mode MSMs (the tiered k_accumulate runs, k_combine_single) and a factored zero-check round, for
compute-sanitizer memcheck / racecheck."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk
from oracle import pyoracle as po
from plonkish_b200 import sumcheck

for n in (20000, 70001):
    sc = po.random_scalars(n, n)
    d = torch.from_numpy(sc.view(np.int64)).cuda()
    b = pk.synth_bases_device(n, 3, 5)
    reg = pk.G1Bases(b, mode=pk.G1Bases.TABLE)
    want = po.known_dlog_answer(3, 5, sc)
    assert (pk.variable_base_msm_device(d, reg).cpu().numpy().view(np.uint64) == want).all()
    for tiers in ("1", "3"):
        os.environ["PLONKISH_CUDA_ACC_TIERS"] = tiers
        assert (pk.variable_base_msm_device(d, reg).cpu().numpy().view(np.uint64) == want).all()
    os.environ.pop("PLONKISH_CUDA_ACC_TIERS")
    reg.release()
k = 8
tables = [pk.ResidentScalars(pk.random_scalars(1 << k, seed=5 + i)) for i in range(4)]
one = sumcheck._to_mont(1)
prover = sumcheck.SumCheckProver(tables, [(one, [1, 2]), (one, [3])], common=0)
a = prover.round_evals(); g = prover.round_evals_factored()
assert a.shape[0] == g.shape[0] + 1
prover.free()
print("sanitize tiers ok")
