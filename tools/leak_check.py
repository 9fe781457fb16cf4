"""Repeats setup -> batch_commit(keep) -> eq_table -> sum check -> merge -> open -> release, then a Zeromorph and a Gemini open
over a univariate SRS with prefix slices, and watches free device memory."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import plonkish_b200 as pk
from plonkish_b200 import gemini, kzg, sumcheck, zeromorph
from plonkish_b200.transcript import Keccak256Transcript
from bench import g1_generator

k = 14
one = sumcheck._to_mont(1)
free0 = None
for it in range(60):
    pp = kzg.setup(g1_generator(np), pk.random_scalars(k, seed=it))
    polys = [pk.random_scalars(1 << k, seed=100 + j) for j in range(3)]
    comms, res = kzg.batch_commit(pp, polys, keep=True)
    eq = pk.eq_table(pk.random_scalars(k, seed=7))
    prover = sumcheck.SumCheckProver([eq] + res, [(one, [1, 2]), (one, [3])], common=0)
    for _ in range(k):
        prover.round_evals(); prover.fix_var(one)
    prover.final_evals(); prover.free()
    g = kzg.linear_combination(res, pk.random_scalars(3, seed=9))
    kzg.open_resident(pp, g, pk.random_scalars(k, seed=11))
    srs = kzg.univariate_setup(g1_generator(np), one, 1 << 10); srs.release()
    powers = kzg.univariate_setup(g1_generator(np), pk.random_scalars(1, seed=it)[0], 1 << k)
    zpp = zeromorph.trim(powers, 1 << k, prefix_tables=(it % 2 == 0))
    t = Keccak256Transcript()
    point = t.squeeze_challenges(k)
    zeromorph.open(zpp, res[0], point, 0, t)
    gemini.open(gemini.GeminiKzgProverParam(powers), res[1], point, t)
    assert not gemini.GpuOps._quotients
    zpp.release()
    pk.fixed_base_msm(g1_generator(np), polys[0][:100])
    for r in res + [eq, g]:
        r.release()
    pp.release()
    torch.cuda.synchronize()
    free, total = torch.cuda.mem_get_info()
    if it == 5:
        free0 = free
    if it in (5, 20, 40, 59):
        print(f"iteration {it}: free {free / 2**20:.0f} MiB", flush=True)
assert free0 - free < 64 * 2**20, f"device memory shrank by {(free0 - free) / 2**20:.0f} MiB over 54 iterations"
print("no leak: free device memory stable")
