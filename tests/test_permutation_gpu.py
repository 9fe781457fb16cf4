"""GPU parity suite (-m gpu): permutation_z_polys (backend/hyperplonk/prover.rs:252-345) through the C ABI against the
oracle, bit for bit, and against the argument's own recurrence on a satisfied copy constraint."""
import numpy as np
import pytest

from test_permutation_cpu import PRIMITIVES, R, copy_constrained_instance, to_ints, to_mont

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk():
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import plonkish_b200

    plonkish_b200._lib.lib()
    return plonkish_b200


@pytest.mark.parametrize("k,count,num_chunks", [(1, 1, 1), (2, 2, 2), (5, 3, 1), (9, 3, 3), (10, 5, 2), (13, 3, 1), (16, 3, 1), (18, 4, 2), (20, 3, 1), (20, 2, 2)])
def test_z_polys_match_the_oracle(pk, oracle, k, count, num_chunks):
    n = 1 << k
    values_h = [pk.random_scalars(n, seed=10 * k + i) for i in range(count)]
    sigmas_h = [pk.random_scalars(n, seed=500 + 10 * k + i) for i in range(count)]
    beta, gamma = pk.random_scalars(2, seed=77)
    values = [pk.ResidentScalars(v) for v in values_h]
    sigmas = [pk.ResidentScalars(v) for v in sigmas_h]
    got = pk.permutation_z_polys(num_chunks, values, sigmas, beta, gamma)
    want = oracle.permutation_z_polys(num_chunks, values_h, sigmas_h, beta, gamma, num_threads=oracle.host_threads())
    assert len(got) == num_chunks
    for g, w in zip(got, want):
        assert g.to_host().tobytes() == w.tobytes()
    for r in values + sigmas + got:
        r.release()


def test_z_poly_of_a_satisfied_copy_constraint_obeys_the_recurrence(pk):
    # z(next(b)) * prod (beta sigma + gamma + w) = z(b) * prod (beta id + gamma + w) on every row (the constraint of
    # backend/hyperplonk/preprocessor.rs:153-166), z(1) = 1, and the product closes after the last row (prover.rs:322-328)
    k, count = 12, 3
    n = 1 << k
    rng = np.random.default_rng(3)
    vals, sigs = copy_constrained_instance(k, count, rng)
    beta, gamma = 0xABCDEF0123456789 % R, 0x1357924680 % R
    values = [pk.ResidentScalars(to_mont(v)) for v in vals]
    sigmas = [pk.ResidentScalars(to_mont(s)) for s in sigs]
    (z,) = pk.permutation_z_polys(1, values, sigmas, to_mont([beta])[0], to_mont([gamma])[0])
    zi = to_ints(z.to_host())
    assert zi[0] == 0 and zi[1] == 1
    for b in range(1, n):
        nxt = (b << 1) ^ (((b << 1) >> k) * PRIMITIVES[k])
        num = den = 1
        for i in range(count):
            num = num * (beta * ((i << k) + b) + gamma + vals[i][b]) % R
            den = den * (beta * sigs[i][b] + gamma + vals[i][b]) % R
        assert zi[nxt] * den % R == zi[b] * num % R, b
    for r in values + sigmas + [z]:
        r.release()
