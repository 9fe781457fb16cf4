"""GPU parity suite (-m gpu): the univariate KZG prover (BASELINE config 4's PCS) — UnivariateKzg::commit / open /
batch_open (pcs/univariate/kzg.rs:242-354) through the C ABI against the oracle, and the proofs it writes against the
reference verifier's own equation (kzg.rs:356-417) evaluated in G1 with the setup's trapdoor."""
import numpy as np
import pytest

from oracle import bigint_ref as br

pytestmark = pytest.mark.gpu
R = br.R


@pytest.fixture(scope="module")
def pk():
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import plonkish_b200

    plonkish_b200._lib.lib()
    return plonkish_b200


def _mont(v):
    from plonkish_b200.sumcheck import _to_mont

    return _to_mont(v)


def _ints(arr):
    from plonkish_b200.sumcheck import _to_int

    return [_to_int(row) for row in np.asarray(arr).reshape(-1, 4)]


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 257, 5000, 16385, (1 << 18) + 3])
def test_division_by_a_linear_factor_matches_the_oracle(pk, oracle, n):
    # poly/univariate.rs:144-168 with divisor (X - z): quotient and remainder, bit for bit
    c = pk.random_scalars(n, seed=n)
    z = pk.random_scalars(1, seed=7)[0]
    poly = pk.ResidentScalars(c)
    q, rem = pk.fr_div_linear(poly, z)
    want_q, want_rem = oracle.fr_div_linear(c, z)
    got_q = q.to_host()
    assert rem.tobytes() == want_rem.tobytes()
    assert got_q[: n - 1].tobytes() == want_q.tobytes() and not got_q[n - 1].any()
    q.release()
    poly.release()


def _point(b):
    return br.point_from_bytes(np.ascontiguousarray(b, dtype=np.uint64).tobytes())


def _horner(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def _proof_points(proof: bytes):
    """write_commitment appends x || y as 32-byte big-endian integers (util/transcript.rs:216-229)."""
    assert len(proof) % 64 == 0
    return [(int.from_bytes(proof[i:i + 32], "big"), int.from_bytes(proof[i + 32:i + 64], "big")) for i in range(0, len(proof), 64)]


from univariate_verify import as_limbs as _as_limbs  # noqa: E402


@pytest.mark.parametrize("log_n", [6, 10, 14])
def test_open_and_batch_open_satisfy_the_verifier(pk, oracle, log_n):
    from plonkish_b200 import kzg, univariate
    from plonkish_b200.transcript import Keccak256Transcript

    n = 1 << log_n
    s = 0x1234567890ABCDEF1234567890ABCDEF % R                       # the trapdoor of this test's setup
    g = oracle.generator()
    srs = kzg.univariate_setup(g, _mont(s), n)                        # powers_of_s_g1 on the device (kzg.rs:175-195)
    polys_h = [pk.random_scalars(n, seed=900 + i) for i in range(4)]
    polys_h[3][n // 2:] = 0                                           # a polynomial of lower degree, zero padded
    coeffs = [_ints(p) for p in polys_h]
    polys = [pk.ResidentScalars(p) for p in polys_h]
    comms = [univariate.commit(srs, p) for p in polys]
    G = br.G
    for c, cf in zip(comms, coeffs):                                   # commit(f) = f(s) * G
        assert _point(c) == br.scalar_mul(_horner(cf, s), G)

    # ---- open (kzg.rs:264-299) and verify (kzg.rs:356-367): e(pi * z + C - eval * G, -G2) e(pi, s G2) = 1  <=>  (s - z) pi = C - eval G
    z = 0xDEADBEEFCAFE % R
    t = Keccak256Transcript()
    eval_ = univariate.open(srs, polys[0], z, t)
    assert eval_ == _horner(coeffs[0], z)
    (pi,) = _proof_points(t.into_proof())
    lhs = br.scalar_mul((s - z) % R, pi)
    rhs = br.add(_point(comms[0]), br.neg(br.scalar_mul(eval_, G)))
    assert lhs == rhs
    # the quotient commitment against the oracle's division + MSM, byte for byte
    want_q, _ = oracle.fr_div_linear(polys_h[0], _mont(z))
    srs_h = srs.to_host()
    assert _as_limbs(pi).tobytes() == oracle.variable_base_msm(want_q, srs_h[: n - 1]).tobytes()

    # ---- batch_open (kzg.rs:301-354) and batch_verify (kzg.rs:380-417)
    points = [0x1111 % R, 0x2222222222 % R, (R - 5)]
    evals = [(0, 0), (0, 1), (1, 0), (2, 2), (3, 0), (3, 1), (1, 0)]   # (poly, point); the last one repeats an entry
    evals = [(p, x, _horner(coeffs[p], points[x])) for p, x in evals]
    t = Keccak256Transcript()
    t.write_commitments(comms)
    univariate.batch_open(srs, polys, points, evals, t)
    proof = t.into_proof()
    pts = _proof_points(proof)
    assert len(pts) == 4 + 2
    q_comm, pi = pts[4], pts[5]
    from univariate_verify import batch_verify_in_g1

    batch_verify_in_g1(comms, points, evals, q_comm, pi, s)
    for p in polys:
        p.release()
    srs.release()
