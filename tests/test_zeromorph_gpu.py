"""GPU parity suite (-m gpu): Zeromorph<UnivariateKzg<Bn256>> (pcs/multilinear/zeromorph.rs) through the C ABI — the
kept quotients, their commitments from resident sub-ranges, q_hat and f against the oracle / Python integers, and the
proofs of open / batch_open byte for byte against the loop-for-loop restatement (tests/zeromorph_ref.py) and the
verifier's equation in G1 with the setup's trapdoor."""
import numpy as np
import pytest

from oracle import bigint_ref as br
from univariate_verify import as_limbs
import zeromorph_ref as zr

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


def _proof_limbs(proof: bytes):
    assert len(proof) % 64 == 0
    return [as_limbs((int.from_bytes(proof[i:i + 32], "big"), int.from_bytes(proof[i + 32:i + 64], "big"))) for i in range(0, len(proof), 64)]


@pytest.mark.parametrize("num_vars", [0, 1, 2, 5, 10, 11, 13, 17])
def test_kept_quotients_q_hat_and_f_match_the_oracle(pk, oracle, num_vars):
    # pcs/multilinear.rs:72-107 (packed: q_i at element offset 2^i), zeromorph.rs:157-168 and :175-180
    n = 1 << num_vars
    poly_h = pk.random_scalars(n, seed=300 + num_vars)
    point = pk.random_scalars(max(num_vars, 1), seed=301)[:num_vars]
    poly = pk.ResidentScalars(poly_h)
    q, value = pk.fr_quotients(poly, point)
    got_q = q.to_host()
    if num_vars:
        want_qs, want_value = oracle.quotients(poly_h, point)
        assert value.tobytes() == want_value.tobytes()
        for i, wq in enumerate(want_qs):
            assert got_q[1 << i: 2 << i].tobytes() == wq.tobytes(), i
    else:
        assert value.tobytes() == poly_h[0].tobytes()
    assert not got_q[0].any()
    w = pk.random_scalars(max(num_vars, 1), seed=302)[:num_vars]
    s = pk.random_scalars(max(num_vars, 1), seed=303)[:num_vars]
    z, c0 = pk.random_scalars(2, seed=304)
    q_hat = pk.zeromorph_q_hat(q, w)
    f = pk.zeromorph_f(poly, q_hat, q, z, c0, s)
    # the same sums through the oracle's vector operations
    want_hat = np.zeros((n, 4), dtype=np.uint64)
    want_f = oracle.fr_linear_combination([poly_h], z.reshape(1, 4))
    for i in range(num_vars):
        seg = got_q[1 << i: 2 << i]
        lo = n - (1 << i)
        want_hat[lo:] = oracle.fr_vec_op("add", want_hat[lo:], oracle.fr_linear_combination([seg], w[i].reshape(1, 4)))
        want_f[: 1 << i] = oracle.fr_vec_op("add", want_f[: 1 << i], oracle.fr_linear_combination([seg], s[i].reshape(1, 4)))
    want_f = oracle.fr_vec_op("add", want_f, want_hat)
    want_f[0] = oracle.fr_vec_op("add", want_f[:1], c0.reshape(1, 4))[0]
    assert q_hat.to_host().tobytes() == want_hat.tobytes()
    assert f.to_host().tobytes() == want_f.tobytes()
    for x in (poly, q, q_hat, f):
        x.release()


@pytest.mark.parametrize("num_vars", [1, 4, 9, 12, 16])
def test_quotient_commitments_from_resident_sub_ranges(pk, oracle, num_vars):
    # UnivariateKzg::batch_commit_and_write over the quotients (zeromorph.rs:150): q_i against powers_of_s_g1[..2^i]
    from plonkish_b200 import kzg

    n = 1 << num_vars
    srs = kzg.univariate_setup(oracle.generator(), _mont(0xBADC0FFEE), n)
    srs_h = srs.to_host()
    poly = pk.ResidentScalars(pk.random_scalars(n, seed=400 + num_vars))
    q, _ = pk.fr_quotients(poly, pk.random_scalars(num_vars, seed=401))
    sizes = [1 << i for i in range(num_vars)]
    got = pk.variable_base_msm_many_resident(q, sizes, [srs] * num_vars, sizes)
    q_h = q.to_host()
    for i in range(num_vars):
        assert got[i].tobytes() == oracle.variable_base_msm(q_h[1 << i: 2 << i], srs_h[: 1 << i]).tobytes(), i
    # ragged: unaligned offsets, a zero-length entry, the whole vector
    offs, ns = [0, 3, n - 1, 1], [n, min(5, n - 3) if n > 3 else 0, 1, 0]
    got = pk.variable_base_msm_many_resident(q, offs, [srs] * 4, ns)
    for j, (o, m) in enumerate(zip(offs, ns)):
        want = oracle.variable_base_msm(q_h[o: o + m], srs_h[:m]) if m else np.zeros(8, dtype=np.uint64)
        assert got[j].tobytes() == want.tobytes(), j
    with pytest.raises(pk.PlonkishCudaError):
        pk.variable_base_msm_many_resident(q, [n - 1], [srs], [2])
    for x in (poly, q):
        x.release()
    srs.release()


@pytest.mark.parametrize("num_vars,extra", [(1, 0), (2, 0), (3, 0), (6, 0), (10, 0), (8, 37), (13, 0)])
def test_open_writes_the_reference_proof_bytes_and_verifies(pk, oracle, num_vars, extra):
    # run_commit_open_verify (pcs/multilinear.rs:293-335) for Zeromorph: commit, squeeze the point, write the evaluation,
    # open; extra > 0: setup longer than poly_size, open_pp = the last 2^num_vars powers (trim, zeromorph.rs:84-102)
    from plonkish_b200 import kzg, zeromorph
    from plonkish_b200.transcript import Keccak256Transcript

    n, s = 1 << num_vars, 0x1F2E3D4C5B6A79881 % R
    full = kzg.univariate_setup(oracle.generator(), _mont(s), n + extra)
    pp = zeromorph.trim(full, n)
    assert pp.degree() == n - 1
    commit_h, open_h = pp.commit_pp.to_host(), pp.open_pp.to_host()
    poly_h = pk.random_scalars(n, seed=500 + num_vars)
    evals = _ints(poly_h)
    poly = pk.ResidentScalars(poly_h)
    t, t_ref = Keccak256Transcript(), Keccak256Transcript()
    comm = zeromorph.commit(pp, poly)
    assert comm.tobytes() == zr.commit_coeffs(oracle, commit_h, evals).tobytes()
    for tr in (t, t_ref):
        tr.write_commitment(comm)
    point = t.squeeze_challenges(num_vars)
    assert t_ref.squeeze_challenges(num_vars) == point
    value = _ints(oracle.evaluate_multilinear(poly_h, zr.mont_rows(point)))[0]
    for tr in (t, t_ref):
        tr.write_field_element(value)
    remainder = zeromorph.open(pp, poly, point, value, t)
    want_remainder, f_at_x = zr.open_reference(oracle, commit_h, open_h, evals, point, value, t_ref)
    assert remainder == want_remainder == value and f_at_x == 0
    proof = t.into_proof()
    assert proof == t_ref.into_proof()
    pts = _proof_limbs(proof[64 + 32:])
    assert len(pts) == num_vars + 2
    v = Keccak256Transcript()
    v.write_commitment(comm)
    v.squeeze_challenges(num_vars)
    v.write_field_element(value)
    zr.verify_in_g1(comm, point, value, pts[:num_vars], pts[num_vars], pts[num_vars + 1], v, s, extra)
    poly.release()
    if extra:
        pp.release()
    full.release()


@pytest.mark.parametrize("num_vars", [16, 18])
def test_open_of_a_larger_polynomial_verifies(pk, oracle, num_vars):
    # sizes past the restatement's reach: the verifier's equation alone
    from plonkish_b200 import kzg, zeromorph
    from plonkish_b200.transcript import Keccak256Transcript

    n, s = 1 << num_vars, 0x7777777123456789ABCDEF % R
    pp = zeromorph.trim(kzg.univariate_setup(oracle.generator(), _mont(s), n), n)
    poly_h = pk.random_scalars(n, seed=600 + num_vars)
    poly = pk.ResidentScalars(poly_h)
    t = Keccak256Transcript()
    comm = zeromorph.commit(pp, poly)
    t.write_commitment(comm)
    point = t.squeeze_challenges(num_vars)
    value = _ints(oracle.evaluate_multilinear(poly_h, zr.mont_rows(point), oracle.host_threads()))[0]
    t.write_field_element(value)
    assert zeromorph.open(pp, poly, point, value, t) == value
    pts = _proof_limbs(t.into_proof()[64 + 32:])
    v = Keccak256Transcript()
    v.write_commitment(comm)
    v.squeeze_challenges(num_vars)
    v.write_field_element(value)
    zr.verify_in_g1(comm, point, value, pts[:num_vars], pts[num_vars], pts[num_vars + 1], v, s, 0)
    poly.release()
    pp.release()


@pytest.mark.parametrize("num_vars", [3, 7])
def test_batch_open_writes_the_reference_proof_bytes(pk, oracle, num_vars):
    # run_batch_commit_open_verify's shape (pcs/multilinear.rs:337-406): several polynomials, several points;
    # additive::batch_open with Zeromorph's open on g_prime (zeromorph.rs:188-204)
    from batch_open_ref import batch_open_reference, to_int
    from plonkish_b200 import kzg, zeromorph
    from plonkish_b200.transcript import Keccak256Transcript

    n, s = 1 << num_vars, 0x4242424242424242DEADBEEF % R
    pp = zeromorph.trim(kzg.univariate_setup(oracle.generator(), _mont(s), n), n)
    srs_h = pp.commit_pp.to_host()
    polys_h = [pk.random_scalars(n, seed=700 + i) for i in range(4)]
    polys_i = [[to_int(r) for r in p] for p in polys_h]
    rng = np.random.default_rng(num_vars)
    points = [[int(x) for x in rng.integers(1, 1 << 62, num_vars)] for _ in range(3)]
    pairs = [(0, 0), (0, 1), (1, 0), (2, 2), (3, 1), (3, 0)]

    def evaluate(poly, pt):
        cur = poly
        for x in pt:
            cur = [(cur[2 * b] + (cur[2 * b + 1] - cur[2 * b]) * x) % R for b in range(len(cur) // 2)]
        return cur[0]

    evals = [(p, x, evaluate(polys_i[p], points[x])) for p, x in pairs]
    resident = [pk.ResidentScalars(p) for p in polys_h]
    t_gpu, t_ref = Keccak256Transcript(), Keccak256Transcript()
    for tr in (t_gpu, t_ref):
        tr.write_commitments(zeromorph.batch_commit(pp, resident))
        tr.write_field_elements([v for _, _, v in evals])
    zeromorph.batch_open(pp, num_vars, resident, points, evals, t_gpu)

    def open_ref(g_prime, challenges, transcript):
        return zr.open_reference(oracle, srs_h, srs_h, g_prime, challenges, 0, transcript)   # eval = ZERO, multilinear.rs:224-226

    batch_open_reference(oracle, None, num_vars, polys_i, points, evals, t_ref, open_fn=open_ref)
    assert t_gpu.into_proof() == t_ref.into_proof()
    for r_ in resident:
        r_.release()
    pp.release()


def test_prefix_tables_leave_the_proofs_unchanged(pk, oracle):
    # zeromorph.build_prefix_tables: own resident slices for &powers_of_s_g1[..2^i] change the MSMs' window sizes, not a byte
    from plonkish_b200 import gemini, kzg, zeromorph
    from plonkish_b200.transcript import Keccak256Transcript

    num_vars = 11
    n, s = 1 << num_vars, 0x1CEB00DA % R
    poly = pk.ResidentScalars(pk.random_scalars(n, seed=4321))
    point = [int(x) for x in np.random.default_rng(5).integers(1, 1 << 62, num_vars)]
    proofs = {}
    for with_tables in (False, True):
        powers = kzg.univariate_setup(oracle.generator(), _mont(s), n)
        zpp = zeromorph.trim(powers, n, prefix_tables=with_tables)
        assert len(getattr(powers, "prefix_tables", {})) == (num_vars if with_tables else 0)
        t = Keccak256Transcript()
        zeromorph.open(zpp, poly, point, 0, t)
        gemini.open(gemini.GeminiKzgProverParam(powers), poly, point, t)   # finds the tables on the same slice
        proofs[with_tables] = t.into_proof()
        zpp.release()
    assert proofs[False] == proofs[True]
    poly.release()


@pytest.mark.parametrize("idx", [0, 1])
def test_golden_proofs_through_the_c_abi(pk, idx):
    # the committed known-answer proofs (tests/golden/pcs_vectors.json, Python integers only) through the GPU path
    from plonkish_b200 import zeromorph
    from plonkish_b200.transcript import Keccak256Transcript
    from test_zeromorph_cpu import _golden_cases, golden_inputs

    case = _golden_cases()[idx]
    powers, evals, point, value, proof = golden_inputs(case)
    n = 1 << case["num_vars"]
    full = pk.G1Bases(powers)
    pp = zeromorph.trim(full, n)
    poly = pk.ResidentScalars(evals)
    t = Keccak256Transcript()
    t.write_commitment(zeromorph.commit(pp, poly))
    assert t.squeeze_challenges(case["num_vars"]) == point
    t.write_field_element(value)
    assert zeromorph.open(pp, poly, point, value, t) == value
    assert t.into_proof() == proof
    poly.release()
    if case["extra"]:
        pp.release()
    full.release()
