"""GPU parity (-m gpu) for the callers either side of the MSM:
resident polynomials, `quotients` + open on the device, the g_prime merge, fixed-base MSM
and the device-built SRS — each against the oracle's restatement of the reference lines,
bit-exact."""
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


def _fr(oracle, v):
    return oracle.from_canonical(1, oracle.int_to_limbs(v % R))[0]


def _eq_points(oracle, ss):
    """eqs[k] as affine points through the oracle (kzg.rs:174-208)."""
    g = oracle.generator()
    return [oracle.fixed_base_msm(g, e) for e in oracle.kzg_eq_scalars(ss)]


def test_fixed_base_msm_matches_oracle(pk, oracle):
    # fixed_base_msm + batch_normalize (msm.rs:16-31, 50-81; kzg.rs:204-207)
    g = oracle.generator()
    sc = oracle.random_scalars(5000, 21)
    edge = [0, 1, 2, R - 1, R - 2, 0x8000, 0x8001, 0x7FFF, 0xFFFF, 0x10000, (1 << 253) + 0xFFFFFFFF, (R - 1) // 2, (R + 1) // 2,
            int("8000" * 15, 16), int("7fff" * 15, 16), int("ffff" * 15, 16)]
    for i, v in enumerate(edge):
        sc[i] = _fr(oracle, v)
    got = pk.fixed_base_msm(g, sc)
    assert got.tobytes() == oracle.fixed_base_msm(g, sc).tobytes()
    assert not got[0].any()  # 0 * G = identity = (0, 0)
    assert got[1].tobytes() == g.tobytes()
    # another base (a multiple of G), and the identity as base
    h = oracle.scalar_mul(g, 0xDEADBEEFCAFE)
    assert pk.fixed_base_msm(h, sc[:300]).tobytes() == oracle.fixed_base_msm(h, sc[:300]).tobytes()
    assert not pk.fixed_base_msm(np.zeros(8, dtype=np.uint64), sc[:40]).any()
    # n = 1 and an empty call
    assert pk.fixed_base_msm(g, sc[5:6]).tobytes() == oracle.fixed_base_msm(g, sc[5:6]).tobytes()
    assert pk.fixed_base_msm(g, np.zeros((0, 4), dtype=np.uint64)).shape == (0, 8)


def test_fixed_base_msm_spans_several_launch_batches(pk, oracle):
    # more than one 2^22 batch; checked by sampling (each output is independent)
    g = oracle.generator()
    n = (1 << 22) + 12345
    sc = oracle.random_scalars(n, 22)
    got = pk.fixed_base_msm(g, sc)
    idx = np.concatenate([np.arange(0, 64), np.arange((1 << 22) - 32, (1 << 22) + 32), np.arange(n - 64, n)])
    assert got[idx].tobytes() == oracle.fixed_base_msm(g, sc[idx]).tobytes()


@pytest.mark.parametrize("k", [0, 1, 4, 11])
def test_device_built_srs_matches_oracle(pk, oracle, k):
    from plonkish_b200 import kzg

    ss = oracle.random_scalars(max(k, 1), 30 + k)[:k]
    pp = kzg.setup(oracle.generator(), ss)
    want = _eq_points(oracle, ss)
    assert pp.num_vars() == k
    for i in range(k + 1):
        assert len(pp.eq(i)) == 1 << i
        assert pp.eq(i).to_host().tobytes() == want[i].tobytes(), i
    # the device-built slices commit like host-registered ones
    evals = oracle.random_scalars(1 << k, 40 + k)
    assert kzg.commit(pp, evals).tobytes() == oracle.variable_base_msm(evals, want[k]).tobytes()
    pp.release()


@pytest.mark.parametrize("k", [0, 1, 3, 10, 11, 14])
def test_open_on_resident_polynomial_matches_oracle(pk, oracle, k):
    # kzg.rs:276-302 with `quotients` (multilinear.rs:72-107) on the GPU
    from plonkish_b200 import kzg

    ss = oracle.random_scalars(max(k, 1), 50 + k)[:k]
    pp = kzg.setup(oracle.generator(), ss)
    eqs_host = [pp.eq(i).to_host() for i in range(k + 1)]
    evals = oracle.random_scalars(1 << k, 60 + k)
    point = oracle.random_scalars(max(k, 1), 70 + k)[:k]
    poly = pk.ResidentScalars(evals)
    assert poly.to_host().tobytes() == evals.tobytes()
    assert kzg.commit(pp, poly).tobytes() == oracle.variable_base_msm(evals, eqs_host[k]).tobytes()
    comms, value = kzg.open_resident(pp, poly, point)
    if k:
        qs, want_value = oracle.quotients(evals, point)
        assert value.tobytes() == want_value.tobytes()
        assert len(comms) == k
        for i in range(k):
            assert comms[i].tobytes() == oracle.variable_base_msm(qs[i], eqs_host[i]).tobytes(), i
        # and the host-scalar open (many entry) agrees
        host_comms = pk.variable_base_msm_many(qs, [pp.eq(i) for i in range(k)])
        assert host_comms.tobytes() == np.stack(comms).tobytes()
    else:
        assert value.tobytes() == evals[0].tobytes() and len(comms) == 0
    # opening does not disturb the resident polynomial
    assert poly.to_host().tobytes() == evals.tobytes()
    poly.release()
    pp.release()


def test_open_satisfies_the_kzg_identity_in_the_exponent(pk, oracle):
    # With eqs[k][j] = eq_j(s) * G the commitments are evaluations at s in the exponent, and the
    # verifier's pairing check (kzg.rs:330-361) reduces to
    #   f(s) - f(x) = sum_i (s_i - x_i) * q_i(s_0..s_{i-1})       (mod r)
    # which pins open() at a size the CPU port would need minutes for.
    from plonkish_b200 import kzg

    k = 18
    g = oracle.generator()
    ss = oracle.random_scalars(k, 81)
    pp = kzg.setup(g, ss)
    evals = oracle.random_scalars(1 << k, 82)
    point = oracle.random_scalars(k, 83)
    (comm,), (poly,) = kzg.batch_commit(pp, [evals], keep=True)
    comms, value = kzg.open_resident(pp, poly, point)
    # right-hand side in the exponent: sum_i (s_i - x_i) * Q_i, as one small MSM through the oracle
    diffs = np.stack([oracle.fe_op("sub", 1, ss[i], point[i]) for i in range(k)])
    rhs = oracle.variable_base_msm(diffs, np.stack(comms))
    # left-hand side: C - f(x) * G
    neg_value = oracle.fe_op("sub", 1, np.zeros(4, dtype=np.uint64), value)
    one = _fr(oracle, 1)
    lhs = oracle.variable_base_msm(np.stack([one, neg_value]), np.stack([comm, g]))
    assert lhs.tobytes() == rhs.tobytes()
    # f(x) itself: the multilinear evaluation by the oracle's fold
    _, want_value = oracle.quotients(evals, point)
    assert value.tobytes() == want_value.tobytes()
    poly.release()
    pp.release()


def test_g_prime_merge_matches_oracle(pk, oracle):
    # pcs/multilinear.rs:203-213 over resident polynomials, incl. more terms than one launch takes
    n = 3000
    for count in (1, 4, 13, 27):
        polys = [oracle.random_scalars(n, 100 + i) for i in range(count)]
        coeffs = oracle.random_scalars(count, 99)
        coeffs[0] = _fr(oracle, 1)
        res = [pk.ResidentScalars(p) for p in polys]
        merged = pk.fr_linear_combination(res, coeffs)
        assert merged.to_host().tobytes() == oracle.fr_linear_combination(polys, coeffs).tobytes()
        for r in res + [merged]:
            r.release()


def test_batch_commit_keep_then_merge_then_open(pk, oracle):
    # the prove-time sequence (backend/hyperplonk.rs:201,251 commit; :287 batch_open's final open)
    from plonkish_b200 import kzg

    k = 12
    ss = oracle.random_scalars(k, 111)
    pp = kzg.setup(oracle.generator(), ss)
    eq_k = pp.eq(k).to_host()
    polys = [oracle.random_scalars(1 << k, 120 + i) for i in range(3)]
    comms, resident = kzg.batch_commit(pp, polys, keep=True)
    for p, c, r in zip(polys, comms, resident):
        assert c.tobytes() == oracle.variable_base_msm(p, eq_k).tobytes()
        assert r.to_host().tobytes() == p.tobytes()
    coeffs = oracle.random_scalars(3, 130)
    g_prime = kzg.linear_combination(resident, coeffs)
    g_prime_host = oracle.fr_linear_combination(polys, coeffs)
    point = oracle.random_scalars(k, 131)
    q_comms, value = kzg.open_resident(pp, g_prime, point)
    qs, want_value = oracle.quotients(g_prime_host, point)
    assert value.tobytes() == want_value.tobytes()
    for i in range(k):
        assert q_comms[i].tobytes() == oracle.variable_base_msm(qs[i], pp.eq(i).to_host()).tobytes()
    # additivity of the commitment: commit(g_prime) = sum coeffs[i] * commit(poly_i)  (pcs.rs:166-177)
    assert kzg.commit(pp, g_prime).tobytes() == oracle.variable_base_msm(coeffs, np.stack(comms)).tobytes()
    for r in resident + [g_prime]:
        r.release()
    pp.release()


def test_univariate_setup_matches_powers_times_generator(pk, oracle):
    # UnivariateKzg::setup, G1 half (pcs/univariate/kzg.rs:175-195): powers_of_s_g1[i] = s^i * g1; commit_coeffs on it
    from plonkish_b200 import kzg

    g = oracle.generator()
    s = oracle.random_scalars(1, 301)[0]
    rinv = pow(br.MONT, -1, R)
    si = int.from_bytes(s.tobytes(), "little") * rinv % R
    for n in (1, 5, 300, 1 << 12):
        srs = kzg.univariate_setup(g, s, n)
        powers = np.stack([_fr(oracle, pow(si, i, R)) for i in range(n)])
        want = oracle.fixed_base_msm(g, powers)
        assert srs.to_host().tobytes() == want.tobytes(), n
        coeffs = oracle.random_scalars(n, 302)
        assert kzg.commit_coeffs(srs, coeffs).tobytes() == oracle.variable_base_msm(coeffs, want).tobytes()
        if n > 4:  # a prefix, as trim hands out (univariate/kzg.rs:217-229)
            assert kzg.commit_coeffs(srs, coeffs[: n // 2]).tobytes() == oracle.variable_base_msm(coeffs[: n // 2], want[: n // 2]).tobytes()
        srs.release()


def test_eq_table_matches_the_product_formula_and_the_oracle(pk, oracle):
    # MultilinearPolynomial::eq_xy (poly/multilinear.rs:91-130): evals[b] = prod_i (b_i ? y_i : 1 - y_i)
    for k in (0, 1, 5, 12):
        y = oracle.random_scalars(max(k, 1), 200 + k)[:k]
        table = pk.eq_table(y)
        got = table.to_host()
        assert len(table) == 1 << k
        want = oracle.kzg_eq_scalars(y)[k] if k else _fr(oracle, 1).reshape(1, 4)  # same recurrence as the SRS tables (kzg.rs:178-192)
        assert got.tobytes() == np.ascontiguousarray(want).tobytes()
        if k == 5:
            rinv = pow(br.MONT, -1, R)
            yi = [int.from_bytes(row.tobytes(), "little") * rinv % R for row in y]
            for b in (0, 1, 6, 21, 31):
                v = 1
                for i in range(k):
                    v = v * (yi[i] if (b >> i) & 1 else (1 - yi[i])) % R
                assert int.from_bytes(got[b].tobytes(), "little") * rinv % R == v
        table.release()


def test_argument_errors(pk, oracle):
    from plonkish_b200 import _lib, kzg

    ss = oracle.random_scalars(3, 141)
    pp = kzg.setup(oracle.generator(), ss)
    poly = pk.ResidentScalars(oracle.random_scalars(8, 142))
    with pytest.raises(ValueError):
        kzg.open_resident(pp, poly, oracle.random_scalars(2, 143))  # wrong number of variables
    with pytest.raises(_lib.PlonkishCudaError):
        pp.eq(3).to_host(offset=4, n=8)  # range past the slice
    with pytest.raises(_lib.PlonkishCudaError):
        poly.to_host(offset=9, n=0)
    poly.release()
    with pytest.raises(_lib.PlonkishCudaError):
        _lib.check(_lib.lib().plonkish_cuda_scalars_release(987654321), "scalars_release")
    pp.release()


def test_widened_rows_on_a_second_device(pk, oracle):
    # every handle carries its device: the same flow on cuda:1 while cuda:0 holds other state
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from plonkish_b200 import kzg
    from plonkish_b200.sumcheck import SumCheckProver

    k = 10
    ss = oracle.random_scalars(k, 401)
    pp0 = kzg.setup(oracle.generator(), ss, device=0)  # state on device 0 as well
    pp = kzg.setup(oracle.generator(), ss, device=1)
    eqs = [pp.eq(i).to_host() for i in range(k + 1)]
    assert eqs[k].tobytes() == pp0.eq(k).to_host().tobytes()
    polys = [oracle.random_scalars(1 << k, 410 + j) for j in range(2)]
    res = [pk.ResidentScalars(p, device=1) for p in polys]
    for p, r in zip(polys, res):
        assert kzg.commit(pp, r).tobytes() == oracle.variable_base_msm(p, eqs[k]).tobytes()
    coeffs = oracle.random_scalars(2, 420)
    merged = kzg.linear_combination(res, coeffs)
    assert merged.device == 1
    merged_h = oracle.fr_linear_combination(polys, coeffs)
    point = oracle.random_scalars(k, 421)
    comms, value = kzg.open_resident(pp, merged, point)
    qs, want = oracle.quotients(merged_h, point)
    assert value.tobytes() == want.tobytes()
    assert all(c.tobytes() == oracle.variable_base_msm(q, eqs[i]).tobytes() for i, (c, q) in enumerate(zip(comms, qs)))
    eq = pk.eq_table(point, device=1)
    terms = [(_fr(oracle, 1), [1, 2])]
    prover = SumCheckProver([eq] + res, terms, common=0)
    assert prover.round_evals().tobytes() == oracle.sumcheck_round([eq.to_host()] + polys, terms, 0).tobytes()
    prover.free()
    # mixing devices is refused
    with pytest.raises(Exception):
        kzg.open_resident(pp0, merged, point)
    assert pk.fixed_base_msm(oracle.generator(), polys[0][:50], device=1).tobytes() == oracle.fixed_base_msm(oracle.generator(), polys[0][:50]).tobytes()
    for r in res + [merged, eq]:
        r.release()
    pp.release()
    pp0.release()


@pytest.mark.parametrize("k,shape", [(3, "single"), (6, "mixed"), (9, "mixed"), (7, "two-points-one-poly")])
def test_batch_open_writes_the_reference_proof_bytes(pk, oracle, k, shape):
    """additive::batch_open (pcs/multilinear.rs:134-235) through the GPU entry points against a restatement with Python
    integers and the oracle's MSM (tests/batch_open_ref.py): identical transcript bytes — the coefficient messages of the
    degree-2 sum check and the quotient commitments of the final opening."""
    from batch_open_ref import batch_open_reference, to_int, to_mont
    from plonkish_b200 import kzg
    from plonkish_b200.transcript import Keccak256Transcript

    n = 1 << k
    g = oracle.generator()
    ss = pk.random_scalars(k, seed=31)
    pp = kzg.setup(g, ss)
    eqs_host = [e.to_host() for e in pp.eqs]
    rng = np.random.default_rng(k)
    num_polys = {"single": 1, "mixed": 5, "two-points-one-poly": 1}[shape]
    polys_h = [pk.random_scalars(n, seed=70 + i) for i in range(num_polys)]
    polys_i = [[to_int(r) for r in p] for p in polys_h]
    points = [[int(x) for x in rng.integers(1, 1 << 62, k)] for _ in range({"single": 1, "mixed": 3, "two-points-one-poly": 2}[shape])]
    pairs = {"single": [(0, 0)], "two-points-one-poly": [(0, 0), (0, 1)],
             "mixed": [(0, 0), (1, 0), (2, 0), (3, 1), (4, 2), (3, 2), (0, 2)]}[shape]

    def evaluate(poly, pt):
        cur = poly
        for x in pt:
            cur = [(cur[2 * b] + (cur[2 * b + 1] - cur[2 * b]) * x) % br.R for b in range(len(cur) // 2)]
        return cur[0]

    evals = [(p, x, evaluate(polys_i[p], points[x])) for p, x in pairs]
    resident = [pk.ResidentScalars(p) for p in polys_h]
    t_gpu, t_ref = Keccak256Transcript(), Keccak256Transcript()
    for t in (t_gpu, t_ref):
        t.write_field_elements([v for _, _, v in evals])           # the evaluations are in the transcript before batch_open
    kzg.batch_open(pp, k, resident, points, evals, t_gpu)
    challenges, g_prime_eval = batch_open_reference(oracle, eqs_host, k, polys_i, points, evals, t_ref)
    assert t_gpu.into_proof() == t_ref.into_proof()
    assert len(t_gpu.into_proof()) == 32 * len(evals) + k * 3 * 32 + k * 64
    for r_ in resident:
        r_.release()
    pp.release()
