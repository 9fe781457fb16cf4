"""GPU parity suite (-m gpu): the CUDA path through the C ABI against the oracle,
the committed golden vectors and size-independent identities.  Bit-exact: every
comparison is on the 64 result bytes (integer arithmetic, no tolerance)."""
import ctypes

import numpy as np
import pytest

from conftest import case_arrays
from oracle import bigint_ref as br

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk():
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import plonkish_b200

    plonkish_b200._lib.lib()
    return plonkish_b200


def _limbs32(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint64).reshape(-1, 4).copy()


def _field_op(pk, op, a, b=None):
    from plonkish_b200 import _lib

    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.zeros_like(a)
    bp = None if b is None else np.ascontiguousarray(b, dtype=np.uint64).ctypes.data
    _lib.check(_lib.lib().plonkish_cuda_debug_field_op(0, op, a.ctypes.data, bp, out.ctypes.data, a.shape[0]), "debug_field_op")
    return out


def _ints(arr):
    return [int.from_bytes(row.tobytes(), "little") for row in arr]


def test_ptx_field_arithmetic_matches_python_ints(pk):
    rng = np.random.default_rng(11)
    for op_mul, mod in ((0, br.P), (5, br.R)):
        edge = [0, 1, 2, mod - 1, mod - 2, 1 << 253, (1 << 32) - 1, 1 << 32, mod >> 1]
        vals = edge + [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(4000)]
        a = vals + [v for v in edge for _ in edge]
        b = vals[::-1] + [w for _ in edge for w in edge]
        rinv = pow(br.MONT, -1, mod)
        got = _ints(_field_op(pk, op_mul, _limbs32(a), _limbs32(b)))
        assert got == [x * y * rinv % mod for x, y in zip(a, b)]
    a = [int.from_bytes(rng.bytes(32), "little") % br.P for _ in range(2000)] + [0, br.P - 1, 1]
    b = a[::-1]
    assert _ints(_field_op(pk, 1, _limbs32(a), _limbs32(b))) == [(x + y) % br.P for x, y in zip(a, b)]
    assert _ints(_field_op(pk, 2, _limbs32(a), _limbs32(b))) == [(x - y) % br.P for x, y in zip(a, b)]
    assert _ints(_field_op(pk, 6, _limbs32(a))) == [(-x) % br.P for x in a]
    # fused a*b + c*d with one reduction (ops 8, 9), extremes included: every operand p-1 / r-1
    for op, mod in ((8, br.P), (9, br.R)):
        top = (mod >> 224 << 224) - 1
        edge = [0, 1, mod - 1, mod - 2, top, (1 << 224) - 1, mod - 1, mod - 1]
        x = edge + [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(3000)]
        y = [mod - 1, mod - 1, mod - 1, top, top, 1, 0, mod - 2] + [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(3000)]
        rinv = pow(br.MONT, -1, mod)
        got = _ints(_field_op(pk, op, _limbs32(x), _limbs32(y)))
        n = len(x)
        if op == 8:
            want = [(x[i] * y[i] + y[i] * x[(i + 1) % n]) * rinv % mod for i in range(n)]
        else:
            want = [(x[i] * x[i] + y[i] * y[(i + 1) % n]) * rinv % mod for i in range(n)]
        assert got == want
    # squaring with the symmetric partial products taken once (ops 10, 11)
    for op, mod in ((10, br.P), (11, br.R)):
        top = (mod >> 224 << 224) - 1
        x = [0, 1, 2, mod - 1, mod - 2, top, (1 << 224) - 1, (1 << 253) + ((1 << 32) - 1), int("7fffffff" * 7, 16), int("80000000" * 7, 16),
             int("ffffffff" * 7, 16)] + [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(4000)]
        rinv = pow(br.MONT, -1, mod)
        assert _ints(_field_op(pk, op, _limbs32(x))) == [v * v * rinv % mod for v in x]
    # Fr Montgomery -> canonical (halo2_curves to_repr at msm.rs:153)
    ks = [0, 1, br.R - 1, (br.R - 1) // 2] + [int.from_bytes(rng.bytes(32), "little") % br.R for _ in range(500)]
    mont = np.frombuffer(b"".join(br.scalar_to_bytes(k) for k in ks), dtype=np.uint64).reshape(-1, 4)
    assert _ints(_field_op(pk, 3, mont)) == ks
    inv_in = a[:64] + [0, 1, br.P - 1, 2, (br.P - 1) // 2]
    want_inv = [0 if x == 0 else pow(x, -1, br.P) * br.MONT * br.MONT % br.P for x in inv_in]
    assert _ints(_field_op(pk, 4, _limbs32(inv_in))) == want_inv       # Fermat ladder
    big = a + [0, 1, br.P - 1, 2, (br.P - 1) // 2]
    assert _ints(_field_op(pk, 7, _limbs32(big))) == [0 if x == 0 else pow(x, -1, br.P) * br.MONT * br.MONT % br.P for x in big]  # safegcd


def test_group_law_including_exceptional_cases(pk, oracle):
    from plonkish_b200 import _lib

    n = 64
    pts = oracle.known_dlog_bases(5, 3, n)  # (5 + 3i) G
    one = np.frombuffer(br.fe_to_mont_bytes(1, br.P), dtype=np.uint64)
    a = np.zeros((n, 16), dtype=np.uint64)
    b = np.zeros((n, 16), dtype=np.uint64)
    a[:, :8] = pts
    a[:, 8:12] = one
    a[:, 12:16] = one
    b[:, :8] = pts[::-1]
    b[:, 8:12] = one
    b[:, 12:16] = one
    # exceptional rows: P + P, P + (-P), identity + P, P + identity
    b[0, :8] = pts[0]
    negp = br.point_to_bytes(br.neg(br.point_from_bytes(pts[1].tobytes())))
    b[1, :8] = np.frombuffer(negp, dtype=np.uint64)
    a[2, 8:] = 0
    b[3, :] = 0

    def run(op, x, y):
        out = np.zeros_like(x)
        _lib.check(_lib.lib().plonkish_cuda_debug_point_op(0, op, x.ctypes.data, y.ctypes.data, out.ctypes.data, x.shape[0]), "debug_point_op")
        aff = np.zeros_like(x)
        _lib.check(_lib.lib().plonkish_cuda_debug_point_op(0, 3, out.ctypes.data, out.ctypes.data, aff.ctypes.data, x.shape[0]), "debug_point_op")
        return [br.point_from_bytes(row[:8].tobytes()) for row in aff]

    def pt(row):
        return None if not row[8:12].any() else br.point_from_bytes(row[:8].tobytes())

    def pt_b_affine(row):
        return br.point_from_bytes(row[:8].tobytes())

    want_mixed = [br.add(pt(a[i]), pt_b_affine(b[i])) for i in range(n)]
    assert run(0, a, b) == want_mixed
    want_full = [br.add(pt(a[i]), pt(b[i])) for i in range(n)]
    assert run(1, a, b) == want_full
    assert run(2, a, b) == [br.add(pt(a[i]), pt(a[i])) for i in range(n)]


def test_golden_vectors_through_the_c_abi(pk, golden):
    for case in golden["cases"]:
        sc, bs, want = case_arrays(case)
        got = pk.variable_base_msm(sc, bs)
        assert got.tobytes() == want.tobytes(), case["name"]
        # iterator-of-references callers (pcs/univariate/kzg.rs:346): gather entry point
        got = pk.variable_base_msm([s for s in sc], [b for b in bs])
        assert got.tobytes() == want.tobytes(), case["name"] + " (gather)"


def test_empty_input_is_identity(pk):
    out = pk.variable_base_msm(np.zeros((0, 4), np.uint64), np.zeros((0, 8), np.uint64))
    assert not out.any()


def test_length_mismatch_asserts(pk):
    with pytest.raises(AssertionError):  # msm.rs:90
        pk.variable_base_msm(pk.random_scalars(3, 1), np.zeros((4, 8), np.uint64))


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 1000, 4097, (1 << 14) - 1, 1 << 16, (1 << 16) + 1])
def test_matches_oracle_on_random_inputs(pk, oracle, n):
    sc = oracle.random_scalars(n, 100 + n)
    bs = oracle.known_dlog_bases(7, 11, n)
    want = oracle.variable_base_msm(sc, bs)
    got = pk.variable_base_msm(sc, bs)
    assert got.tobytes() == want.tobytes()
    assert oracle.transcript_bytes(got) == oracle.transcript_bytes(want)  # util/transcript.rs:216-229


@pytest.mark.parametrize("c", [8, 9, 10, 11, 12, 13, 14, 15, 16])
def test_every_window_size_gives_the_same_point(pk, oracle, c):
    import torch

    n = 20000
    sc = oracle.random_scalars(n, 5)
    bs = oracle.known_dlog_bases(3, 5, n)
    want = oracle.known_dlog_answer(3, 5, sc)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_bs = torch.from_numpy(bs.view(np.int64)).cuda()
    got = pk.variable_base_msm_device(d_sc, d_bs, window_bits=c).cpu().numpy().view(np.uint64)
    assert got.tobytes() == want.tobytes()


def test_skewed_scalar_distributions(pk, oracle):
    # SURVEY.md §8(d) skew set: selector columns (backend/hyperplonk/util.rs:133-152),
    # small integers (preprocessor.rs:184-190), one repeated wide value, all -1.
    n = 1 << 15
    rng = np.random.default_rng(3)
    bs = oracle.known_dlog_bases(2, 9, n)

    def mont(vals):
        return np.frombuffer(b"".join(br.scalar_to_bytes(int(v)) for v in vals), dtype=np.uint64).reshape(-1, 4).copy()

    cases = {
        "selector": [[0, 1, br.R - 1][int(x)] for x in rng.integers(0, 3, n)],
        "small": [int(x) for x in rng.integers(0, 3 * n, n)],
        "all-minus-one": [br.R - 1] * n,
        "all-one": [1] * n,
        "all-zero": [0] * n,
        "same-wide": [0x2AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA % br.R] * n,
        "limbs68": [int(x) << 4 for x in rng.integers(0, 1 << 62, n)],
    }
    for name, vals in cases.items():
        sc = mont(vals)
        want = oracle.known_dlog_answer(2, 9, sc)
        assert pk.variable_base_msm(sc, bs).tobytes() == want.tobytes(), name


def test_duplicate_negated_and_identity_bases(pk, oracle):
    n = 4096
    sc = oracle.random_scalars(n, 8)
    base = oracle.known_dlog_bases(4, 1, n)
    dup = np.repeat(base[:1], n, axis=0)  # forces P + P inside buckets
    assert pk.variable_base_msm(sc, dup).tobytes() == oracle.variable_base_msm(sc, dup).tobytes()
    neg = base.copy()
    for i in range(0, 64, 2):  # P, -P pairs with equal scalars cancel
        neg[i + 1] = np.frombuffer(br.point_to_bytes(br.neg(br.point_from_bytes(neg[i].tobytes()))), dtype=np.uint64)
    sc2 = sc.copy()
    sc2[1:64:2] = sc2[0:64:2]
    assert pk.variable_base_msm(sc2, neg).tobytes() == oracle.variable_base_msm(sc2, neg).tobytes()
    ident = base.copy()
    ident[::5] = 0
    assert pk.variable_base_msm(sc, ident).tobytes() == oracle.variable_base_msm(sc, ident).tobytes()


def test_registered_bases_and_prefixes(pk, oracle):
    # One resident SRS, MSMs over prefixes: the shape of MultilinearKzg::open
    # (pcs/multilinear/kzg.rs:291-293, sizes 2^(k-1)..1) and UnivariateKzg::commit_coeffs.
    n = 1 << 13
    bs = oracle.known_dlog_bases(6, 7, n)
    reg = pk.G1Bases(bs)
    for m in (n, n // 2, 1000, 3, 1):
        sc = oracle.random_scalars(m, m)
        assert pk.variable_base_msm(sc, reg).tobytes() == oracle.known_dlog_answer(6, 7, sc).tobytes()
    reg.release()


@pytest.mark.parametrize("n", [1, 33, 1000, 1 << 13, (1 << 16) + 7])
def test_table_of_window_multiples_matches_oracle(pk, oracle, n):
    # Resident bases expanded into T[w][i] = 2^(c*w) * P_i: one bucket set, wider windows.
    import torch

    bs = oracle.known_dlog_bases(6, 7, n)
    reg = pk.G1Bases(bs, mode=pk.G1Bases.TABLE)
    plain = pk.G1Bases(bs, mode=pk.G1Bases.PLAIN)
    for m in sorted({n, max(1, n // 2), max(1, n - 1), 1}):
        sc = oracle.random_scalars(m, m + 13)
        want = oracle.known_dlog_answer(6, 7, sc)
        assert pk.variable_base_msm(sc, reg).tobytes() == want.tobytes(), m
        assert pk.variable_base_msm(sc, plain).tobytes() == want.tobytes(), m
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        assert pk.variable_base_msm_device(d_sc, reg).cpu().numpy().view(np.uint64).tobytes() == want.tobytes(), m
    reg.release()
    plain.release()


def test_table_with_skew_identity_and_duplicate_bases(pk, oracle):
    n = 1 << 12
    rng = np.random.default_rng(9)
    bs = oracle.known_dlog_bases(2, 9, n)
    bs[::9] = 0            # identity bases stay identity in every table row
    bs[5::64] = bs[5]      # repeated bases force P + P inside buckets
    reg = pk.G1Bases(bs, mode=pk.G1Bases.TABLE)

    def mont(vals):
        return np.frombuffer(b"".join(br.scalar_to_bytes(int(v)) for v in vals), dtype=np.uint64).reshape(-1, 4).copy()

    for name, vals in {
        "uniform": None,
        "selector": [[0, 1, br.R - 1][int(x)] for x in rng.integers(0, 3, n)],
        "all-minus-one": [br.R - 1] * n,
        "all-zero": [0] * n,
        "small": [int(x) for x in rng.integers(0, 3 * n, n)],
    }.items():
        sc = oracle.random_scalars(n, 4) if vals is None else mont(vals)
        assert pk.variable_base_msm(sc, reg).tobytes() == oracle.variable_base_msm(sc, bs).tobytes(), name
    reg.release()


def test_table_at_2pow22_known_discrete_log(pk, oracle):
    import torch

    n = 1 << 22
    sc = pk.random_scalars(n, seed=2222)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_bs = pk.synth_bases_device(n, 3, 5)
    reg = pk.G1Bases(d_bs, mode=pk.G1Bases.TABLE)
    got = pk.variable_base_msm_device(d_sc, reg).cpu().numpy().view(np.uint64)
    assert got.tobytes() == oracle.known_dlog_answer(3, 5, sc).tobytes()
    reg.release()


@pytest.mark.parametrize("n", [(1 << 20) + 3, (1 << 22) + 5, 1 << 23])
def test_host_path_chunk_pipeline(pk, oracle, n):
    # Host scalars at 2^20 and above are copied in chunks (two, three from 2^23) that overlap with
    # compute; every chunk fills its own bucket array and one reduce adds them (api.cu enqueue_host_msm).
    sc = pk.random_scalars(n, seed=n % 1000)
    d_bs = pk.synth_bases_device(n, 3, 5)
    want = oracle.known_dlog_answer(3, 5, sc)
    import os

    for mode in (pk.G1Bases.TABLE, pk.G1Bases.PLAIN):
        reg = pk.G1Bases(d_bs, mode=mode)
        for var, chunks in (("", ""), ("PLONKISH_CUDA_HOST_CHUNKS", "3"), ("PLONKISH_CUDA_HOST_CHUNKS", "1"), ("PLONKISH_CUDA_HOST_CUTS", "0.01,0.3,0.31,1")):
            if var:
                os.environ[var] = chunks
            try:
                assert pk.variable_base_msm(sc, reg).tobytes() == want.tobytes(), (mode, var, chunks)
            finally:
                os.environ.pop(var, None)
        reg.release()


def test_multilinear_kzg_commit_open_round_trip(pk, oracle):
    # The reference's PCS test is setup -> commit -> open -> verify (pcs/multilinear.rs:293-333)
    # with the pairing check e(C - v*G, g2) = prod e(Q_i, [s_i - x_i]_2) (kzg.rs:330-361).
    # With the trapdoor s known to the test the same identity is checked in G1:
    #   sum_i (s_i - x_i) * Q_i == C - f(x) * G.
    from plonkish_b200 import kzg

    rng = np.random.default_rng(17)
    k = 7
    ss = [int.from_bytes(rng.bytes(32), "little") % br.R for _ in range(k)]
    eq_scalars = [[1]]
    for s_i in ss:  # kzg.rs:174-194
        last = eq_scalars[-1]
        hi = [s_i * e % br.R for e in last]
        lo = [(e - h) % br.R for e, h in zip(last, hi)]
        eq_scalars.append(lo + hi)
    g = oracle.generator()
    eqs = [np.array([oracle.scalar_mul(g, e) for e in row]) for row in eq_scalars]
    for mode in (pk.G1Bases.PLAIN, pk.G1Bases.TABLE):
        pp = kzg.MultilinearKzgProverParam(eqs, mode=mode)
        evals_int = [int.from_bytes(rng.bytes(32), "little") % br.R for _ in range(1 << k)]
        evals = kzg.fr_to_montgomery(evals_int)
        comm = kzg.commit(pp, evals)
        f_s = sum(e * q for e, q in zip(evals_int, eq_scalars[k])) % br.R
        assert br.point_from_bytes(comm.tobytes()) == br.scalar_mul(f_s, br.G)  # commit(f) = f(s) * G
        assert [c.tobytes() for c in kzg.batch_commit(pp, [evals, evals[: 1 << (k - 1)]])][0] == comm.tobytes()
        assert all(c.tobytes() == comm.tobytes() for c in kzg.batch_commit(pp, [evals, evals, evals]))
        x = [int.from_bytes(rng.bytes(32), "little") % br.R for _ in range(k)]
        q_comms, value = kzg.open(pp, evals, x)
        assert len(q_comms) == k
        lhs = None
        for s_i, x_i, q in zip(ss, x, q_comms):
            lhs = br.add(lhs, br.scalar_mul((s_i - x_i) % br.R, br.point_from_bytes(q.tobytes())))
        rhs = br.add(br.point_from_bytes(comm.tobytes()), br.neg(br.scalar_mul(value, br.G)))
        assert lhs == rhs
        # each quotient commitment equals the oracle's MSM of the same quotient
        qs, _ = kzg.quotients(evals_int, x)
        for i, (q, c) in enumerate(zip(qs, q_comms)):
            assert oracle.variable_base_msm(kzg.fr_to_montgomery(q), eqs[i]).tobytes() == c.tobytes()
        pp.release()
    with pytest.raises(ValueError):
        kzg.commit(kzg.MultilinearKzgProverParam(eqs[:3], mode=pk.G1Bases.PLAIN), evals)


def test_batch_entry_matches_single_calls(pk, oracle):
    # batch_commit shape (kzg.rs:259-274): several polynomials against one SRS slice.
    n = 1 << 14
    bs = oracle.known_dlog_bases(5, 9, n)
    for mode in (pk.G1Bases.TABLE, pk.G1Bases.PLAIN):
        reg = pk.G1Bases(bs, mode=mode)
        polys = [oracle.random_scalars(n, 300 + i) for i in range(5)]
        got = pk.variable_base_msm_batch(polys, reg)
        for g, p in zip(got, polys):
            assert g.tobytes() == oracle.known_dlog_answer(5, 9, p).tobytes()
        short = [p[:1000] for p in polys[:2]]
        got = pk.variable_base_msm_batch(short, reg)
        for g, p in zip(got, short):
            assert g.tobytes() == oracle.known_dlog_answer(5, 9, p).tobytes()
        assert pk.variable_base_msm_batch([], reg).shape == (0, 8)
        reg.release()


def test_many_entry_matches_single_calls(pk, oracle):
    # open() shape (kzg.rs:291-293): MSMs of sizes 2^(k-1)..1 against the slices eqs[k-1..0].
    k = 19
    full = oracle.known_dlog_bases(5, 9, 1 << (k - 1))
    regs = [pk.G1Bases(full[: 1 << i]) for i in range(k)]
    scal = [oracle.random_scalars(1 << i, 700 + i) for i in range(k)]
    got = pk.variable_base_msm_many(scal, regs)
    for i in range(k):
        assert got[i].tobytes() == oracle.known_dlog_answer(5, 9, scal[i]).tobytes(), i
    # order, empty entries and repeated slices
    got = pk.variable_base_msm_many([scal[3], np.zeros((0, 4), np.uint64), scal[3], scal[10]], [regs[3], regs[0], regs[3], regs[10]])
    assert got[0].tobytes() == got[2].tobytes() == oracle.known_dlog_answer(5, 9, scal[3]).tobytes()
    assert not got[1].any() and got[3].tobytes() == oracle.known_dlog_answer(5, 9, scal[10]).tobytes()
    for r in regs:
        r.release()


def test_linearity(pk, oracle):
    # MSM(s, B) + MSM(t, B) == MSM(s + t, B), checked through the oracle's field/curve ops.
    n = 5000
    bs = oracle.known_dlog_bases(3, 8, n)
    s, t = oracle.random_scalars(n, 1), oracle.random_scalars(n, 2)
    st = np.array([oracle.fe_op("add", 1, a, b) for a, b in zip(s, t)])
    ps, pt, pst = (pk.variable_base_msm(x, bs) for x in (s, t, st))
    ssum = br.add(br.point_from_bytes(ps.tobytes()), br.point_from_bytes(pt.tobytes()))
    assert br.point_from_bytes(pst.tobytes()) == ssum


@pytest.mark.parametrize("log_n", [20, 22])
def test_large_sizes_against_known_discrete_log(pk, oracle, log_n):
    # BASELINE.json config 2 (2^20) and up: synthetic bases (a + i*d)G on the GPU,
    # answer from two field sums and one scalar multiplication.
    import torch

    n = 1 << log_n
    sc = pk.random_scalars(n, seed=log_n)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_bs = pk.synth_bases_device(n, 3, 5)
    got = pk.variable_base_msm_device(d_sc, d_bs).cpu().numpy().view(np.uint64)
    assert got.tobytes() == oracle.known_dlog_answer(3, 5, sc).tobytes()
    if log_n == 20:
        host_bases = d_bs.cpu().numpy().view(np.uint64)
        assert (host_bases[:: n // 64] == oracle.known_dlog_bases(3, 5, n)[:: n // 64]).all()
        assert pk.variable_base_msm(sc, host_bases).tobytes() == oracle.variable_base_msm(sc, host_bases).tobytes()


def test_full_size_2pow24_known_discrete_log(pk, oracle):
    import torch

    n = 1 << 24
    sc = pk.random_scalars(n, seed=24)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_bs = pk.synth_bases_device(n, 3, 5)
    got = pk.variable_base_msm_device(d_sc, d_bs).cpu().numpy().view(np.uint64)
    assert got.tobytes() == oracle.known_dlog_answer(3, 5, sc).tobytes()


def test_sharded_partials_sum_to_the_whole(pk, oracle):
    # Rank-style use on one GPU: per-slice projective partials, then the fold.
    import torch

    from plonkish_b200.distributed import shard_bounds

    n = 100003
    sc = pk.random_scalars(n, seed=77)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_bs = pk.synth_bases_device(n, 9, 4)
    for world in (2, 8):
        parts = []
        for r in range(world):
            b, e = shard_bounds(n, world, r)
            parts.append(pk.variable_base_msm_device(d_sc[b:e].contiguous(), d_bs[b:e].contiguous(), partial=True))
        got = pk.sum_partials_device(torch.stack(parts).contiguous()).cpu().numpy().view(np.uint64)
        assert got.tobytes() == oracle.known_dlog_answer(9, 4, sc).tobytes()


def test_host_partial_entry(pk, oracle):
    import torch

    n = 50000
    sc = oracle.random_scalars(n, 31)
    bs = oracle.known_dlog_bases(3, 5, n)
    reg = pk.G1Bases(bs)
    halves = [pk.host_partial(sc[: n // 2], reg)]
    reg2 = pk.G1Bases(bs[n // 2:])
    halves.append(pk.host_partial(sc[n // 2:], reg2))
    got = pk.sum_partials_device(torch.stack(halves).contiguous()).cpu().numpy().view(np.uint64)
    assert got.tobytes() == oracle.known_dlog_answer(3, 5, sc).tobytes()
    reg.release()
    reg2.release()


def test_multi_gpu_single_process(pk, oracle):
    from plonkish_b200 import _lib

    gpus = _lib.lib().plonkish_cuda_device_count()
    if gpus < 2:
        pytest.skip("one GPU visible")
    n = 200001
    sc = oracle.random_scalars(n, 5)
    bs = oracle.known_dlog_bases(3, 5, n)
    want = oracle.known_dlog_answer(3, 5, sc)
    for g in (2, gpus):
        assert pk.variable_base_msm(sc, bs, n_gpus=g).tobytes() == want.tobytes()
        reg = pk.ShardedG1Bases(bs, g)
        assert pk.variable_base_msm(sc, reg).tobytes() == want.tobytes()
        reg.release()


def test_concurrent_callers(pk, oracle):
    # The reference calls variable_base_msm from rayon workers (hyrax.rs:176-180).
    import threading

    n = 3000
    bs = oracle.known_dlog_bases(3, 5, n)
    inputs = [oracle.random_scalars(n, 200 + i) for i in range(8)]
    want = [oracle.known_dlog_answer(3, 5, s) for s in inputs]
    got = [None] * 8

    def work(i):
        got[i] = pk.variable_base_msm(inputs[i], bs)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(g.tobytes() == w.tobytes() for g, w in zip(got, want))
