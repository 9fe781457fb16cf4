"""CPU suite: pins the oracle (C restatement of msm.rs:84-181) against the
independent big-int implementation, the public BN254 known answers and the
committed golden vectors.  No GPU, no product code."""
import numpy as np
import pytest

from conftest import case_arrays
from oracle import bigint_ref as br


def _pt(arr):
    return br.point_from_bytes(np.ascontiguousarray(arr, dtype=np.uint64).tobytes())


def test_montgomery_constants_match_survey():
    # SURVEY.md §8(c) constants, recomputed from p and r.
    assert br.MONT % br.P == 0x0E0A77C19A07DF2F666EA36F7879462C0A78EB28F5C70B3DD35D438DC58F0D9D
    assert br.MONT % br.R == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert (-pow(br.P, -1, 1 << 64)) % (1 << 64) == 0x87D20782E4866389
    assert (-pow(br.R, -1, 1 << 64)) % (1 << 64) == 0xC2E1F593EFFFFFFF
    assert br.R.bit_length() == 254  # util/arithmetic.rs:202-205: field_size::<Fr>() == 254


def test_public_known_answers(oracle, golden):
    g = oracle.generator()
    assert g.tobytes().hex() == golden["public_kats"]["generator"]
    assert _pt(g) == br.G
    two_g = oracle.scalar_mul(g, 2)
    assert oracle.transcript_bytes(two_g).hex() == golden["public_kats"]["two_g_canonical_be"]
    assert _pt(two_g) == br.TWO_G
    assert _pt(oracle.scalar_mul(g, br.R)) is None  # group order
    assert _pt(oracle.scalar_mul(g, br.R - 1)) == br.neg(br.G)


def test_field_ops_against_python_ints(oracle):
    rng = np.random.default_rng(1)
    for which, mod in ((0, br.P), (1, br.R)):
        vals = [0, 1, mod - 1, mod - 2, 1 << 253] + [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(50)]
        rinv = pow(br.MONT, -1, mod)
        for a in vals:
            for b in vals[:12]:
                la, lb = oracle.int_to_limbs(a), oracle.int_to_limbs(b)
                assert oracle.limbs_to_int(oracle.fe_op("mul", which, la, lb)) == a * b * rinv % mod
                assert oracle.limbs_to_int(oracle.fe_op("add", which, la, lb)) == (a + b) % mod
                assert oracle.limbs_to_int(oracle.fe_op("sub", which, la, lb)) == (a - b) % mod
            inv = oracle.limbs_to_int(oracle.fe_op("inv", which, oracle.int_to_limbs(a)))
            assert inv == (0 if a == 0 else pow(a, -1, mod) * br.MONT * br.MONT % mod)
        c = oracle.int_to_limbs(vals[7])
        assert oracle.limbs_to_int(oracle.to_canonical(which, oracle.from_canonical(which, c))[0]) == vals[7]


def test_window_helpers_match_reference_semantics(oracle):
    lib = oracle.lib()
    # msm.rs:8-14
    assert [lib.oracle_window_size(n) for n in (1, 31, 32, 1000, 1 << 16, 1 << 20, 1 << 21, 1 << 24)] == [3, 3, 3, 6, 11, 13, 14, 16]
    assert all(lib.oracle_window_size(n) == br.window_size(n) for n in (1, 7, 31, 32, 33, 100, 12345, 1 << 20, (1 << 24) // 8))
    rng = np.random.default_rng(2)
    for _ in range(200):
        rep = rng.bytes(32)
        c = int(rng.integers(1, 20))
        idx = int(rng.integers(0, -(-256 // c)))
        buf = np.frombuffer(rep, dtype=np.uint8).copy()
        got = lib.oracle_windowed_scalar(c, (1 << c) - 1, idx, buf.ctypes.data)
        assert got == br.windowed_scalar(c, (1 << c) - 1, idx, rep)
        assert got == (int.from_bytes(rep, "little") >> (idx * c)) & ((1 << c) - 1)


def test_golden_vectors(oracle, golden):
    for case in golden["cases"]:
        sc, bs, want = case_arrays(case)
        for threads in (1, 3, 8):
            got = oracle.variable_base_msm(sc, bs, threads)
            assert got.tobytes() == want.tobytes(), (case["name"], threads)
        assert oracle.msm_naive(sc, bs).tobytes() == want.tobytes(), case["name"]
        if case["result_transcript_be"] is None:
            assert not want.any()
            with pytest.raises(ValueError):
                oracle.transcript_bytes(want)
        else:
            assert oracle.transcript_bytes(want).hex() == case["result_transcript_be"], case["name"]


def test_restatement_matches_bigint_pippenger(oracle):
    # Same algorithm (msm.rs:84-181) on Python ints, including the thread chunking.
    n = 70
    sc = oracle.random_scalars(n, 3)
    bs = oracle.known_dlog_bases(11, 13, n, 2)
    ks = [br.scalar_from_bytes(s.tobytes()) for s in sc]
    pts = [_pt(b) for b in bs]
    assert all(br.is_on_curve(p) for p in pts)
    assert pts[5] == br.scalar_mul(11 + 5 * 13, br.G)
    want = br.msm(ks, pts)
    for threads in (1, 4):
        assert br.msm_pippenger_reference(ks, pts, threads) == want
        assert _pt(oracle.variable_base_msm(sc, bs, threads)) == want


@pytest.mark.parametrize("n", [1, 31, 32, 1000, 1 << 14])
def test_known_dlog_identity(oracle, n):
    sc = oracle.random_scalars(n, n)
    bs = oracle.known_dlog_bases(3, 5, n)
    assert all(oracle.is_on_curve(b) for b in bs[:: max(1, n // 16)])
    want = oracle.known_dlog_answer(3, 5, sc)
    for threads in (1, 8):
        assert (oracle.variable_base_msm(sc, bs, threads) == want).all()


def test_thread_count_does_not_change_the_affine_value(oracle):
    # msm.rs:101,157: the window size depends on the chunk length, the result must not.
    n = 3000
    sc = oracle.random_scalars(n, 9)
    bs = oracle.known_dlog_bases(2, 7, n)
    ref = oracle.variable_base_msm(sc, bs, 1)
    for threads in (2, 5, 8, 64, 4000):
        assert (oracle.variable_base_msm(sc, bs, threads) == ref).all()


def test_empty_input_returns_identity(oracle):
    # Documented deviation: the reference panics at msm.rs:154 on n == 0.
    out = oracle.variable_base_msm(np.zeros((0, 4), np.uint64), np.zeros((0, 8), np.uint64), 4)
    assert not out.any()
