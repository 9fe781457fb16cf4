"""CPU suite for the callers either side of the MSM: the oracle's
restatements of `quotients`, the g_prime merge, the eq tables and fixed_base_msm pinned
against independent Python integers, and the product's kernels (poly_kernels.cuh) run
through the CPU emulation build against the oracle.  No GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import bigint_ref as br

R = br.R
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


def _ints(limbs):
    rinv = pow(br.MONT, -1, R)
    arr = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
    return [int.from_bytes(row.tobytes(), "little") * rinv % R for row in arr]


def _fr(oracle, v):
    return oracle.from_canonical(1, oracle.int_to_limbs(v % R))[0]


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", EMUL_DIR], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(EMUL_DIR, "libemul_msm.so"))
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    lib.emul_quotients.argtypes = [vp, u32, vp, vp, vp]
    lib.emul_fr_lincomb.argtypes = [vp, vp, u32, u32, vp]
    lib.emul_eq_scalars.argtypes = [vp, u32, vp]
    lib.emul_fixed_base.argtypes = [vp, vp, u32, vp]
    lib.emul_fr_powers.argtypes = [vp, u32, vp]
    return lib


# ------------------------------------------------------------------ the oracle
def test_oracle_quotients_match_python_integers(oracle):
    # pcs/multilinear.rs:72-107 restated twice: C (oracle) and Python integers
    for k in (1, 2, 6):
        evals = oracle.random_scalars(1 << k, k)
        point = oracle.random_scalars(k, 10 + k)
        qs, value = oracle.quotients(evals, point)
        ev, pt = _ints(evals), _ints(point)
        rem = list(ev)
        want = []
        for i in reversed(range(k)):
            half = 1 << i
            want.append([(rem[half + j] - rem[j]) % R for j in range(half)])
            rem = [(rem[j] + (rem[half + j] - rem[j]) * pt[i]) % R for j in range(half)]
        want.reverse()
        assert [_ints(q) for q in qs] == want
        assert _ints(value)[0] == rem[0]
        # f(point) is the multilinear extension: sum_j evals[j] * eq_j(point)
        direct = 0
        for j, e in enumerate(ev):
            w = 1
            for i in range(k):
                w = w * (pt[i] if (j >> i) & 1 else (1 - pt[i])) % R
            direct = (direct + e * w) % R
        assert rem[0] == direct


def test_oracle_eq_scalars_and_merge_match_python_integers(oracle):
    ss = oracle.random_scalars(5, 3)
    si = _ints(ss)
    for k, e in enumerate(oracle.kzg_eq_scalars(ss)):  # kzg.rs:174-193
        assert len(e) == 1 << k
        for j, v in enumerate(_ints(e)):
            w = 1
            for i in range(k):
                w = w * (si[i] if (j >> i) & 1 else (1 - si[i])) % R
            assert v == w, (k, j)
    polys = [oracle.random_scalars(40, 20 + i) for i in range(5)]
    coeffs = oracle.random_scalars(5, 30)
    got = _ints(oracle.fr_linear_combination(polys, coeffs))  # multilinear.rs:203-213
    pi, ci = [_ints(p) for p in polys], _ints(coeffs)
    assert got == [sum(c * p[j] for c, p in zip(ci, pi)) % R for j in range(40)]


def test_oracle_fixed_base_msm_matches_bigint_scalar_mul(oracle):
    # msm.rs:16-31, 50-81 vs affine double-and-add on Python integers
    g = oracle.generator()
    vals = [0, 1, 2, 7, 8, R - 1, (1 << 253) + 12345, 0xFFFFFFFFFFFFFFFF] + [int(x) for x in np.random.default_rng(4).integers(1, 1 << 62, 6)]
    sc = np.stack([_fr(oracle, v) for v in vals])
    for window in (3, 5, 8):
        got = oracle.fixed_base_msm(g, sc, window=window)
        for v, pt in zip(vals, got):
            want = br.scalar_mul(v % R, br.G)
            assert pt.tobytes() == br.point_to_bytes(want), (window, v)


# ------------------------------------------------- the product's kernels, emulated
@pytest.mark.parametrize("k", [0, 1, 3, 10, 12])
def test_emulated_quotient_kernels(emul, oracle, k):
    # k <= 10: the single-block tail kernel only; k = 12: two fold launches, then the tail
    evals = oracle.random_scalars(1 << k, k + 1)
    point = oracle.random_scalars(max(k, 1), 7)[:k]
    q = np.zeros((1 << k, 4), dtype=np.uint64)
    value = np.zeros(4, dtype=np.uint64)
    emul.emul_quotients(evals.ctypes.data, k, point.ctypes.data if k else None, q.ctypes.data, value.ctypes.data)
    if k == 0:
        assert value.tobytes() == evals[0].tobytes()
        return
    qs, want = oracle.quotients(evals, point)
    assert value.tobytes() == want.tobytes()
    for i in range(k):
        assert q[1 << i: 2 << i].tobytes() == qs[i].tobytes(), i


def test_emulated_merge_and_eq_kernels(emul, oracle):
    for count in (1, 3, 14, 25):  # 14, 25: more terms than one launch takes (accumulate path)
        polys = [oracle.random_scalars(200, i) for i in range(count)]
        coeffs = oracle.random_scalars(count, 99)
        ptrs = (ctypes.c_void_p * count)(*[p.ctypes.data for p in polys])
        out = np.zeros((200, 4), dtype=np.uint64)
        emul.emul_fr_lincomb(ctypes.cast(ptrs, ctypes.c_void_p), coeffs.ctypes.data, count, 200, out.ctypes.data)
        assert out.tobytes() == oracle.fr_linear_combination(polys, coeffs).tobytes()
    for k in (0, 1, 5, 9):
        ss = oracle.random_scalars(max(k, 1), 3)[:k]
        out = np.zeros(((2 << k) - 1, 4), dtype=np.uint64)
        emul.emul_eq_scalars(ss.ctypes.data if k else None, k, out.ctypes.data)
        want = oracle.kzg_eq_scalars(ss) if k else [_fr(oracle, 1).reshape(1, 4)]
        for i, e in enumerate(want):
            assert out[(1 << i) - 1: (2 << i) - 1].tobytes() == np.ascontiguousarray(e).tobytes(), (k, i)


def test_emulated_powers_kernel(emul, oracle):
    # powers(s).take(n) (pcs/univariate/kzg.rs:180), chunks of 256 exponents per thread, ragged tail
    s = oracle.random_scalars(1, 77)[0]
    si = _ints(s)[0]
    for n in (1, 2, 255, 256, 257, 1000):
        out = np.zeros((n, 4), dtype=np.uint64)
        emul.emul_fr_powers(s.ctypes.data, n, out.ctypes.data)
        assert _ints(out) == [pow(si, i, R) for i in range(n)], n


def test_emulated_fixed_base_kernels(emul, oracle):
    # signed 16-bit windows: digits at and around the borrow threshold, carries through every window
    g = oracle.generator()
    sc = oracle.random_scalars(48, 4)
    edge = [0, 1, R - 1, 0x8000, 0x8001, 0x7FFF, 0xFFFF, 0x10000, (1 << 253) + 0xFFFFFFFF, int("8000" * 15, 16), int("ffff" * 15, 16),
            int("7fff" * 15, 16), (R - 1) // 2]
    for i, v in enumerate(edge):
        sc[i] = _fr(oracle, v)
    out = np.zeros((48, 8), dtype=np.uint64)
    emul.emul_fixed_base(g.ctypes.data, sc.ctypes.data, 48, out.ctypes.data)
    assert out.tobytes() == oracle.fixed_base_msm(g, sc).tobytes()


def test_emulated_division_by_a_linear_factor_matches_the_oracle(oracle):
    # poly/univariate.rs:144-168 for the divisor (X - z) (UnivariateKzg::open, pcs/univariate/kzg.rs:281-282): the chunked
    # three-pass kernels (csrc/poly_kernels.cuh k_horner_*) run on the CPU against the oracle's long division
    import ctypes
    import os
    import subprocess

    from conftest import ROOT

    emul_dir = os.path.join(ROOT, "tests", "emul")
    subprocess.run(["make", "-C", emul_dir], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(emul_dir, "libemul_msm.so"))
    lib.emul_div_linear.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    z = oracle.random_scalars(1, 11)[0]
    for n in (1, 2, 3, 63, 64, 65, 255, 256, 257, 5000, 16384, 16385, 70001):
        c = oracle.random_scalars(n, n)
        q = np.full((n, 4), 0xAB, dtype=np.uint64)
        rem = np.zeros(4, dtype=np.uint64)
        lib.emul_div_linear(c.ctypes.data, n, z.ctypes.data, q.ctypes.data, rem.ctypes.data)
        want_q, want_rem = oracle.fr_div_linear(c, z)
        assert rem.tobytes() == want_rem.tobytes(), n
        assert q[: n - 1].tobytes() == want_q.tobytes(), n
        assert not q[n - 1].any(), n
    # every chunk size the library accepts (PLONKISH_CUDA_HORNER_LOG_CHUNK), sizes around the recursion's level boundaries
    lib.emul_div_linear_chunk.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    for log_chunk in (4, 5, 6, 8):
        for n in (1, 64, 65, (64 << log_chunk) - 1, (64 << log_chunk) + 1, 40000):
            c = oracle.random_scalars(n, 1000 + n)
            q = np.full((n, 4), 0xAB, dtype=np.uint64)
            rem = np.zeros(4, dtype=np.uint64)
            lib.emul_div_linear_chunk(c.ctypes.data, n, z.ctypes.data, q.ctypes.data, rem.ctypes.data, log_chunk)
            want_q, want_rem = oracle.fr_div_linear(c, z)
            assert rem.tobytes() == want_rem.tobytes(), (log_chunk, n)
            assert q[: n - 1].tobytes() == want_q.tobytes() and not q[n - 1].any(), (log_chunk, n)


def test_univariate_batch_open_host_logic_satisfies_the_verifier(oracle):
    # pcs/univariate/kzg.rs:301-354: eval_sets, challenge powers, set scalars and the order of the transcript writes of the
    # Python mirror, with the three polynomial operations supplied by the oracle (no GPU), against batch_verify's equation
    from oracle import bigint_ref as br
    from plonkish_b200 import univariate
    from plonkish_b200.sumcheck import _to_int, _to_mont
    from plonkish_b200.transcript import Keccak256Transcript
    from univariate_verify import batch_verify_in_g1

    R = br.R

    class OracleOps:
        @staticmethod
        def linear_combination(polys, coeffs):
            return oracle.fr_linear_combination(polys, np.stack([_to_mont(c) for c in coeffs]))

        @staticmethod
        def div_linear(poly, z):
            q, rem = oracle.fr_div_linear(poly, _to_mont(z))
            return np.concatenate([q, np.zeros((1, 4), dtype=np.uint64)]), _to_int(rem)

        @staticmethod
        def commit(srs, poly):
            return oracle.variable_base_msm(poly, srs[: len(poly)])

        @staticmethod
        def release(p):
            pass

    n, s = 128, 0xC0FFEE1234567
    srs = oracle.fixed_base_msm(oracle.generator(), np.stack([_to_mont(pow(s, i, R)) for i in range(n)]))
    polys = [oracle.random_scalars(n, 40 + i) for i in range(4)]
    coeffs = [[_to_int(r) for r in p] for p in polys]

    def horner(c, x):
        acc = 0
        for v in reversed(c):
            acc = (acc * x + v) % R
        return acc

    comms = [OracleOps.commit(srs, p) for p in polys]
    points = [3, 0x55555, R - 2]
    evals = [(0, 0), (0, 1), (1, 0), (2, 2), (3, 1), (3, 0), (2, 2)]
    evals = [(p, x, horner(coeffs[p], points[x])) for p, x in evals]
    t = Keccak256Transcript()
    t.write_commitments(comms)
    univariate.batch_open(srs, polys, points, evals, t, ops=OracleOps)
    proof = t.into_proof()
    pts = [(int.from_bytes(proof[i:i + 32], "big"), int.from_bytes(proof[i + 32:i + 64], "big")) for i in range(0, len(proof), 64)]
    assert len(pts) == 6
    sets = batch_verify_in_g1(comms, points, evals, pts[4], pts[5], s)
    assert [st.polys for st in sets] == [[0, 3], [1], [2]] and sets[0].points == [0, 1] and sets[0].evals[1] == [evals[5][2], evals[4][2]]
