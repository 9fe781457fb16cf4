"""CPU suite: the product's kernel sources (plonkish_b200/csrc/*.cuh) compiled by
g++ against tests/emul/cuda_emul.h and run thread-for-thread on the CPU, compared
with the oracle.  This checks the kernels' logic (signed-digit recoding, the two
counting-sort levels, run partitioning, warp segmented reduction, bucket reduce)
without a GPU; the PTX carry chains themselves are pinned by the -m gpu tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, case_arrays
from oracle import bigint_ref as br

EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", EMUL_DIR], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(EMUL_DIR, "libemul_msm.so"))
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    lib.emul_msm.argtypes = [vp, vp, u32, u32, u32, u32, vp, vp, vp, vp]
    lib.emul_fq_mul.argtypes = [vp, vp, vp]
    lib.emul_fr_to_canonical.argtypes = [vp, vp]
    lib.emul_fq_mul_sum.argtypes = [vp, vp, vp, vp, vp]
    lib.emul_fq_sqr.argtypes = [vp, vp]
    lib.emul_plan.argtypes = [u32, u32, u32, vp]
    lib.emul_msm_table.argtypes = [vp, vp, u32, u32, u32, u32, vp]

    def msm_table(sc, bs, n_use, c=0, sms=148):
        sc = np.ascontiguousarray(sc, dtype=np.uint64)
        bs = np.ascontiguousarray(bs, dtype=np.uint64)
        out = np.zeros(8, dtype=np.uint64)
        lib.emul_msm_table(sc.ctypes.data, bs.ctypes.data, bs.shape[0], n_use, c, sms, out.ctypes.data)
        return out

    lib.msm_table = msm_table

    def msm(sc, bs, c=0, sms=148, serial_items=0):
        sc = np.ascontiguousarray(sc, dtype=np.uint64)
        bs = np.ascontiguousarray(bs, dtype=np.uint64)
        out = np.zeros(8, dtype=np.uint64)
        lib.emul_msm(sc.ctypes.data, bs.ctypes.data, sc.shape[0], c, sms, serial_items, out.ctypes.data, None, None, None)
        return out

    lib.msm = msm
    return lib


def _mont(vals):
    return np.array([np.frombuffer(br.scalar_to_bytes(v), dtype=np.uint64) for v in vals])


def test_field_restatement(emul):
    rng = np.random.default_rng(5)
    rinv = pow(br.MONT, -1, br.P)
    vals = [0, 1, br.P - 1, br.P - 2] + [int.from_bytes(rng.bytes(32), "little") % br.P for _ in range(40)]
    for a in vals:
        for b in vals[:10]:
            la = np.frombuffer(a.to_bytes(32, "little"), dtype=np.uint32).copy()
            lb = np.frombuffer(b.to_bytes(32, "little"), dtype=np.uint32).copy()
            out = np.zeros(8, dtype=np.uint32)
            emul.emul_fq_mul(la.ctypes.data, lb.ctypes.data, out.ctypes.data)
            assert int.from_bytes(out.tobytes(), "little") == a * b * rinv % br.P
    for v in [0, 1, br.R - 1] + [int.from_bytes(rng.bytes(32), "little") % br.R for _ in range(20)]:
        la = np.frombuffer(br.scalar_to_bytes(v), dtype=np.uint32).copy()
        out = np.zeros(8, dtype=np.uint32)
        emul.emul_fr_to_canonical(la.ctypes.data, out.ctypes.data)
        assert int.from_bytes(out.tobytes(), "little") == v


def test_fused_two_product_reduction(emul):
    # (a*b + c*d)/R mod p with one reduction: the y-coordinate form of every point formula.  Extremes
    # (all operands p-1, limbs of all ones below p) exercise the 3p bound of the running value.
    rng = np.random.default_rng(6)
    rinv = pow(br.MONT, -1, br.P)
    top = (br.P >> 224 << 224) - 1  # largest value below p whose low 7 limbs are all ones
    edge = [0, 1, br.P - 1, br.P - 2, top, (1 << 224) - 1, (1 << 253) + ((1 << 32) - 1)]
    vals = edge + [int.from_bytes(rng.bytes(32), "little") % br.P for _ in range(12)]
    limbs = lambda v: np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint32).copy()
    out = np.zeros(8, dtype=np.uint32)
    for a in vals:
        for b in edge + vals[-3:]:
            for c, d in [(a, b), (br.P - 1, br.P - 1), (top, top), (vals[-1], 0), (b, a), (vals[-2], vals[-4])]:
                la, lb, lc, ld = limbs(a), limbs(b), limbs(c), limbs(d)  # keep the arrays alive across the call
                emul.emul_fq_mul_sum(la.ctypes.data, lb.ctypes.data, lc.ctypes.data, ld.ctypes.data, out.ctypes.data)
                assert int.from_bytes(out.tobytes(), "little") == (a * b + c * d) * rinv % br.P, (a, b, c, d)


def test_symmetric_squaring(emul):
    # a*a/R with the symmetric partial products taken once (row i multiplies the doubled tail above limb i only)
    rng = np.random.default_rng(8)
    rinv = pow(br.MONT, -1, br.P)
    top = (br.P >> 224 << 224) - 1
    vals = [0, 1, 2, br.P - 1, br.P - 2, top, (1 << 224) - 1, (1 << 253) + ((1 << 32) - 1), int("7fffffff" * 7, 16), int("80000000" * 7, 16),
            int("ffffffff" * 7, 16)] + [int.from_bytes(rng.bytes(32), "little") % br.P for _ in range(500)]
    out = np.zeros(8, dtype=np.uint32)
    for v in vals:
        la = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint32).copy()
        emul.emul_fq_sqr(la.ctypes.data, out.ctypes.data)
        assert int.from_bytes(out.tobytes(), "little") == v * v * rinv % br.P, v


def test_plan_invariants(emul):
    out = np.zeros(8, dtype=np.uint32)
    for n in (1, 2, 100, 1 << 10, 1 << 16, 1 << 20, 1 << 24, 1 << 26):
        for c in (0, 8, 11, 16):
            emul.emul_plan(n, c, 148, out.ctypes.data)
            cc, w, hi, lo, idx, tile, run, threads = [int(v) for v in out]
            assert 8 <= cc <= 16 and cc * w >= 255 and hi + lo == cc - 1
            assert lo + 1 + idx <= 32 and lo <= 7 and hi <= 10 and (1 << idx) >= n
            assert tile % 8 == 0 and -(-n // tile) <= 1024
            assert threads * run >= n * w and threads % 32 == 0


def test_accumulate_runs_partition_the_entries(emul):
    """pk_acc_run (k_accumulate's tiers of decreasing run length): for any entry count up to the planned one the runs are
    consecutive, cover [0, total) exactly once, busy threads come before idle ones, whole blocks share a run length in
    the non-final tiers, and the runs shrink."""
    import ctypes

    u32 = ctypes.c_uint32
    fn = emul.emul_acc_runs
    fn.argtypes = [ctypes.c_uint64, u32, u32, u32, ctypes.c_void_p, ctypes.c_void_p, u32]
    fn.restype = u32
    rng = np.random.default_rng(11)
    cases = [(148 * 512, e) for e in (600, 5000, 98304, 1 << 20, 7_864_320, 27_262_976, 201_326_592, (1 << 26) * 16)]
    cases += [(512, e) for e in (600, 66000, 1 << 20)] + [(1024, 66000), (2048, 300000)]
    for resident, entries in cases:
        for tiers in (1, 2, 4, 7):
            cap = 1 << 23
            s = np.zeros(cap, dtype=np.uint32)
            e = np.zeros(cap, dtype=np.uint32)
            totals = [entries, entries - 1, entries // 2 + 3, entries // 6, 17, 1, 0] + [int(v) for v in rng.integers(0, entries + 1, 3)]
            for total in totals:
                nthreads = fn(entries, total, tiers, resident, s.ctypes.data, e.ctypes.data, cap)
                assert nthreads % 128 == 0 and nthreads <= cap
                ss, ee = s[:nthreads].astype(np.int64), e[:nthreads].astype(np.int64)
                busy = ss != 0xffffffff
                nb = int(busy.sum())
                assert not busy[nb:].any(), "an idle thread sits before a busy one"
                if total == 0:
                    assert nb == 0
                    continue
                assert ss[0] == 0 and ee[nb - 1] == total
                assert (ss[1:nb] == ee[:nb - 1]).all() and (ee[:nb] > ss[:nb]).all()
                lens = ee[:nb] - ss[:nb]
                assert (lens[:-1] >= lens[1:]).all() or (np.diff(lens[:-1]) <= 0).all(), "runs grow"
                assert lens.max() <= max(1024, -(-total // nthreads) + 8)


def test_golden_vectors(emul, golden):
    for case in golden["cases"]:
        sc, bs, want = case_arrays(case)
        assert emul.msm(sc, bs, 8).tobytes() == want.tobytes(), case["name"]
    sc, bs, want = case_arrays(golden["cases"][5])
    assert emul.msm(sc, bs, 16, 1).tobytes() == want.tobytes()


@pytest.mark.parametrize("n,c,sms,serial", [(1, 8, 148, 0), (257, 10, 148, 0), (3000, 12, 2, 0), (2000, 16, 1, 0), (5000, 13, 1, 16)])
def test_uniform_scalars(emul, oracle, n, c, sms, serial):
    sc = oracle.random_scalars(n, n)
    bs = oracle.known_dlog_bases(3, 5, n)
    assert (emul.msm(sc, bs, c, sms, serial) == oracle.known_dlog_answer(3, 5, sc)).all()


def test_skewed_and_degenerate_inputs(emul, oracle):
    n = 600
    rng = np.random.default_rng(0)
    bs = oracle.known_dlog_bases(9, 2, n)
    for name, vals, c, sms in [
        ("all-zero", [0] * n, 8, 148),
        ("all-minus-one", [br.R - 1] * n, 9, 1),
        ("selector", [[0, 1, br.R - 1][int(x)] for x in rng.integers(0, 3, n)], 10, 2),
        ("same-wide-value", [0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF123456789 % br.R] * n, 12, 1),
    ]:
        sc = _mont(vals)
        assert (emul.msm(sc, bs, c, sms) == oracle.variable_base_msm(sc, bs, 2)).all(), name
    dup = np.repeat(bs[:1], 300, axis=0)
    sc = oracle.random_scalars(300, 3)
    assert (emul.msm(sc, dup, 9, 1) == oracle.variable_base_msm(sc, dup, 2)).all()
    idb = bs[:300].copy()
    idb[::3] = 0
    assert (emul.msm(sc, idb, 10, 1) == oracle.variable_base_msm(sc, idb, 2)).all()


@pytest.mark.parametrize("n,use,c,sms", [(1, 1, 8, 148), (300, 77, 10, 2), (500, 500, 13, 1), (100, 100, 20, 1), (2000, 2000, 0, 1),
                                          (300, 300, 16, 2), (70, 50, 22, 1), (150, 150, 17, 1)])  # c >= 16: level 1 fused with the decomposition
def test_table_of_window_multiples(emul, oracle, n, use, c, sms):
    # mode 1: T[w][i] = 2^(c*w) * P_i built by the table kernels, one bucket set.
    sc = oracle.random_scalars(use, n + use)
    bs = oracle.known_dlog_bases(3, 5, n)
    assert (emul.msm_table(sc, bs, use, c, sms) == oracle.known_dlog_answer(3, 5, sc)).all()


def test_table_with_degenerate_inputs(emul, oracle):
    n = 300
    bs = oracle.known_dlog_bases(9, 2, n)
    bs[::7] = 0
    for vals, c in [([0] * n, 8), ([br.R - 1] * n, 9), ([0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF123456789 % br.R] * n, 12)]:
        sc = _mont(vals)
        assert (emul.msm_table(sc, bs, n, c, 1) == oracle.variable_base_msm(sc, bs, 2)).all()
    dup = np.repeat(oracle.known_dlog_bases(9, 2, 1), 200, axis=0)
    sc = oracle.random_scalars(200, 3)
    assert (emul.msm_table(sc, dup, 200, 9, 1) == oracle.variable_base_msm(sc, dup, 2)).all()
