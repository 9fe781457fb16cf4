"""CPU suite: the FP64-pipe field arithmetic (plonkish_b200/csrc/dpfq.cuh) compiled by g++ (fma under FE_TOWARDZERO for
__fma_rz) against Python integers: exactness of the DFMA partial-product split, the Montgomery product with radix
2^288 at the extremes of its input range, the subtraction offsets, the limb <-> word conversions, and the point
formulas against the integer-pipe ones bit for bit (including P + P, P + (-P), identity operands)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import bigint_ref as br

EMUL_DIR = os.path.join(ROOT, "tests", "emul")
P = br.P
R288 = 1 << 288
MASK48 = (1 << 48) - 1


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", EMUL_DIR], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(EMUL_DIR, "libemul_msm.so"))
    vp = ctypes.c_void_p
    lib.emul_dp_raw.argtypes = [ctypes.c_int, vp, vp, vp]
    lib.emul_dp_mul_words.argtypes = [vp, vp, vp]
    lib.emul_dp_words_roundtrip.argtypes = [vp, vp]
    lib.emul_dxyzz_madd.argtypes = [vp, vp, ctypes.c_uint32]
    lib.emul_xyzz_madd.argtypes = [vp, vp]
    lib.emul_fq_mul.argtypes = [vp, vp, vp]
    return lib


def limbs(v):
    return np.array([(v >> (48 * i)) & MASK48 for i in range(6)], dtype=np.uint64)


def value(l):
    return sum(int(x) << (48 * i) for i, x in enumerate(l))


def raw(emul, op, a, b=0):
    out = np.zeros(6, dtype=np.uint64)
    la, lb = limbs(a), limbs(b)
    emul.emul_dp_raw(op, la.ctypes.data, lb.ctypes.data, out.ctypes.data)
    assert all(int(x) < (1 << 48) for x in out), ("limb not normalised / not an integer", [hex(int(x)) for x in out])
    return value(out)


def test_product_is_exact_over_its_whole_input_range(emul):
    rng = np.random.default_rng(1)
    r288_inv = pow(R288, -1, P)
    tops = [0, 1, P - 1, P, P + 1, 2 * P, 9 * P + (1 << 240), (1 << 258) - 1, (1 << 270) - 1,
            MASK48 * sum(1 << (48 * i) for i in range(5)) + (0xffff << 240)]  # every limb at its maximum
    vals = tops + [int.from_bytes(rng.bytes(33), "little") % (10 * P) for _ in range(60)]
    for i, a in enumerate(vals):
        for b in (vals[(i * 7 + 3) % len(vals)], vals[-1 - i % 5], a):
            got = raw(emul, 0, a, b)
            assert got % P == a * b * r288_inv % P, (hex(a), hex(b))
            assert got < P + (1 << 240) + (a * b >> 288), "product out of its documented range"
        sq = raw(emul, 1, a)
        assert sq % P == a * a * r288_inv % P and sq < P + (1 << 240) + (a * a >> 288)


def test_subtraction_offsets_and_additions(emul):
    rng = np.random.default_rng(2)
    eps = 1 << 240
    for K, op in ((2, 2), (4, 3), (8, 4)):
        bmax = (K - 1) * P + eps
        for a in (0, 1, P, 9 * P + eps, int.from_bytes(rng.bytes(32), "little") % (9 * P)):
            for b in (0, 1, P - 1, P, bmax, bmax - 1, int.from_bytes(rng.bytes(32), "little") % bmax):
                assert raw(emul, op, a, b) == a - b + K * P, (K, hex(a), hex(b))
    for _ in range(50):
        a, b = (int.from_bytes(rng.bytes(33), "little") % (10 * P) for _ in range(2))
        assert raw(emul, 5, a, b) == a + b


def _words(v):
    return np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint32).copy()


def test_word_conversions_and_product_match_the_integer_pipe(emul):
    rng = np.random.default_rng(3)
    vals = [0, 1, P - 1, P - 2, (1 << 254) - 1, (1 << 256) - 1] + [int.from_bytes(rng.bytes(32), "little") for _ in range(40)]
    for v in vals:
        out = np.zeros(8, dtype=np.uint32)
        emul.emul_dp_words_roundtrip(_words(v).ctypes.data, out.ctypes.data)
        assert int.from_bytes(out.tobytes(), "little") == v
    field = [v % P for v in vals]
    for i, a in enumerate(field):
        b = field[(3 * i + 1) % len(field)]
        o1, o2 = np.zeros(8, dtype=np.uint32), np.zeros(8, dtype=np.uint32)
        aw, bw = _words(a), _words(b)
        emul.emul_dp_mul_words(aw.ctypes.data, bw.ctypes.data, o1.ctypes.data)
        emul.emul_fq_mul(aw.ctypes.data, bw.ctypes.data, o2.ctypes.data)
        assert o1.tobytes() == o2.tobytes(), (hex(a), hex(b))


def _affine_words(pt):
    return np.frombuffer(br.point_to_bytes(pt), dtype=np.uint32).copy()


def test_point_formulas_match_the_integer_pipe_bit_for_bit(emul):
    rng = np.random.default_rng(4)
    g = br.G
    pts = [br.scalar_mul(int(k), g) for k in rng.integers(1, 1 << 40, 12)]
    seqs = {
        "random": pts,
        "doubling": [pts[0], pts[0], pts[1]],                       # P + P inside the accumulator
        "cancel": [pts[0], br.neg(pts[0]), pts[2], pts[3]],         # P + (-P) -> identity, then on
        "identity-bases": [None, pts[1], None, pts[2]],
        "same-thrice": [pts[4]] * 3,
    }
    for name, seq in seqs.items():
        flat = np.concatenate([_affine_words(p) for p in seq])
        a1 = np.zeros(32, dtype=np.uint32)
        a2 = np.zeros(32, dtype=np.uint32)
        emul.emul_dxyzz_madd(a1.ctypes.data, flat.ctypes.data, len(seq))
        for k in range(len(seq)):
            emul.emul_xyzz_madd(a2.ctypes.data, flat[16 * k:].ctypes.data)
        assert a1.tobytes() == a2.tobytes(), name


def test_redundant_range_field_forms_of_the_accumulate_loop(emul):
    # fq_*_lz (csrc/fq.cuh): operands and results in [0, 2p), no subtraction after a product; checked at the range's ends
    emul.emul_fq_lazy.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5
    rng = np.random.default_rng(9)
    rinv = pow(1 << 256, -1, P)
    ends = [0, 1, P - 1, P, P + 1, 2 * P - 1, 2 * P - 2, (1 << 254), (1 << 255) - 19]
    vals = [v for v in ends if v < 2 * P] + [int.from_bytes(rng.bytes(32), "little") % (2 * P) for _ in range(40)]

    def run(op, a, b=0, c=0, d=0):
        out = np.zeros(8, dtype=np.uint32)
        arrs = [_words(v) for v in (a, b, c, d)]
        emul.emul_fq_lazy(op, *[x.ctypes.data for x in arrs], out.ctypes.data)
        return int.from_bytes(out.tobytes(), "little")

    for i, a in enumerate(vals):
        b, c, d = vals[(5 * i + 2) % len(vals)], vals[(3 * i + 7) % len(vals)], vals[-1 - i % 9]
        for x, y in ((a, b), (a, a), (2 * P - 1, 2 * P - 1), (a, 2 * P - 1)):
            got = run(0, x, y)
            assert got < 2 * P and got % P == x * y * rinv % P
        got = run(1, a)
        assert got < 2 * P and got % P == a * a * rinv % P
        got = run(2, a, b)
        assert got < 2 * P and got % P == (a + b) % P
        got = run(3, a, b)
        assert got < 2 * P and got % P == (a - b) % P
        got = run(4, a)
        assert got < 2 * P and got % P == (-a) % P
        for w, x, y, z in ((a, b, c, d), (2 * P - 1, 2 * P - 1, 2 * P - 1, 2 * P - 1), (a, 2 * P - 1, 2 * P - 1, d)):
            got = run(5, w, x, y, z)
            assert got < 2 * P and got % P == (w * x + y * z) * rinv % P


def test_lazy_point_formula_matches_the_canonical_one(emul):
    emul.emul_xyzz_madd_lazy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    rng = np.random.default_rng(10)
    pts = [br.scalar_mul(int(k), br.G) for k in rng.integers(1, 1 << 40, 40)]
    seqs = {
        "random": pts,
        "doubling": [pts[0], pts[0], pts[1]],
        "cancel": [pts[0], br.neg(pts[0]), pts[2], pts[3]],
        "identity-bases": [None, pts[1], None, pts[2]],
        "same-thrice": [pts[4]] * 3,
        "double-late": pts[:5] + [br.msm([1] * 5, pts[:5])],   # the running sum itself comes back as a base: P + P at depth
    }
    for name, seq in seqs.items():
        flat = np.concatenate([_affine_words(p) for p in seq])
        a1 = np.zeros(32, dtype=np.uint32)
        a2 = np.zeros(32, dtype=np.uint32)
        emul.emul_xyzz_madd_lazy(a1.ctypes.data, flat.ctypes.data, len(seq))
        for k in range(len(seq)):
            emul.emul_xyzz_madd(a2.ctypes.data, flat[16 * k:].ctypes.data)
        assert a1.tobytes() == a2.tobytes(), name
