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
