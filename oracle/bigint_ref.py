"""oracle/bigint_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Independent big-integer restatement of BN254 G1 arithmetic (affine coordinates,
Python ints) used to pin the C oracle (oracle/bn254_oracle.c) and to generate
the committed golden vectors under tests/golden/.  It shares no code with the C
oracle or with the CUDA path: textbook affine chord-and-tangent formulas,
`pow(x, -1, p)` inversions, double-and-add.

Reference anchors (paths relative to /root/reference/plonkish_backend/src):
  * the value computed is the one `variable_base_msm` returns after the callers'
    `.into()` / `.to_affine()` (pcs/multilinear/kzg.rs:255,271,292;
    pcs/univariate/kzg.rs:28; pcs.rs:175);
  * byte encodings follow halo2_curves 0.3.3 [ext, un-vendored]: Fr/Fq are
    4x64-bit little-endian Montgomery limbs with R = 2^256, G1Affine is x||y with
    (0, 0) for the identity;
  * transcript encoding follows util/transcript.rs:216-229.
Parity status: "parity unpinned" against reference binaries (no Rust toolchain,
no golden vectors in the reference); pinned by public BN254/EIP-196 constants.
"""
from __future__ import annotations

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq modulus
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr modulus (group order)
MONT = 1 << 256
B = 3
G = (1, 2)
IDENTITY = None  # affine point at infinity

# EIP-196 / alt_bn128 public known answer: 2*G.
TWO_G = (
    0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3,
    0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4,
)


def is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B) % P == 0


def neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def scalar_mul(k: int, pt):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc


def msm(scalars, points):
    """sum_i scalars[i] * points[i] with canonical integer scalars."""
    assert len(scalars) == len(points)  # msm.rs:90
    acc = None
    for k, pt in zip(scalars, points):
        acc = add(acc, scalar_mul(k, pt))
    return acc


# ---------------------------------------------------------------- encodings

def fe_to_mont_bytes(v: int, modulus: int) -> bytes:
    return ((v % modulus) * MONT % modulus).to_bytes(32, "little")


def fe_from_mont_bytes(b: bytes, modulus: int) -> int:
    v = int.from_bytes(b, "little")
    assert v < modulus, "non-canonical Montgomery representation"
    return v * pow(MONT, -1, modulus) % modulus


def scalar_to_bytes(k: int) -> bytes:
    return fe_to_mont_bytes(k, R)


def scalar_from_bytes(b: bytes) -> int:
    return fe_from_mont_bytes(b, R)


def point_to_bytes(pt) -> bytes:
    if pt is None:
        return bytes(64)
    return fe_to_mont_bytes(pt[0], P) + fe_to_mont_bytes(pt[1], P)


def point_from_bytes(b: bytes):
    assert len(b) == 64
    if b == bytes(64):
        return None
    return (fe_from_mont_bytes(b[:32], P), fe_from_mont_bytes(b[32:], P))


def transcript_bytes(pt) -> bytes:
    """util/transcript.rs:216-229: x then y, 32-byte big-endian canonical."""
    if pt is None:
        raise ValueError("identity has no coordinates (reference unwrap() panics)")
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


# ------------------------------------------------- reference window helpers

def window_size(num_scalars: int) -> int:
    """util/arithmetic/msm.rs:8-14."""
    import math

    if num_scalars < 32:
        return 3
    return int(math.floor(math.log(float(num_scalars))))


def windowed_scalar(window_size_: int, window_mask: int, idx: int, repr_le: bytes) -> int:
    """util/arithmetic/msm.rs:33-48."""
    skip_bits = idx * window_size_
    skip_bytes = skip_bits // 8
    value = bytearray(8)
    for k, src in zip(range(8), repr_le[skip_bytes:]):
        value[k] = src
    return (int.from_bytes(value, "little") >> (skip_bits - skip_bytes * 8)) & window_mask


def msm_pippenger_reference(scalars, points, num_threads: int = 1):
    """Restates util/arithmetic/msm.rs:84-181 on Python ints (small n only)."""
    n = len(scalars)
    assert n == len(points)
    if n == 0:
        return None

    def serial(sc, pts):
        reprs = [(k % R).to_bytes(32, "little") for k in sc]
        c = window_size(len(sc))
        num_buckets = (1 << c) - 1
        num_windows = -(-256 // c)
        result = None
        for idx in reversed(range(num_windows)):
            for _ in range(c):
                result = add(result, result)
            buckets = [None] * num_buckets
            for rep, pt in zip(reprs, pts):
                d = windowed_scalar(c, num_buckets, idx, rep)
                if d != 0:
                    buckets[d - 1] = add(buckets[d - 1], pt)
            running = None
            for b in reversed(buckets):
                running = add(b, running)
                result = add(result, running)
        return result

    if n <= num_threads:
        return serial(scalars, points)
    chunk = -(-n // num_threads)
    acc = None
    for s in range(0, n, chunk):
        acc = add(acc, serial(scalars[s:s + chunk], points[s:s + chunk]))
    return acc
