"""Host-side mirror of the reference's Keccak256 Fiat-Shamir transcript.

Mirrors `FiatShamirTranscript<Keccak256, Cursor<Vec<u8>>>` (= `Keccak256Transcript`,
/root/reference/plonkish_backend/src/util/transcript.rs:100-235) for `bn256::G1Affine` / `bn256::Fr`:

  common_field_element   state.update(fe.to_repr())           32 bytes little-endian canonical   (:132-135, util/hash.rs:19-21)
  write_field_element    common + stream.write(repr reversed) 32 bytes big-endian                (:158-166)
  common_commitment      update x.to_repr(), y.to_repr()      identity is an error               (:172-184)
  write_commitment       common + stream x, y big-endian                                        (:216-229)
  squeeze_challenge      hash = finalize_fixed_reset(); state.update(hash); fe_mod_from_le_bytes(hash)   (:126-131)

Keccak256 is the original Keccak padding (0x01 ... 0x80, rate 136), which Python's hashlib does not offer
(sha3_256 pads with 0x06), so Keccak-f[1600] is written out here; `sha3_256` below exists to pin the permutation and
the sponge against hashlib in the tests.  Pure Python: the transcript sees a few kilobytes per proof.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np

FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_MONT = 1 << 256
_MASK = (1 << 64) - 1

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
    0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
    0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]  # [x][y]


def _rol(v: int, n: int) -> int:
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _MASK if n else v


def keccak_f1600(a: List[int]) -> List[int]:
    """The 24-round permutation on 25 lanes, lane (x, y) at index x + 5 y."""
    for rc in _RC:
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], _ROT[x][y])
        a = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= rc
    return a


def _permute(state: List[int]) -> List[int]:
    """Keccak-f through the library's host routine (plonkish_cuda_keccak_f1600, same permutation in C: a pure-Python
    round takes 0.6 ms) when the shared object is there, else the Python rounds above; the tests compare the two."""
    try:
        from . import _lib

        lib = _lib.load()
    except Exception:  # noqa: BLE001  (library not built: the transcript still works)
        return keccak_f1600(state)
    arr = np.array(state, dtype=np.uint64)
    lib.plonkish_cuda_keccak_f1600(arr.ctypes.data)
    return [int(v) for v in arr]


class _Sponge:
    RATE = 136  # 1600 - 2 * 256 bits

    def __init__(self, domain: int):
        self.domain = domain
        self.state = [0] * 25
        self.buf = b""

    def _absorb_block(self, block: bytes) -> None:
        for i in range(self.RATE // 8):
            self.state[i] ^= int.from_bytes(block[8 * i: 8 * i + 8], "little")
        self.state = _permute(self.state)

    def update(self, data: bytes) -> None:
        self.buf += bytes(data)
        while len(self.buf) >= self.RATE:
            self._absorb_block(self.buf[: self.RATE])
            self.buf = self.buf[self.RATE:]

    def finalize_reset(self) -> bytes:
        pad = bytearray(self.RATE - len(self.buf))
        pad[0] ^= self.domain
        pad[-1] ^= 0x80
        self._absorb_block(self.buf + bytes(pad))
        out = b"".join(lane.to_bytes(8, "little") for lane in self.state[:4])
        self.state = [0] * 25
        self.buf = b""
        return out


def keccak256(data: bytes) -> bytes:
    s = _Sponge(0x01)
    s.update(data)
    return s.finalize_reset()


def sha3_256(data: bytes) -> bytes:
    """Same sponge with the SHA-3 domain byte; only here so the tests can compare with hashlib."""
    s = _Sponge(0x06)
    s.update(data)
    return s.finalize_reset()


def _fe_canonical(limbs, modulus: int) -> int:
    v = int.from_bytes(np.ascontiguousarray(limbs, dtype=np.uint64).tobytes(), "little")
    return v * pow(_MONT, -1, modulus) % modulus


def fr_to_montgomery(v: int) -> np.ndarray:
    return np.frombuffer((v % FR_MODULUS * _MONT % FR_MODULUS).to_bytes(32, "little"), dtype=np.uint64).copy()


class Keccak256Transcript:
    """The prover side of `Keccak256Transcript` (write + squeeze) over Montgomery-limb arrays as they cross the C ABI."""

    def __init__(self):
        self.state = _Sponge(0x01)
        self.stream = bytearray()

    # -- field elements (canonical integers) ------------------------------------------------------------------
    def common_field_element(self, value: int) -> None:
        self.state.update((value % FR_MODULUS).to_bytes(32, "little"))

    def write_field_element(self, value: int) -> None:
        self.common_field_element(value)
        self.stream += (value % FR_MODULUS).to_bytes(32, "big")

    def write_field_elements(self, values: Iterable[int]) -> None:
        for v in values:
            self.write_field_element(v)

    # -- commitments ([8] uint64 Montgomery limbs, x || y) -------------------------------------------------------
    def write_commitment(self, point) -> None:
        p = np.ascontiguousarray(point, dtype=np.uint64).reshape(8)
        if not p.any():
            raise ValueError("Invalid elliptic curve point encoding")  # coordinates() of the identity (transcript.rs:175-181)
        x, y = _fe_canonical(p[:4], FQ_MODULUS), _fe_canonical(p[4:], FQ_MODULUS)
        self.state.update(x.to_bytes(32, "little"))
        self.state.update(y.to_bytes(32, "little"))
        self.stream += x.to_bytes(32, "big") + y.to_bytes(32, "big")

    def write_commitments(self, points: Sequence) -> None:
        for p in points:
            self.write_commitment(p)

    # -- challenges -------------------------------------------------------------------------------------------------
    def squeeze_challenge(self) -> int:
        h = self.state.finalize_reset()
        self.state.update(h)
        return int.from_bytes(h, "little") % FR_MODULUS  # fe_mod_from_le_bytes (util/arithmetic.rs:150-152)

    def squeeze_challenges(self, n: int) -> List[int]:
        return [self.squeeze_challenge() for _ in range(n)]

    def into_proof(self) -> bytes:
        return bytes(self.stream)
