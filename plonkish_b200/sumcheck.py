"""Host-side mirror of the reference's classic sum-check prover over the GPU round kernels.

Mirrors `ClassicSumCheck<EvaluationsProver>::prove`
(/root/reference/plonkish_backend/src/piop/sum_check/classic.rs:208-240): per round the message is
the evaluations of the round polynomial at 0..degree (classic/eval.rs:101-131), the challenge comes
from the caller's transcript (classic.rs:226-229), and every table is fixed at it
(classic.rs:90-141).  Round sums and table folds run on the GPU (csrc/sumcheck_kernels.cuh); the
O(degree^2) interpolation of `msg.evaluate(challenge)` (eval.rs:50-52, util/arithmetic.rs:108-136)
is integer arithmetic here, as it is scalar CPU work in the reference.

The expression arrives flattened — [(coeff, [poly indices]), ...] plus an optional common factor
polynomial (the eq(x, y) of a zero check) — over explicit resident tables.
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .msm import ResidentScalars, _as_u64

FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_MONT = 1 << 256
_RINV = pow(_MONT, -1, FR_MODULUS)


def _to_int(limbs) -> int:
    return int.from_bytes(np.ascontiguousarray(limbs, dtype=np.uint64).tobytes(), "little") * _RINV % FR_MODULUS


def _to_mont(v: int) -> np.ndarray:
    return np.frombuffer((v % FR_MODULUS * _MONT % FR_MODULUS).to_bytes(32, "little"), dtype=np.uint64).copy()


class SumCheckProver:
    """`ProverState` + `EvaluationsProver` (classic.rs:24-150, classic/eval.rs:68-131) for one virtual polynomial."""

    def __init__(self, polys: Sequence[ResidentScalars], terms: Sequence[Tuple[np.ndarray, Sequence[int]]], common: int = -1):
        n = polys[0].n
        assert n >= 2 and n & (n - 1) == 0 and all(p.n == n for p in polys), "tables must hold 2^num_vars evaluations"  # classic.rs:41
        self.num_vars = n.bit_length() - 1
        self.num_polys = len(polys)
        handles = np.array([p.handle for p in polys], dtype=np.uint64)
        coeffs = np.stack([_as_u64(c, 4, "coeff").reshape(4) for c, _ in terms])
        offsets = np.zeros(len(terms) + 1, dtype=np.uint32)
        flat: List[int] = []
        for t, (_, idx) in enumerate(terms):
            flat.extend(int(i) for i in idx)
            offsets[t + 1] = len(flat)
        flat_arr = np.array(flat if flat else [0], dtype=np.uint32)
        state = ctypes.c_uint64(0)
        rc = _lib.lib().plonkish_cuda_sumcheck_new(handles.ctypes.data, len(polys), self.num_vars, coeffs.ctypes.data, offsets.ctypes.data,
                                                   flat_arr.ctypes.data, len(terms), int(common), ctypes.byref(state))
        _lib.check(rc, "plonkish_cuda_sumcheck_new")
        self.state = state.value
        self.degree = _lib.lib().plonkish_cuda_sumcheck_degree(self.state)
        self.round = 0

    def round_evals(self) -> np.ndarray:
        """Evaluations of the round polynomial at X = 1..degree ([degree, 4] Montgomery limbs)."""
        out = np.zeros((self.degree, 4), dtype=np.uint64)
        _lib.check(_lib.lib().plonkish_cuda_sumcheck_round(self.state, out.ctypes.data), "plonkish_cuda_sumcheck_round")
        return out

    def fix_var(self, challenge) -> None:
        ch = _as_u64(challenge, 4, "challenge").reshape(4)
        _lib.check(_lib.lib().plonkish_cuda_sumcheck_fix_var(self.state, ch.ctypes.data), "plonkish_cuda_sumcheck_fix_var")
        self.round += 1

    def final_evals(self) -> np.ndarray:
        out = np.zeros((self.num_polys, 4), dtype=np.uint64)
        _lib.check(_lib.lib().plonkish_cuda_sumcheck_final_evals(self.state, out.ctypes.data), "plonkish_cuda_sumcheck_final_evals")
        return out

    def free(self) -> None:
        if self.state:
            _lib.check(_lib.lib().plonkish_cuda_sumcheck_free(self.state), "plonkish_cuda_sumcheck_free")
            self.state = 0


def interpolate_at(evals: Sequence[int], x: int) -> int:
    """`Evaluations::evaluate`: the polynomial through (i, evals[i]), i = 0..degree, at x
    (barycentric form of util/arithmetic.rs:108-136; any exact interpolation gives the same value)."""
    r = FR_MODULUS
    d = len(evals) - 1
    if 0 <= x <= d:
        return evals[x] % r
    total = 0
    for j, e in enumerate(evals):
        num, den = 1, 1
        for i in range(d + 1):
            if i != j:
                num = num * (x - i) % r
                den = den * (j - i) % r
        total = (total + e * num % r * pow(den, -1, r)) % r
    return total


def prove(polys: Sequence[ResidentScalars], terms, claimed_sum: int, squeeze_challenge: Callable[[List[int]], int],
          common: int = -1) -> Tuple[List[List[int]], List[int], List[int]]:
    """`ClassicSumCheck::prove` (classic.rs:208-240).  `squeeze_challenge(message)` stands for
    `msg.write(transcript)` + `transcript.squeeze_challenge()`; values are canonical integers.
    Returns (round messages, challenges, evaluations of every polynomial at the challenges)."""
    prover = SumCheckProver(polys, terms, common)
    msgs: List[List[int]] = []
    challenges: List[int] = []
    total = claimed_sum % FR_MODULUS
    try:
        for _ in range(prover.num_vars):
            tail = [_to_int(row) for row in prover.round_evals()]
            msg = [(total - tail[0]) % FR_MODULUS] + tail  # evals[0] = sum - evals[1]  (eval.rs:128)
            msgs.append(msg)
            ch = squeeze_challenge(msg) % FR_MODULUS
            challenges.append(ch)
            total = interpolate_at(msg, ch)              # state.next_round(msg.evaluate(..), ..)  (classic.rs:232)
            prover.fix_var(_to_mont(ch))
        evals = [_to_int(row) for row in prover.final_evals()]
    finally:
        prover.free()
    return msgs, challenges, evals


def prove_coefficients_to_transcript(polys: Sequence[ResidentScalars], terms, claimed_sum: int, transcript):
    """`ClassicSumCheck<CoefficientsProver>::prove` (classic.rs:208-240 with classic/coeff.rs:68-147) for the degree-2
    expressions of `additive::batch_open` (pcs/multilinear.rs:176-196: sum_j scalar_j * eq_xy(j) * poly_j): the round message
    is the COEFFICIENT vector [c0, c1, c2] of the round polynomial (coeff.rs:21-23), with c0 = sum lhs0 * rhs0,
    c2 = sum (lhs1 - lhs0)(rhs1 - rhs0) and c1 = sum - 2 c0 - c2 (coeff.rs:139-146).  The GPU round kernel returns the
    polynomial's values at X = 1, 2; the same three coefficients follow from them and the running sum:
    c0 = h(0) = sum - h(1), c2 = (h(2) - 2 h(1) + h(0)) / 2.  Returns (challenges, evals)."""
    r = FR_MODULUS
    inv2 = pow(2, -1, r)
    prover = SumCheckProver(polys, terms, -1)
    assert prover.degree == 2, "the coefficient form is implemented for products of two tables (coeff.rs:132-146)"
    challenges: List[int] = []
    total = claimed_sum % r
    try:
        for _ in range(prover.num_vars):
            h1, h2 = (_to_int(row) for row in prover.round_evals())
            c0 = (total - h1) % r
            c2 = (h2 - 2 * h1 + c0) * inv2 % r
            c1 = (total - 2 * c0 - c2) % r
            transcript.write_field_elements([c0, c1, c2])
            ch = transcript.squeeze_challenge()
            challenges.append(ch)
            total = (c0 + ch * (c1 + ch * c2)) % r       # horner(coeffs, challenge), coeff.rs:36-38
            prover.fix_var(_to_mont(ch))
        evals = [_to_int(row) for row in prover.final_evals()]
    finally:
        prover.free()
    return challenges, evals


def prove_to_transcript(polys: Sequence[ResidentScalars], terms, claimed_sum: int, transcript, common: int = -1):
    """`ClassicSumCheck::prove` with the reference's transcript calls (classic.rs:226-229): every round message goes
    down with `write_field_elements` (eval.rs:37-39), the challenge comes from `squeeze_challenge`.
    `transcript` is a plonkish_b200.transcript.Keccak256Transcript.  Returns (challenges, evals)."""

    def squeeze(msg: List[int]) -> int:
        transcript.write_field_elements(msg)
        return transcript.squeeze_challenge()

    _, challenges, evals = prove(polys, terms, claimed_sum, squeeze, common)
    return challenges, evals
