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

    def round_evals_factored(self) -> np.ndarray:
        """Zero check with eq(x, y) as the common factor: G(1..degree-1) of h(X) = (1 - y_r + X (2 y_r - 1)) G(X)
        ([degree - 1, 4] Montgomery limbs); see `zero_check_message`."""
        out = np.zeros((self.degree - 1, 4), dtype=np.uint64)
        _lib.check(_lib.lib().plonkish_cuda_sumcheck_round_factored(self.state, out.ctypes.data), "plonkish_cuda_sumcheck_round_factored")
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


_BARY_WEIGHTS: dict = {}


def _bary_weights(d: int) -> List[int]:
    """1 / prod_{i != j} (j - i) for the points 0..d (they depend on d alone: computed once per degree)."""
    w = _BARY_WEIGHTS.get(d)
    if w is None:
        r = FR_MODULUS
        w = []
        for j in range(d + 1):
            den = 1
            for i in range(d + 1):
                if i != j:
                    den = den * (j - i) % r
            w.append(pow(den, -1, r))
        _BARY_WEIGHTS[d] = w
    return w


def interpolate_at(evals: Sequence[int], x: int) -> int:
    """`Evaluations::evaluate`: the polynomial through (i, evals[i]), i = 0..degree, at x
    (barycentric form of util/arithmetic.rs:108-136; any exact interpolation gives the same value)."""
    r = FR_MODULUS
    d = len(evals) - 1
    if 0 <= x <= d:
        return evals[x] % r
    weights = _bary_weights(d)
    diffs = [(x - i) % r for i in range(d + 1)]
    suffix = [1] * (d + 2)  # suffix[j] = prod_{i >= j} (x - i)
    for i in range(d, -1, -1):
        suffix[i] = suffix[i + 1] * diffs[i] % r
    total, prefix = 0, 1
    for j, e in enumerate(evals):
        total += e * (prefix * suffix[j + 1] % r) % r * weights[j]
        prefix = prefix * diffs[j] % r
    return total % r


def batch_inverse(vals: Sequence[int]) -> List[int]:
    """Inverses mod r of the non-zero entries with one modular inversion (Montgomery's trick); 0 stays 0."""
    r = FR_MODULUS
    prefix, acc = [], 1
    for v in vals:
        prefix.append(acc)
        if v % r:
            acc = acc * v % r
    inv = pow(acc, -1, r)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        if vals[i] % r:
            out[i] = inv * prefix[i] % r
            inv = inv * vals[i] % r
    return out


def zero_check_message(g_tail: Sequence[int], total: int, y_r: int, inv_e0: Optional[int] = None) -> Optional[List[int]]:
    """The round message h(0..D) of a zero check from the factored round.  The common factor eq(x, y) of the round's
    pair b is (S_b (1 - y_r), S_b y_r), so h(X) = l(X) G(X) with l(X) = 1 - y_r + X (2 y_r - 1) and
    G(X) = sum_b S_b expr_b(X) of degree D - 1.  g_tail = G(1..D-1) from the GPU; G(0) follows from the running sum
    h(0) + h(1) = total, G(D) from the D points of G.  The values are the field elements the reference's
    `EvaluationsProver` sums up pair by pair (eval.rs:101-131) — the message is identical.  None when 1 - y_r is zero
    (no G(0) from the sum: the caller runs the plain round)."""
    r = FR_MODULUS
    e0, e1 = (1 - y_r) % r, y_r % r
    if e0 == 0:
        return None
    if inv_e0 is None:
        inv_e0 = pow(e0, -1, r)
    g = [(total - e1 * g_tail[0]) * inv_e0 % r] + [v % r for v in g_tail]            # G(0..D-1)
    g.append(interpolate_at(g, len(g)))                                              # G(D)
    return [(e0 + x * (e1 - e0)) % r * gx % r for x, gx in enumerate(g)]


def prove(polys: Sequence[ResidentScalars], terms, claimed_sum: int, squeeze_challenge: Callable[[List[int]], int],
          common: int = -1, zero_check_point: Optional[Sequence[int]] = None) -> Tuple[List[List[int]], List[int], List[int]]:
    """`ClassicSumCheck::prove` (classic.rs:208-240).  `squeeze_challenge(message)` stands for
    `msg.write(transcript)` + `transcript.squeeze_challenge()`; values are canonical integers.
    Returns (round messages, challenges, evaluations of every polynomial at the challenges).
    zero_check_point = y (canonical integers) states that polys[common] is eq_xy(y) (classic.rs:57-61) times a constant:
    the rounds then run factored (`zero_check_message`), the messages are the same."""
    prover = SumCheckProver(polys, terms, common)
    msgs: List[List[int]] = []
    challenges: List[int] = []
    total = claimed_sum % FR_MODULUS
    factored = zero_check_point is not None and common >= 0 and prover.degree >= 2
    if factored:
        assert len(zero_check_point) == prover.num_vars, "zero_check_point must hold one value per variable"
        inv_e0 = batch_inverse([(1 - int(v)) % FR_MODULUS for v in zero_check_point])  # y is known up front: one inversion
    try:
        for rnd in range(prover.num_vars):
            msg = None
            if factored:
                msg = zero_check_message([_to_int(row) for row in prover.round_evals_factored()], total, int(zero_check_point[rnd]), inv_e0[rnd])
            if msg is None:
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


def prove_to_transcript(polys: Sequence[ResidentScalars], terms, claimed_sum: int, transcript, common: int = -1,
                        zero_check_point: Optional[Sequence[int]] = None):
    """`ClassicSumCheck::prove` with the reference's transcript calls (classic.rs:226-229): every round message goes
    down with `write_field_elements` (eval.rs:37-39), the challenge comes from `squeeze_challenge`.
    `transcript` is a plonkish_b200.transcript.Keccak256Transcript.  Returns (challenges, evals)."""

    def squeeze(msg: List[int]) -> int:
        transcript.write_field_elements(msg)
        return transcript.squeeze_challenge()

    _, challenges, evals = prove(polys, terms, claimed_sum, squeeze, common, zero_check_point)
    return challenges, evals
