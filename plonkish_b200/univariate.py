"""Host-side mirror of the reference's univariate KZG prover (the PCS of the halo2 system, BASELINE config 4).

Mirrors ``UnivariateKzg::{commit, open, batch_open}``
(/root/reference/plonkish_backend/src/pcs/univariate/kzg.rs:242-354) over the C ABI: coefficient vectors live in HBM
(``ResidentScalars``), linear combinations (``fr_linear_combination``), divisions by ``X - z``
(``fr_div_linear``: ``div_rem``, poly/univariate.rs:144-168) and commitments (the MSM against the resident
``powers_of_s_g1``) run on the GPU; what stays here is what is scalar work in the reference too: the grouping of the
evaluations into sets (``eval_sets``, kzg.rs:454-512), the challenge powers and the set scalars (kzg.rs:514-552).

Polynomials are passed with one common length (shorter ones zero padded): a division by ``X - z`` keeps the length
and leaves a zero top coefficient, so sums of quotients of different degrees need no bookkeeping.
Values are canonical integers except where ``*_mont`` says Montgomery limbs.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .msm import G1Bases, ResidentScalars, fr_div_linear, fr_linear_combination, variable_base_msm
from .sumcheck import FR_MODULUS, _to_int, _to_mont


class GpuOps:
    """The three polynomial operations batch_open needs, on resident coefficient vectors."""

    @staticmethod
    def linear_combination(polys: Sequence[ResidentScalars], coeffs: Sequence[int]) -> ResidentScalars:
        return fr_linear_combination(polys, np.stack([_to_mont(c) for c in coeffs]))

    @staticmethod
    def div_linear(poly: ResidentScalars, z: int) -> Tuple[ResidentScalars, int]:
        q, rem = fr_div_linear(poly, _to_mont(z))
        return q, _to_int(rem)

    @staticmethod
    def commit(powers_of_s_g1: G1Bases, poly: ResidentScalars) -> np.ndarray:
        return variable_base_msm(poly, powers_of_s_g1)  # commit_coeffs, kzg.rs:24-30

    @staticmethod
    def release(poly) -> None:
        poly.release()


def commit(powers_of_s_g1: G1Bases, poly, ops=GpuOps) -> np.ndarray:
    """UnivariateKzg::commit (kzg.rs:242-252)."""
    if len(poly) > len(powers_of_s_g1):
        raise ValueError(f"Too large degree of poly to commit (param supports degree up to {len(powers_of_s_g1) - 1} but got {len(poly) - 1})")
    return ops.commit(powers_of_s_g1, poly)


def open(powers_of_s_g1: G1Bases, poly, point: int, transcript, ops=GpuOps) -> int:
    """UnivariateKzg::open (kzg.rs:264-299): quotient of the division by (X - point), its commitment written to the
    transcript.  Returns the remainder, i.e. poly(point) (what the sanity check at kzg.rs:284-290 compares with eval)."""
    if len(poly) > len(powers_of_s_g1):
        raise ValueError(f"Too large degree of poly to open (param supports degree up to {len(powers_of_s_g1) - 1} but got {len(poly) - 1})")
    quotient, remainder = ops.div_linear(poly, point % FR_MODULUS)
    transcript.write_commitment(ops.commit(powers_of_s_g1, quotient))
    ops.release(quotient)
    return remainder


class EvaluationSet:
    """kzg.rs:423-428."""

    def __init__(self, poly: int, points: List[int], diffs: List[int], evals: List[int]):
        self.polys, self.points, self.diffs, self.evals = [poly], points, diffs, [evals]


def eval_sets(evals: Sequence[Tuple[int, int, int]]) -> Tuple[List[EvaluationSet], List[int]]:
    """kzg.rs:454-512.  evals: (poly index, point index, value) in the caller's order -> (sets, sorted superset)."""
    poly_shifts: List[Tuple[int, List[int], List[int]]] = []
    superset = set()
    for poly, point, value in evals:
        for entry in poly_shifts:
            if entry[0] == poly:
                if point not in entry[1]:
                    entry[1].append(point)
                    entry[2].append(value)
                break
        else:
            poly_shifts.append((poly, [point], [value]))
        superset.add(point)
    ordered = sorted(superset)  # BTreeSet iteration order
    sets: List[EvaluationSet] = []
    for poly, points, values in poly_shifts:
        for s in sets:
            if set(s.points) == set(points):
                if poly not in s.polys:
                    s.polys.append(poly)
                    s.evals.append([values[points.index(p)] for p in s.points])
                break
        else:
            sets.append(EvaluationSet(poly, points, [i for i in ordered if i not in points], values))
    return sets, ordered


def _powers(x: int, n: int) -> List[int]:
    out, cur = [], 1
    for _ in range(n):
        out.append(cur)
        cur = cur * x % FR_MODULUS
    return out


def set_scalars(sets: Sequence[EvaluationSet], powers_of_gamma: Sequence[int], points: Sequence[int], z: int) -> Tuple[List[int], int]:
    """kzg.rs:514-533 (fflonk's normalisation by the first set)."""
    r = FR_MODULUS
    diff_evals = []
    for s in sets:
        v = 1
        for idx in s.diffs:
            v = v * (z - points[idx]) % r
        diff_evals.append(v)
    normalizer = pow(diff_evals[0], -1, r) if diff_evals[0] else 1
    return [normalizer * v % r * g % r for g, v in zip(powers_of_gamma, diff_evals)], normalizer


def vanishing_eval(points: Sequence[int], z: int) -> int:
    v = 1
    for p in points:
        v = v * (z - p) % FR_MODULUS
    return v


def batch_open(powers_of_s_g1: G1Bases, polys: Sequence, points: Sequence[int], evals: Sequence[Tuple[int, int, int]], transcript, ops=GpuOps) -> None:
    """UnivariateKzg::batch_open (kzg.rs:301-354), the non-sanity-check path: two commitments go to the transcript
    (the combined quotient q, then the opening of f at z)."""
    r = FR_MODULUS
    points = [p % r for p in points]
    sets, superset = eval_sets(evals)
    beta = transcript.squeeze_challenge()
    gamma = transcript.squeeze_challenge()
    powers_of_beta = _powers(beta, max(len(s.polys) for s in sets))
    powers_of_gamma = _powers(gamma, len(sets))
    fs, qs, owned = [], [], []
    for s in sets:
        if len(s.polys) == 1:   # powers_of_beta[0] = 1: the combination of kzg.rs:324-325 is the polynomial itself, borrowed
            f, own = polys[s.polys[0]], False
        else:
            f, own = ops.linear_combination([polys[i] for i in s.polys], powers_of_beta[: len(s.polys)]), True
        # f.div_rem(vanishing_poly): the quotient by prod (X - point) is the chain of quotients by each factor
        q = f
        for idx in s.points:
            nxt, _ = ops.div_linear(q, points[idx])
            if q is not f:
                ops.release(q)
            q = nxt
        fs.append(f)
        owned.append(own)
        qs.append(q)
    q = ops.linear_combination(qs, powers_of_gamma)                                                      # kzg.rs:330
    for t in qs:
        ops.release(t)
    q_comm = commit(powers_of_s_g1, q, ops)
    transcript.write_commitment(q_comm)                                                                   # commit_and_write, kzg.rs:332
    z = transcript.squeeze_challenge()
    normalized_scalars, normalizer = set_scalars(sets, powers_of_gamma, points, z)
    q_scalar = (-vanishing_eval([points[i] for i in superset], z) * normalizer) % r
    f = ops.linear_combination(fs + [q], normalized_scalars + [q_scalar])                                 # kzg.rs:339-343
    for t, own in zip(fs + [q], owned + [True]):
        if own:
            ops.release(t)
    open(powers_of_s_g1, f, z, transcript, ops)                                                           # kzg.rs:353
    ops.release(f)
