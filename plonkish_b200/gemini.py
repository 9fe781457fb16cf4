"""Host-side mirror of the reference's Gemini PCS over the univariate KZG SRS — the third BN254 multilinear caller of the MSM.

Mirrors ``Gemini<UnivariateKzg<Bn256>>::{commit, batch_commit, open, batch_open}``
(/root/reference/plonkish_backend/src/pcs/multilinear/gemini.rs:58-160) over the C ABI.  ``commit`` is ``commit_coeffs`` of
the evaluations (univariate/kzg.rs:24-30).  ``open`` folds the polynomial num_vars - 1 times (``merge_into``,
poly/multilinear.rs:599-618), commits the folds against prefixes of ``powers_of_s_g1``, evaluates them at beta and at the
negated squares of beta and hands everything to ``UnivariateKzg::batch_open`` (univariate.py).  On the GPU the folds are
one packed resident vector (``fr_gemini_folds``) committed in one call (``variable_base_msm_many_resident``); each fold
is then a slice handle (``scalars_slice``) for the divisions and the mixed-length sums of batch_open
(``fr_linear_combination_padded``).  Values are canonical integers.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import univariate
from .msm import (G1Bases, ResidentScalars, fr_div_linear, fr_gemini_folds, fr_linear_combination_padded, scalars_slice,
                  variable_base_msm_many_resident)
from .sumcheck import FR_MODULUS, _to_int, _to_mont


class GpuOps(univariate.GpuOps):
    """The polynomial operations of Gemini::open on resident vectors (plus UnivariateKzg::batch_open's, inherited; its sums
    run over polynomials of different lengths here)."""

    @staticmethod
    def linear_combination(polys: Sequence[ResidentScalars], coeffs: Sequence[int]) -> ResidentScalars:
        return fr_linear_combination_padded(polys, np.stack([_to_mont(c) for c in coeffs]))

    @staticmethod
    def folds(poly: ResidentScalars, point: Sequence[int]) -> ResidentScalars:
        return fr_gemini_folds(poly, np.stack([_to_mont(x) for x in point]))

    @staticmethod
    def fold_views(folds: ResidentScalars, num_vars: int) -> List[ResidentScalars]:
        return [scalars_slice(folds, 1 << (num_vars - i), 1 << (num_vars - i)) for i in range(1, num_vars)]

    @staticmethod
    def commit_folds(powers_of_s_g1: G1Bases, folds: ResidentScalars, num_vars: int) -> np.ndarray:
        sizes = [1 << (num_vars - i) for i in range(1, num_vars)]  # f_i: 2^(n-i) scalars at offset 2^(n-i)
        from .zeromorph import prefix_bases

        return variable_base_msm_many_resident(folds, sizes, [prefix_bases(powers_of_s_g1, m) for m in sizes], sizes)

    # evaluate() leaves its quotient here: batch_open starts the set of that polynomial with the same division
    # (univariate.batch_open divides a single-polynomial set's polynomial itself), which then costs nothing
    _quotients: dict = {}

    @staticmethod
    def evaluate(poly: ResidentScalars, x: int) -> int:
        q, rem = fr_div_linear(poly, _to_mont(x))  # the remainder of the division by (X - x) is poly(x)
        old = GpuOps._quotients.pop((poly.handle, x), None)
        if old is not None:
            old[0].release()
        GpuOps._quotients[(poly.handle, x)] = (q, _to_int(rem))
        return _to_int(rem)

    @staticmethod
    def div_linear(poly: ResidentScalars, z: int):
        hit = GpuOps._quotients.pop((poly.handle, z), None)
        return hit if hit is not None else univariate.GpuOps.div_linear(poly, z)

    @staticmethod
    def drop_quotients() -> None:
        """Release what evaluate() left and no division took over."""
        for q, _ in GpuOps._quotients.values():
            q.release()
        GpuOps._quotients.clear()


class GeminiKzgProverParam:
    """Gemini's ProverParam is UnivariateKzg's (gemini.rs:38): powers_of_s_g1 (univariate/kzg.rs:41-55)."""

    def __init__(self, powers_of_s_g1: G1Bases, prefix_tables: bool = False):
        self.powers_of_s_g1 = powers_of_s_g1
        if prefix_tables:  # own resident slices for the fold commitments' prefixes (zeromorph.build_prefix_tables)
            from .zeromorph import build_prefix_tables

            powers_of_s_g1.prefix_tables = build_prefix_tables(powers_of_s_g1, len(powers_of_s_g1))

    def degree(self) -> int:
        return len(self.powers_of_s_g1) - 1

    def release(self) -> None:
        from .zeromorph import release_prefix_tables

        release_prefix_tables(self.powers_of_s_g1)
        self.powers_of_s_g1.release()


def commit(pp: GeminiKzgProverParam, poly, ops=GpuOps) -> np.ndarray:
    """Gemini::commit (gemini.rs:58-68)."""
    if pp.degree() + 1 < len(poly):
        raise ValueError(f"Too large degree of poly to commit (param supports degree up to {pp.degree()} but got {len(poly)})")
    return ops.commit(pp.powers_of_s_g1, poly)


def batch_commit(pp: GeminiKzgProverParam, polys: Sequence, keep: bool = False, ops=GpuOps):
    """gemini.rs:70-78; keep=True: the pipelined batch entry over host polynomials of one size, which leaves them
    resident — returns (commitments, [ResidentScalars])."""
    polys = list(polys)
    if keep:
        from .msm import variable_base_msm_batch_keep

        if not polys:
            return [], []
        if pp.degree() + 1 < len(polys[0]):
            raise ValueError(f"Too large degree of poly to commit (param supports degree up to {pp.degree()} but got {len(polys[0])})")
        comms, resident = variable_base_msm_batch_keep(polys, pp.powers_of_s_g1)
        return list(comms), resident
    return [commit(pp, poly, ops) for poly in polys]


def open_points(beta: int, num_vars: int) -> List[int]:
    """chain![[beta], squares(beta).map(Neg::neg)].take(num_vars + 1) (gemini.rs:131-133)."""
    r = FR_MODULUS
    points, sq = [beta % r], beta % r
    for _ in range(num_vars):
        points.append((-sq) % r)
        sq = sq * sq % r
    return points


def open_queries(num_vars: int) -> List[Tuple[int, int]]:
    """chain!([(0, 0), (0, 1)], (1..num_vars).zip(2..)) (gemini.rs:135)."""
    return [(0, 0), (0, 1)] + [(i, i + 1) for i in range(1, num_vars)]


def open(pp: GeminiKzgProverParam, poly, point: Sequence[int], transcript, ops=GpuOps) -> None:
    """Gemini::open (gemini.rs:80-141), the non-sanity-check path (commitment and evaluation are only read by the sanity
    checks).  Writes num_vars - 1 fold commitments, num_vars evaluations and the two points of UnivariateKzg::batch_open."""
    r = FR_MODULUS
    num_vars = len(point)
    assert num_vars >= 1, "the reference indexes point[..num_vars - 1]"
    if pp.degree() + 1 < len(poly):
        raise ValueError(f"Too large degree of poly to open (param supports degree up to {pp.degree()} but got {len(poly)})")
    point = [int(x) % r for x in point]
    folds = ops.folds(poly, point)                                                        # :98-108
    views = []
    try:
        views = ops.fold_views(folds, num_vars)
        fs = [poly] + views
        transcript.write_commitments(ops.commit_folds(pp.powers_of_s_g1, folds, num_vars))   # batch_commit_and_write(&fs[1..]), :124-128
        beta = transcript.squeeze_challenge()
        points = open_points(beta, num_vars)
        evals = [(idx, pt, ops.evaluate(fs[idx], points[pt])) for idx, pt in open_queries(num_vars)]   # :135-137
        transcript.write_field_elements([v for _, _, v in evals[1:]])                        # :138
        univariate.batch_open(pp.powers_of_s_g1, fs, points, evals, transcript, ops)          # :140
    finally:
        if hasattr(ops, "drop_quotients"):
            ops.drop_quotients()
        for v in views:
            ops.release(v)
        ops.release(folds)


def batch_open(pp: GeminiKzgProverParam, num_vars: int, polys: Sequence, points: Sequence[Sequence[int]], evals: Sequence[Tuple[int, int, int]],
               transcript) -> None:
    """Gemini::batch_open (gemini.rs:143-158) = additive::batch_open (pcs/multilinear.rs:134-235) with this PCS's open."""
    from . import kzg

    kzg.batch_open(None, num_vars, polys, points, evals, transcript, open_fn=lambda g_prime, challenges: open(pp, g_prime, challenges, transcript))
