"""Host-side mirror of the reference's MultilinearKzg commit paths over the GPU MSM.

Mirrors, for BN254 (`M = Bn256`):
  MultilinearKzg::commit        /root/reference/plonkish_backend/src/pcs/multilinear/kzg.rs:252-257
  MultilinearKzg::batch_commit  kzg.rs:259-274
  MultilinearKzg::open          kzg.rs:276-302  (quotient commitments written to the transcript)
  quotients                     pcs/multilinear.rs:72-107
  UnivariateKzg::commit_coeffs  pcs/univariate/kzg.rs:24-30
  MultilinearKzg::setup (prover half)  kzg.rs:167-212   -> setup()
  g_prime merge                 pcs/multilinear.rs:203-213  -> linear_combination()
Every one of them is `variable_base_msm(scalars, srs_slice)`; only the MSM runs on the
GPU.  The scalar-field bookkeeping of `quotients` is done here with Python integers — it
is the caller's CPU work in the reference too and is meant for the parity tests and
small sizes.  The throughput path keeps polynomials resident (ResidentScalars): commit with
keep=True, merge with linear_combination(), open with open_resident(); there `quotients`
runs on the GPU as well and no scalar crosses PCIe twice.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .msm import (G1Bases, ResidentScalars, fr_linear_combination, kzg_open_resident, kzg_setup_eqs, variable_base_msm,
                  variable_base_msm_batch, variable_base_msm_batch_keep, variable_base_msm_many)

FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_MONT = 1 << 256


def fr_to_montgomery(values: Sequence[int]) -> np.ndarray:
    """Canonical integers -> [n, 4] uint64 Montgomery limbs (the bn256::Fr layout)."""
    buf = b"".join((int(v) % FR_MODULUS * _MONT % FR_MODULUS).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def fr_from_montgomery(limbs: np.ndarray) -> List[int]:
    rinv = pow(_MONT, -1, FR_MODULUS)
    arr = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
    return [int.from_bytes(row.tobytes(), "little") * rinv % FR_MODULUS for row in arr]


class MultilinearKzgProverParam:
    """`MultilinearKzgProverParams { g1, eqs }` (kzg.rs:55-77): eqs[k] holds the 2^k bases
    eq_j(s_0..s_{k-1}) * G.  Each slice is made resident on the GPU once."""

    def __init__(self, eqs: Sequence, device: int = 0, mode: int = 0):
        self.eqs = [e if isinstance(e, G1Bases) else G1Bases(e, device=device, mode=mode) for e in eqs]
        for k, e in enumerate(self.eqs):
            assert len(e) == 1 << k, f"eqs[{k}] must hold 2^{k} bases"

    def num_vars(self) -> int:
        return len(self.eqs) - 1  # kzg.rs:68-70

    def eq(self, num_vars: int) -> G1Bases:
        return self.eqs[num_vars]  # kzg.rs:74-76

    def release(self) -> None:
        for e in self.eqs:
            e.release()


def setup(g1: np.ndarray, ss: np.ndarray, device: int = 0) -> MultilinearKzgProverParam:
    """The prover half of MultilinearKzg::setup (kzg.rs:167-212) for the toxic-waste point
    ss (num_vars Montgomery Fr): eq tables, fixed-base MSM by g1, batch normalise — all on the
    GPU, every eqs[k] resident when it returns."""
    return MultilinearKzgProverParam(kzg_setup_eqs(g1, ss, device=device), device=device)


def _num_vars_of(evals) -> int:
    n = len(evals) if isinstance(evals, ResidentScalars) else np.asarray(evals).reshape(-1, 4).shape[0]
    assert n and n & (n - 1) == 0, "a multilinear polynomial has 2^k evaluations"
    return n.bit_length() - 1


def commit(pp: MultilinearKzgProverParam, evals: np.ndarray) -> np.ndarray:
    """kzg.rs:252-257: variable_base_msm(poly.evals(), pp.eq(poly.num_vars())).into()"""
    k = _num_vars_of(evals)  # evals: [2^k, 4] host array or ResidentScalars
    if k > pp.num_vars():  # validate_input, pcs/multilinear.rs:26-58
        raise ValueError(f"Too many variates of poly to commit (param supports variates up to {pp.num_vars()} but got {k})")
    return variable_base_msm(evals, pp.eq(k))


def batch_commit(pp: MultilinearKzgProverParam, polys: Sequence[np.ndarray], keep: bool = False):
    """kzg.rs:259-274: one MSM per polynomial, in order.  Polynomials of equal size go down
    as one pipelined batch (upload of the next overlaps the current MSM).  keep=True also
    returns the polynomials as ResidentScalars: (commitments, resident)."""
    polys = list(polys)
    if not polys:
        return ([], []) if keep else []
    if keep:
        k = _num_vars_of(polys[0])
        if k > pp.num_vars():
            raise ValueError(f"Too many variates of poly to batch commit (param supports variates up to {pp.num_vars()} but got {k})")
        comms, resident = variable_base_msm_batch_keep(polys, pp.eq(k))
        return list(comms), resident
    sizes = {_num_vars_of(p) for p in polys}
    if len(sizes) == 1 and len(polys) > 1:
        k = sizes.pop()
        if k > pp.num_vars():
            raise ValueError(f"Too many variates of poly to batch commit (param supports variates up to {pp.num_vars()} but got {k})")
        return list(variable_base_msm_batch(polys, pp.eq(k)))
    return [commit(pp, p) for p in polys]


def quotients(evals: Sequence[int], point: Sequence[int]) -> Tuple[List[List[int]], int]:
    """pcs/multilinear.rs:72-107 on canonical integers: for i = k-1..0 the quotient
    q_i = hi - lo (2^i values) and the remainder folds with x_i; returns ([q_0..q_{k-1}], f(x))."""
    k = len(point)
    assert len(evals) == 1 << k
    r = FR_MODULUS
    remainder = [int(v) % r for v in evals]
    qs: List[List[int]] = []
    for i in reversed(range(k)):
        half = 1 << i
        lo, hi = remainder[:half], remainder[half:2 * half]
        qs.append([(h - l) % r for h, l in zip(hi, lo)])
        remainder = [(l + (h - l) * point[i]) % r for h, l in zip(hi, lo)]
    qs.reverse()
    return qs, remainder[0]


def open(pp: MultilinearKzgProverParam, evals: np.ndarray, point: Sequence[int]) -> Tuple[List[np.ndarray], int]:
    """kzg.rs:276-302: the k quotient commitments (MSMs of 2^(k-1), ..., 2, 1 points against
    eqs[k-1..0]) the reference writes to the transcript, and the evaluation f(point)."""
    k = _num_vars_of(evals)
    if k > pp.num_vars() or len(point) != k:
        raise ValueError("Invalid point / polynomial size for open")
    qs, value = quotients(fr_from_montgomery(evals), [int(x) for x in point])
    # the quotients do not depend on their commitments: one call, small MSMs run concurrently
    comms = variable_base_msm_many([fr_to_montgomery(q) for q in qs], [pp.eq(i) for i in range(k)])
    return list(comms), value


def linear_combination(polys: Sequence[ResidentScalars], coeffs: np.ndarray) -> ResidentScalars:
    """pcs/multilinear.rs:203-213 (g_prime): sum_i coeffs[i] * polys[i], resident in, resident out."""
    return fr_linear_combination(polys, coeffs)


def open_resident(pp: MultilinearKzgProverParam, poly: ResidentScalars, point: np.ndarray) -> Tuple[List[np.ndarray], np.ndarray]:
    """kzg.rs:276-302 on a resident polynomial; point is [k, 4] Montgomery Fr.  Returns the k
    quotient commitments and f(point) as Montgomery limbs."""
    k = _num_vars_of(poly)
    pt = np.ascontiguousarray(point, dtype=np.uint64).reshape(-1, 4)
    if k > pp.num_vars() or pt.shape[0] != k:
        raise ValueError("Invalid point / polynomial size for open")
    comms, value = kzg_open_resident(poly, pp.eqs, pt)
    return list(comms), value


def open_to_transcript(pp: MultilinearKzgProverParam, poly: ResidentScalars, point: np.ndarray, transcript) -> np.ndarray:
    """`MultilinearKzg::open` as the prover runs it (kzg.rs:276-302): the quotient commitments are written to the
    transcript in order (`write_commitments`, kzg.rs:299); returns f(point) (the `remainder`, Montgomery limbs)."""
    comms, value = open_resident(pp, poly, point)
    transcript.write_commitments(comms)
    return value


def eq_xy_table(point: Sequence[int]) -> List[int]:
    """MultilinearPolynomial::eq_xy(y).evals() on canonical integers, lowest variable first (poly/multilinear.rs:91-130)."""
    r = FR_MODULUS
    evals = [1]
    for y in point:
        evals = [e * (1 - y) % r for e in evals] + [e * y % r for e in evals]
    return evals


def eq_xy_eval(x: Sequence[int], y: Sequence[int]) -> int:
    """util::arithmetic::eq_xy_eval on canonical integers: prod_i (x_i y_i + (1 - x_i)(1 - y_i))."""
    r = FR_MODULUS
    out = 1
    for a, b in zip(x, y):
        out = out * ((a * b + (1 - a) * (1 - b)) % r) % r
    return out


def batch_open(pp: MultilinearKzgProverParam, num_vars: int, polys: Sequence[ResidentScalars], points: Sequence[Sequence[int]],
               evals: Sequence[Tuple[int, int, int]], transcript, open_fn=None) -> None:
    """`additive::batch_open` (pcs/multilinear.rs:134-235), what MultilinearKzg::batch_open runs (kzg.rs:304-313), the
    non-sanity-check path.  polys: resident polynomials of 2^num_vars evaluations; points: canonical-integer points;
    evals: (poly index, point index, value) in the caller's order.  Writes the degree-2 sum check's coefficient messages
    and the quotient commitments of the final opening to the transcript; every polynomial operation runs on the GPU.
    open_fn(g_prime, challenges): another additive PCS's `open` for the last step (Zeromorph, zeromorph.rs:188-204);
    default MultilinearKzg::open."""
    from . import sumcheck
    from .msm import eq_table
    from .transcript import fr_to_montgomery

    r = FR_MODULUS
    assert all(p.n == 1 << num_vars for p in polys) and all(len(pt) == num_vars for pt in points)  # validate_input, multilinear.rs:26-58
    ell = max(len(evals) - 1, 0).bit_length()                                    # evals.len().next_power_of_two().ilog2()
    t = transcript.squeeze_challenges(ell)
    eq_xt = eq_xy_table(t)
    # merged_polys (multilinear.rs:150-167): per point, the eq_xt-weighted sum of the polynomials evaluated there; a point
    # with a single evaluation borrows the polynomial and keeps eq_xt_i as the scalar of its term
    by_point: List[List[Tuple[int, int]]] = [[] for _ in points]
    for (poly, point, _), w in zip(evals, eq_xt):
        by_point[point].append((poly, w))
    assert all(by_point), "every point needs at least one evaluation"
    merged, owned = [], []
    for entries in by_point:
        if len(entries) == 1:
            merged.append((entries[0][1], polys[entries[0][0]]))
        else:
            m = linear_combination([polys[i] for i, _ in entries], np.stack([fr_to_montgomery(w) for _, w in entries]))
            owned.append(m)
            merged.append((1, m))
    eqs = [eq_table(np.stack([fr_to_montgomery(v) for v in pt])) for pt in points]
    # expression sum_j eq_xy(j) * poly_j * scalar_j over the tables [eq_0, .., eq_{P-1}, poly_0, .., poly_{P-1}]
    tables = eqs + [m for _, m in merged]
    terms = [(fr_to_montgomery(scalar), [j, len(points) + j]) for j, (scalar, _) in enumerate(merged)]
    tilde_gs_sum = sum(v * w for (_, _, v), w in zip(evals, eq_xt)) % r       # multilinear.rs:197-198
    challenges, _ = sumcheck.prove_coefficients_to_transcript(tables, terms, tilde_gs_sum, transcript)
    # g_prime (multilinear.rs:203-213) and its opening at the sum check's point (:227-234)
    coeffs = [scalar * eq_xy_eval(challenges, pt) % r for (scalar, _), pt in zip(merged, points)]
    g_prime = linear_combination([m for _, m in merged], np.stack([fr_to_montgomery(c) for c in coeffs]))
    if open_fn is None:
        open_to_transcript(pp, g_prime, np.stack([fr_to_montgomery(c) for c in challenges]), transcript)
    else:
        open_fn(g_prime, challenges)
    for x in owned + eqs + [g_prime]:
        x.release()


def univariate_setup(g1: np.ndarray, s: np.ndarray, poly_size: int, device: int = 0) -> G1Bases:
    """The G1 half of UnivariateKzg::setup (pcs/univariate/kzg.rs:175-195): powers_of_s_g1, built and kept on the GPU."""
    from .msm import kzg_setup_powers

    return kzg_setup_powers(g1, s, poly_size, device=device)


def commit_coeffs(powers_of_s_g1: G1Bases, coeffs: np.ndarray) -> np.ndarray:
    """pcs/univariate/kzg.rs:24-30: variable_base_msm(coeffs, &powers_of_s_g1[..coeffs.len()])."""
    return variable_base_msm(coeffs, powers_of_s_g1)
