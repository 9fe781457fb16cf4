"""Host-side mirror of the reference's Zeromorph PCS over the univariate KZG SRS — another multilinear caller of the MSM.

Mirrors ``Zeromorph<UnivariateKzg<Bn256>>::{trim, commit, batch_commit, open, batch_open}``
(/root/reference/plonkish_backend/src/pcs/multilinear/zeromorph.rs:84-206) over the C ABI.  A multilinear polynomial's
2^n evaluations are committed as the coefficients of a univariate polynomial (``commit_coeffs``, univariate/kzg.rs:24-30:
``variable_base_msm(evals, &powers_of_s_g1[..2^n])``); ``open`` commits the n quotients of ``quotients``
(pcs/multilinear.rs:72-107) against prefixes of the same SRS, folds them into ``q_hat`` and ``f`` and opens ``f`` with
``UnivariateKzg::open``.  Everything that touches 2^n scalars runs on the GPU on resident vectors (``GpuOps``):
quotients (``fr_quotients``), their n commitments in one call (``variable_base_msm_many_resident``), ``q_hat`` and ``f``
(``zeromorph_q_hat`` / ``zeromorph_f``), the division by ``X - x`` and the last MSM; what stays here is the reference's
own scalar work — challenge powers and ``eval_and_quotient_scalars`` (zeromorph.rs:259-294) — and the transcript.

Values are canonical integers; polynomials are ``ResidentScalars`` (or whatever the ``ops`` class works on: the CPU
tests drive the same host logic through the oracle).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import univariate
from .msm import G1Bases, ResidentScalars, fr_quotients, variable_base_msm_many_resident, zeromorph_f, zeromorph_q_hat
from .sumcheck import FR_MODULUS, _to_int, _to_mont


def _mont_rows(values: Sequence[int]) -> np.ndarray:
    return np.stack([_to_mont(v) for v in values]) if len(values) else np.zeros((0, 4), dtype=np.uint64)


def build_prefix_tables(powers_of_s_g1: G1Bases, poly_size: int) -> dict:
    """Optional: one resident slice of its own (with its own table of window multiples and window size) for every proper
    power-of-two prefix of powers_of_s_g1 — the bases of the quotient / fold commitments (`&powers_of_s_g1[..2^i]`,
    univariate/kzg.rs:28).  Without them a prefix inherits the full slice's window size (and pays its bucket set) down to
    1/16 of the slice and falls back to the plain layout below; with them the chain of halving MSMs runs like
    MultilinearKzg::open's over eqs[i].  Costs about as much HBM again as the full slice's table (14 GB at 2^24) and one
    pass of the points through host memory.  Attach the result as `powers_of_s_g1.prefix_tables`."""
    tables, size = {}, 1
    while size < poly_size:
        tables[size] = G1Bases(powers_of_s_g1.to_host(0, size), device=powers_of_s_g1.device)
        size <<= 1
    return tables


def prefix_bases(powers_of_s_g1: G1Bases, size: int) -> G1Bases:
    return getattr(powers_of_s_g1, "prefix_tables", {}).get(size, powers_of_s_g1)


def release_prefix_tables(powers_of_s_g1: G1Bases) -> None:
    for b in getattr(powers_of_s_g1, "prefix_tables", {}).values():
        b.release()
    powers_of_s_g1.prefix_tables = {}


class GpuOps(univariate.GpuOps):
    """The polynomial operations of Zeromorph::open on resident vectors (plus UnivariateKzg::open's, inherited)."""

    @staticmethod
    def quotients(poly: ResidentScalars, point: Sequence[int]) -> Tuple[ResidentScalars, int]:
        q, value = fr_quotients(poly, _mont_rows(point))
        return q, _to_int(value)

    @staticmethod
    def commit_quotients(powers_of_s_g1: G1Bases, q: ResidentScalars, num_vars: int) -> np.ndarray:
        sizes = [1 << i for i in range(num_vars)]  # q_i: 2^i scalars at offset 2^i, against powers_of_s_g1[..2^i]
        return variable_base_msm_many_resident(q, sizes, [prefix_bases(powers_of_s_g1, m) for m in sizes], sizes)

    @staticmethod
    def q_hat(q: ResidentScalars, weights: Sequence[int]) -> ResidentScalars:
        return zeromorph_q_hat(q, _mont_rows(weights))

    @staticmethod
    def f(poly: ResidentScalars, q_hat: ResidentScalars, q: ResidentScalars, z: int, c0: int, q_scalars: Sequence[int]) -> ResidentScalars:
        return zeromorph_f(poly, q_hat, q, _to_mont(z), _to_mont(c0), _mont_rows(q_scalars))


class ZeromorphKzgProverParam:
    """``ZeromorphKzgProverParam { commit_pp, open_pp }`` (zeromorph.rs:30-39): two views of powers_of_s_g1."""

    def __init__(self, commit_pp: G1Bases, open_pp: G1Bases):
        self.commit_pp, self.open_pp = commit_pp, open_pp

    def degree(self) -> int:
        return len(self.commit_pp) - 1  # zeromorph.rs:36-38

    def release(self) -> None:
        release_prefix_tables(self.commit_pp)
        if self.open_pp is not self.commit_pp:
            self.open_pp.release()
        self.commit_pp.release()


def trim(powers_of_s_g1: G1Bases, poly_size: int, prefix_tables: bool = False) -> ZeromorphKzgProverParam:
    """Zeromorph::trim (zeromorph.rs:84-102), prover half: commit_pp = the first poly_size powers (UnivariateKzg::trim,
    univariate/kzg.rs:214-233), open_pp = the LAST poly_size powers (offset = len - poly_size).  With a setup of exactly
    poly_size powers the two coincide and share one resident slice; otherwise open_pp becomes a slice of its own (the
    points pass through host memory once, at trim time).  prefix_tables: build_prefix_tables for the quotient commitments."""
    total = len(powers_of_s_g1)
    if poly_size > total:
        raise ValueError(f"Too large poly_size to trim to (param supports poly_size up to {total} but got {poly_size})")
    offset = total - poly_size
    if offset == 0:
        pp = ZeromorphKzgProverParam(powers_of_s_g1, powers_of_s_g1)
    else:
        commit_pp = G1Bases(powers_of_s_g1.to_host(0, poly_size), device=powers_of_s_g1.device)
        open_pp = G1Bases(powers_of_s_g1.to_host(offset, poly_size), device=powers_of_s_g1.device)
        pp = ZeromorphKzgProverParam(commit_pp, open_pp)
    if prefix_tables:  # see build_prefix_tables
        pp.commit_pp.prefix_tables = build_prefix_tables(pp.commit_pp, poly_size)
    return pp


def commit(pp: ZeromorphKzgProverParam, poly, ops=GpuOps) -> np.ndarray:
    """Zeromorph::commit (zeromorph.rs:104-114): commit_coeffs of the evaluations."""
    if pp.degree() + 1 < len(poly):
        raise ValueError(f"Too large degree of poly to commit (param supports degree up to {pp.degree()} but got {len(poly)})")
    return ops.commit(pp.commit_pp, poly)


def batch_commit(pp: ZeromorphKzgProverParam, polys: Sequence, keep: bool = False, ops=GpuOps):
    """zeromorph.rs:116-124: one commit per polynomial, in order.  keep=True (host polynomials of one size): the pipelined
    batch entry, which also leaves every polynomial resident — returns (commitments, [ResidentScalars])."""
    polys = list(polys)
    if keep:
        from .msm import variable_base_msm_batch_keep

        if not polys:
            return [], []
        if pp.degree() + 1 < len(polys[0]):
            raise ValueError(f"Too large degree of poly to commit (param supports degree up to {pp.degree()} but got {len(polys[0])})")
        comms, resident = variable_base_msm_batch_keep(polys, pp.commit_pp)
        return list(comms), resident
    return [commit(pp, poly, ops) for poly in polys]


def _powers(x: int, n: int) -> List[int]:
    return univariate._powers(x, n)


def eval_and_quotient_scalars(y: int, x: int, z: int, u: Sequence[int]) -> Tuple[int, List[int]]:
    """zeromorph.rs:259-294 on canonical integers.  With v_i = (x^(2^n) - 1) / (x^(2^i) - 1) and offset_i = x^(2^n - 2^i):
    eval_scalar = -v_0 z and q_scalars[i] = -(y^i offset_i + z (x^(2^i) v_(i+1) - u_i v_i))."""
    r = FR_MODULUS
    num_vars = len(u)
    squares_of_x = [x % r]
    for _ in range(num_vars):
        squares_of_x.append(squares_of_x[-1] * squares_of_x[-1] % r)       # squares(x).take(num_vars + 1)
    offsets_of_x, state = [], 1
    for power_of_x in reversed(squares_of_x[:-1]):                         # .rev().skip(1).scan(ONE, ..)
        state = state * power_of_x % r
        offsets_of_x.append(state)
    offsets_of_x.reverse()
    v_numer = (squares_of_x[num_vars] - 1) % r
    # BatchInvert leaves a zero where the denominator is zero (x a 2^i-th root of unity: negligible for a challenge)
    vs = [v_numer * (pow(sq - 1, -1, r) if (sq - 1) % r else 0) % r for sq in squares_of_x]
    q_scalars = [(-(power_of_y * offset_of_x + z * (square_of_x * v_j - u_i * v_i))) % r
                 for power_of_y, offset_of_x, square_of_x, v_i, v_j, u_i in zip(_powers(y, num_vars), offsets_of_x, squares_of_x, vs, vs[1:], u)]
    return (-vs[0] * z) % r, q_scalars


def open(pp: ZeromorphKzgProverParam, poly, point: Sequence[int], eval: int, transcript, ops=GpuOps) -> int:
    """Zeromorph::open (zeromorph.rs:126-186), the non-sanity-check path.  Writes n quotient commitments, the commitment
    of q_hat and the opening of f at x (one more point) to the transcript.  Returns the remainder of `quotients`, i.e.
    poly(point) — what the sanity check at :152-154 compares with `eval`."""
    r = FR_MODULUS
    num_vars = len(point)
    if pp.degree() + 1 < len(poly):
        raise ValueError(f"Too large degree of poly to open (param supports degree up to {pp.degree()} but got {len(poly)})")
    point = [int(p) % r for p in point]
    q, remainder = ops.quotients(poly, point)                                              # :149
    owned = [q]
    try:
        transcript.write_commitments(ops.commit_quotients(pp.commit_pp, q, num_vars))      # batch_commit_and_write, :150
        y = transcript.squeeze_challenge()
        q_hat = ops.q_hat(q, _powers(y, num_vars))                                         # :158-168
        owned.append(q_hat)
        transcript.write_commitment(ops.commit(pp.commit_pp, q_hat))                       # commit_and_write, :169
        x = transcript.squeeze_challenge()
        z = transcript.squeeze_challenge()
        eval_scalar, q_scalars = eval_and_quotient_scalars(y, x, z, point)
        f = ops.f(poly, q_hat, q, z, eval_scalar * (int(eval) % r) % r, q_scalars)         # :175-180
        owned.append(f)
        univariate.open(pp.open_pp, f, x, transcript, ops)                                 # UnivariateKzg::open(&pp.open_pp, &f, .., &x, &ZERO), :185
    finally:
        for t in owned:
            ops.release(t)
    return remainder


def batch_open(pp: ZeromorphKzgProverParam, num_vars: int, polys: Sequence, points: Sequence[Sequence[int]], evals: Sequence[Tuple[int, int, int]],
               transcript) -> None:
    """Zeromorph::batch_open (zeromorph.rs:188-204) = additive::batch_open (pcs/multilinear.rs:134-235) with this PCS's
    open on g_prime; the non-sanity-check path hands open a zero evaluation (multilinear.rs:224-226), which leaves the
    proof unchanged (only f's constant term depends on it, the quotient by X - x does not)."""
    from . import kzg

    def open_g_prime(g_prime, challenges):
        open(pp, g_prime, challenges, 0, transcript)

    kzg.batch_open(None, num_vars, polys, points, evals, transcript, open_fn=open_g_prime)
