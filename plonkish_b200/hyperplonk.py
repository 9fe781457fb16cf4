"""Host-side mirror of the reference's HyperPlonk prover over the GPU entry points — the k = 24 "prove" of BASELINE.json.

Mirrors /root/reference/plonkish_backend/src/backend (M = Bn256, Pcs = MultilinearKzg, Keccak256 transcript):
  PlonkishCircuitInfo                       backend.rs:40-123
  vanilla_plonk_circuit_info                backend/hyperplonk/util.rs:30-49
  compose / max_degree / lookup_constraints /
  permutation_constraints / permutation_polys   backend/hyperplonk/preprocessor.rs:25-203
  HyperPlonk::preprocess                    backend/hyperplonk.rs:97-162
  HyperPlonk::prove                         backend/hyperplonk.rs:164-291
  instance_polys, prove_zero_check,
  prove_sum_check                           backend/hyperplonk/prover.rs:32-48, 347-409
  pcs_query, point_offset, points           backend/hyperplonk/verifier.rs:147-182
  rotation_eval_points (+ point pattern)    poly/multilinear.rs:478-545

Everything that touches 2^k field elements runs on the GPU: the commitments (MSM), permutation_z_polys, the tables of the
compiled zero-check expression (expression.py), the sum-check rounds and folds, the evaluations at the rotated points and
additive::batch_open.  The host keeps what is scalar work in the reference too: the transcript, the expression tree, the
interpolation of a round message.  Lookup arguments run on the GPU as well (lookup_compressed_polys / lookup_m_polys /
lookup_h_polys, prover.rs:50-250: csrc/lookup_kernels.cuh).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import kzg, sumcheck
from .expression import BooleanHypercube, CompiledExpression, Expression, Query, compile_expression
from .msm import (ResidentScalars, eq_table, fr_affine_table, fr_evaluate, fr_expression_table, lookup_h_poly, lookup_m_poly,
                  permutation_z_polys)
from .transcript import fr_to_montgomery

FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R = FR_MODULUS
_MONT = 1 << 256


def div_ceil(a: int, b: int) -> int:
    return -(-a // b)


@dataclass
class PlonkishCircuitInfo:
    """backend.rs:40-73.  preprocess_polys: [2^k, 4] Montgomery arrays (Vec<F>)."""
    k: int
    num_instances: List[int]
    preprocess_polys: List[np.ndarray]
    num_witness_polys: List[int]
    num_challenges: List[int]
    constraints: List[Expression]
    lookups: List[List[Tuple[Expression, Expression]]]
    permutations: List[List[Tuple[int, int]]]
    max_degree: Optional[int] = None

    def num_poly(self) -> int:  # backend.rs:108-112
        return len(self.num_instances) + len(self.preprocess_polys) + sum(self.num_witness_polys)

    def permutation_polys(self) -> List[int]:  # backend.rs:114-121
        return sorted({poly for cycle in self.permutations for poly, _ in cycle})


def vanilla_plonk_circuit_info(num_vars: int, num_instances: int, preprocess_polys: Sequence[np.ndarray],
                               permutations: List[List[Tuple[int, int]]]) -> PlonkishCircuitInfo:
    """backend/hyperplonk/util.rs:30-49."""
    assert len(preprocess_polys) == 5
    pi, q_l, q_r, q_m, q_o, q_c, w_l, w_r, w_o = (Expression.polynomial(p) for p in range(9))
    return PlonkishCircuitInfo(k=num_vars, num_instances=[num_instances], preprocess_polys=list(preprocess_polys), num_witness_polys=[3],
                               num_challenges=[0], constraints=[q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi], lookups=[],
                               permutations=permutations, max_degree=4)


# ------------------------------------------------------------------------------------ preprocessor.rs
def lookup_constraints(info: PlonkishCircuitInfo, beta: Expression, gamma: Expression):
    """preprocessor.rs:79-109."""
    m_offset = info.num_poly() + len(info.permutation_polys())
    h_offset = m_offset + len(info.lookups)
    constraints = []
    for i, lookup in enumerate(info.lookups):
        m, h = Expression.polynomial(m_offset + i), Expression.polynomial(h_offset + i)
        inp = Expression.distribute_powers([a for a, _ in lookup], beta)
        table = Expression.distribute_powers([b for _, b in lookup], beta)
        constraints.append(h * (inp + gamma) * (table + gamma) - (table + gamma) + m * (inp + gamma))
    sum_check = [Expression.polynomial(h_offset + i) for i in range(len(info.lookups))]
    return constraints, sum_check


def max_degree(info: PlonkishCircuitInfo, lookup_cs: Optional[Sequence[Expression]] = None) -> int:
    """preprocessor.rs:62-77."""
    if lookup_cs is None:
        lookup_cs = lookup_constraints(info, Expression.zero(), Expression.zero())[0]
    degrees = [c.degree() for c in info.constraints] + [c.degree() for c in lookup_cs]
    if info.max_degree is not None:
        degrees.append(info.max_degree)
    return max(degrees + [2])


def permutation_constraints(info: PlonkishCircuitInfo, max_deg: int, beta: Expression, gamma: Expression, num_builtin_witness_polys: int):
    """preprocessor.rs:111-170."""
    perm_polys = info.permutation_polys()
    chunk_size = max_deg - 1
    num_chunks = div_ceil(len(perm_polys), chunk_size)
    permutation_offset = info.num_poly()
    z_offset = permutation_offset + len(perm_polys) + num_builtin_witness_polys
    polys = [Expression.polynomial(idx) for idx in perm_polys]
    ids = [Expression.constant(idx << info.k) + Expression.identity() for idx in range(len(polys))]
    permutations = [Expression.polynomial(permutation_offset + i) for i in range(len(perm_polys))]
    zs = [Expression.polynomial(z_offset + i) for i in range(num_chunks)]
    z_0_next = Expression.polynomial(z_offset, 1)
    l_1, one = Expression.lagrange(1), Expression.one()

    def product(factors: List[Expression]) -> Expression:  # Product::product: reduce(|acc, item| acc * item)
        acc = factors[0]
        for f in factors[1:]:
            acc = acc * f
        return acc

    constraints = [l_1 * (zs[0] - one)] if zs else []
    for c in range(num_chunks):
        sl = slice(c * chunk_size, (c + 1) * chunk_size)
        z_lhs, z_rhs = zs[c], (zs[c + 1] if c + 1 < num_chunks else z_0_next)
        lhs = product([p + beta * i + gamma for p, i in zip(polys[sl], ids[sl])])
        rhs = product([p + beta * s + gamma for p, s in zip(polys[sl], permutations[sl])])
        constraints.append(z_lhs * lhs - z_rhs * rhs)
    return num_chunks, constraints


def compose(info: PlonkishCircuitInfo) -> Tuple[int, Expression]:
    """preprocessor.rs:25-60."""
    challenge_offset = sum(info.num_challenges)
    beta, gamma, alpha = (Expression.challenge(challenge_offset + i) for i in range(3))
    lookup_cs, lookup_zero_checks = lookup_constraints(info, beta, gamma)
    max_deg = max_degree(info, lookup_cs)
    num_z, perm_cs = permutation_constraints(info, max_deg, beta, gamma, 2 * len(info.lookups))
    constraints = list(info.constraints) + lookup_cs + perm_cs
    zero_check_on_every_row = Expression.distribute_powers(constraints, alpha) * Expression.eq_xy(0)
    return num_z, Expression.distribute_powers(lookup_zero_checks + [zero_check_on_every_row], alpha)


def permutation_polys_canonical(num_vars: int, perm_polys: Sequence[int], cycles: Sequence[Sequence[Tuple[int, int]]]) -> List[np.ndarray]:
    """preprocessor.rs:172-203 on canonical integers (every value is a cell id < 2^(k + 5)): uint64 arrays of 2^k."""
    poly_index = {poly: idx for idx, poly in enumerate(perm_polys)}
    n = 1 << num_vars
    permutations = [np.arange(idx << num_vars, (idx << num_vars) + n, dtype=np.uint64) for idx in range(len(perm_polys))]
    for cycle in cycles:
        i0, j0 = cycle[0]
        last = permutations[poly_index[i0]][j0]
        for i, j in list(cycle[1:]) + [cycle[0]]:
            assert j != 0
            cell = permutations[poly_index[i]]
            cell[j], last = last, cell[j]
    return permutations


def canonical_u64_to_resident(values: np.ndarray, device: int = 0) -> ResidentScalars:
    """F::from(u64) for a whole column: upload the raw integers, multiply by R^2 in the Montgomery kernel (a * R^2 / R)."""
    raw = np.zeros((values.shape[0], 4), dtype=np.uint64)
    raw[:, 0] = values
    tmp = ResidentScalars(raw, device=device)
    try:
        return fr_affine_table(values.shape[0].bit_length() - 1, [tmp], np.stack([fr_to_montgomery(_MONT % R)]))
    finally:
        tmp.release()


# ------------------------------------------------------------------------------------ hyperplonk.rs
def _pcs_of(pcs_pp):
    """The PolynomialCommitmentScheme a prover parameter belongs to (HyperPlonk<Pcs>, hyperplonk.rs:76-95; pcs.rs:22-130):
    MultilinearKzg (kzg.py), Zeromorph<UnivariateKzg> (zeromorph.py) or Gemini<UnivariateKzg> (gemini.py) — the three the
    reference's tests instantiate over Bn256 (hyperplonk.rs:424-426).  Each module exposes commit / batch_commit(keep=True) /
    batch_open."""
    from . import gemini, zeromorph

    if isinstance(pcs_pp, zeromorph.ZeromorphKzgProverParam):
        return zeromorph
    return gemini if isinstance(pcs_pp, gemini.GeminiKzgProverParam) else kzg


@dataclass
class HyperPlonkProverParam:
    """backend/hyperplonk.rs:38-57; polynomials resident in HBM."""
    pcs: object  # kzg.MultilinearKzgProverParam, zeromorph.ZeromorphKzgProverParam or gemini.GeminiKzgProverParam
    num_instances: List[int]
    num_witness_polys: List[int]
    num_challenges: List[int]
    lookups: List
    num_permutation_z_polys: int
    num_vars: int
    expression: Expression
    preprocess_polys: List[ResidentScalars]
    preprocess_comms: List[np.ndarray]
    permutation_polys: List[Tuple[int, ResidentScalars]]
    permutation_comms: List[np.ndarray]

    def release(self) -> None:
        for p in self.preprocess_polys + [p for _, p in self.permutation_polys]:
            p.release()


@dataclass
class HyperPlonkVerifierParam:
    """backend/hyperplonk.rs:59-74 without the PCS half (the G2 side of the SRS is out of scope)."""
    num_instances: List[int]
    num_witness_polys: List[int]
    num_challenges: List[int]
    num_lookups: int
    num_permutation_z_polys: int
    num_vars: int
    expression: Expression
    preprocess_comms: List[np.ndarray]
    permutation_comms: List[Tuple[int, np.ndarray]]


def preprocess(pcs_pp, info: PlonkishCircuitInfo, permutation_columns: Optional[Sequence[np.ndarray]] = None,
               device: int = 0) -> Tuple[HyperPlonkProverParam, HyperPlonkVerifierParam]:
    """HyperPlonk::preprocess (hyperplonk.rs:97-162): commit the preprocessed and the permutation polynomials, compose the
    expression.  permutation_columns: the permutation polynomials as canonical uint64 columns when the caller already has
    them (a 2^24-row circuit builds them vectorised); default = permutation_polys_canonical over info.permutations."""
    num_vars = info.k
    preprocess_polys = [p if isinstance(p, ResidentScalars) else ResidentScalars(p, device=device) for p in info.preprocess_polys]
    pcs = _pcs_of(pcs_pp)
    preprocess_comms = [pcs.commit(pcs_pp, p) for p in preprocess_polys]
    perm_idx = info.permutation_polys()
    cols = permutation_columns if permutation_columns is not None else permutation_polys_canonical(num_vars, perm_idx, info.permutations)
    assert len(cols) == len(perm_idx)
    permutation_polys = [canonical_u64_to_resident(np.ascontiguousarray(c, dtype=np.uint64), device) for c in cols]
    permutation_comms = [pcs.commit(pcs_pp, p) for p in permutation_polys]
    num_z, expression = compose(info)
    vp = HyperPlonkVerifierParam(list(info.num_instances), list(info.num_witness_polys), list(info.num_challenges), len(info.lookups), num_z,
                                 num_vars, expression, preprocess_comms, list(zip(perm_idx, permutation_comms)))
    pp = HyperPlonkProverParam(pcs_pp, list(info.num_instances), list(info.num_witness_polys), list(info.num_challenges), list(info.lookups),
                               num_z, num_vars, expression, preprocess_polys, preprocess_comms, list(zip(perm_idx, permutation_polys)),
                               permutation_comms)
    return pp, vp


def row_mapping(k: int) -> List[int]:
    """WitnessEncoding for HyperPlonk (hyperplonk.rs:365-369): bh.iter().skip(1).chain([0])."""
    rows = list(BooleanHypercube(k).iter())
    return rows[1:] + [0]


def instance_polys(num_vars: int, instances: Sequence[Sequence[int]], device: int = 0) -> List[ResidentScalars]:
    """prover.rs:32-48: instance i of a column sits on row bh[i + 1]; everything else is zero.  Built in HBM (a zero table
    plus a handful of rows), instances are canonical integers."""
    bh = BooleanHypercube(num_vars)
    out = []
    for column in instances:
        rows, b = [], 1
        for _ in column:
            rows.append(b)
            b = bh.next(b)
        out.append(fr_affine_table(num_vars, sparse_rows=rows, sparse_values=np.stack([fr_to_montgomery(v) for v in column]) if rows else None,
                                   device=device))
    return out


# ---------------------------------------------------------------------- poly/multilinear.rs:478-545
def rotation_eval_point_pattern(num_vars: int, distance: int, is_next: bool) -> List[int]:
    bh = BooleanHypercube(num_vars)
    remainder = bh.primitive if is_next else bh.x_inv
    pattern = [0] * (1 << distance)
    for depth in range(distance):
        step = 1 << (distance - depth)
        for e in range(0, len(pattern), step):
            o = e + (step >> 1)
            rotated = pattern[e] << 1 if is_next else pattern[e] >> 1
            pattern[o] = rotated ^ remainder
            pattern[e] = rotated
    return pattern


def rotation_eval_points(x: Sequence[int], rotation: int) -> List[List[int]]:
    if rotation == 0:
        return [list(x)]
    distance = abs(rotation)
    num_x = len(x) - distance
    bit = lambda pat, i: (pat >> i) & 1  # noqa: E731
    if rotation < 0:
        pattern = rotation_eval_point_pattern(len(x), distance, False)
        xs = list(x[distance:])
        return [[(1 - xs[i]) % R if bit(pat, i) else xs[i] for i in range(num_x)] + [bit(pat, i + num_x) for i in range(distance)] for pat in pattern]
    pattern = rotation_eval_point_pattern(len(x), distance, True)
    xs = list(x[:num_x])
    return [[bit(pat, i) for i in range(distance)] + [(1 - xs[i]) % R if bit(pat, i + distance) else xs[i] for i in range(num_x)] for pat in pattern]


# ---------------------------------------------------------------------- verifier.rs:147-182
def pcs_query(expression: Expression, num_instance_poly: int) -> List[Query]:
    return [q for q in expression.used_query() if q.poly >= num_instance_poly]


def point_offset(queries: Sequence[Query]) -> Dict[int, int]:
    rotations = sorted({q.rotation for q in queries})
    offsets, off = {}, 0
    for r in rotations:
        offsets[r] = off
        off += 1 << abs(r)
    return offsets


def points(queries: Sequence[Query], x: Sequence[int]) -> List[List[int]]:
    return [p for r in sorted({q.rotation for q in queries}) for p in rotation_eval_points(x, r)]


# ---------------------------------------------------------------------- the compiled zero check on the GPU
@dataclass
class SumCheckTables:
    tables: List[ResidentScalars]
    owned: List[ResidentScalars] = field(default_factory=list)
    query_table: Dict[int, int] = field(default_factory=dict)   # polynomial index -> table holding it unrotated

    def release(self) -> None:
        for t in self.owned:
            t.release()
        self.owned = []


def build_tables(compiled: CompiledExpression, num_vars: int, polys: Sequence[ResidentScalars], ys: Sequence[Sequence[int]],
                 extra_polys: Sequence[int] = ()) -> SumCheckTables:
    """Materialise the atoms of a compiled expression (classic.rs:40-83 keeps them implicit): a polynomial queried at the
    current row is its own table; eq_xy(y), identity, Lagrange(i) (a single row, classic.rs:44-55), rotated queries
    (rotation_map, classic.rs:105-125) and every other linear factor come out of plonkish_cuda_fr_affine_table.
    extra_polys: polynomials whose evaluation at the sum check's point is wanted although no term reads them as a plain
    table (they ride along through the folds; classic.rs:143-149 returns the evaluation of every polynomial)."""
    bh = BooleanHypercube(num_vars)
    device = polys[0].device
    out = SumCheckTables([])
    eq_tables: Dict[int, ResidentScalars] = {}

    def eq_of(idx: int) -> ResidentScalars:
        if idx not in eq_tables:
            eq_tables[idx] = eq_table(np.stack([fr_to_montgomery(v) for v in ys[idx]]), device=device)
            out.owned.append(eq_tables[idx])
        return eq_tables[idx]

    for atom in compiled.atoms:
        if atom.is_leaf() and atom.leaf()[0] == "poly" and atom.leaf()[2] == 0:
            out.query_table[atom.leaf()[1]] = len(out.tables)
            out.tables.append(polys[atom.leaf()[1]])
            continue
        if atom.is_leaf() and atom.leaf()[0] == "eq_xy":
            out.tables.append(eq_of(atom.leaf()[1]))
            continue
        srcs, coeffs, rots, rows, vals, id_coeff = [], [], [], [], [], None
        for leaf, c in atom.terms.items():
            if leaf[0] == "poly":
                srcs.append(polys[leaf[1]]); coeffs.append(c); rots.append(leaf[2])
            elif leaf[0] == "eq_xy":
                srcs.append(eq_of(leaf[1])); coeffs.append(c); rots.append(0)
            elif leaf[0] == "identity":
                id_coeff = c
            else:  # ("lagrange", i): one at row bh[i mod 2^k] (classic.rs:50-53)
                rows.append(bh.nth(leaf[1] % (1 << num_vars))); vals.append(c)
        t = fr_affine_table(num_vars, srcs, np.stack([fr_to_montgomery(c) for c in coeffs]) if srcs else None, rots,
                            constant=fr_to_montgomery(atom.const) if atom.const else None,
                            identity_coeff=None if id_coeff is None else fr_to_montgomery(id_coeff), sparse_rows=rows,
                            sparse_values=np.stack([fr_to_montgomery(v) for v in vals]) if rows else None, device=device)
        out.owned.append(t)
        out.tables.append(t)
    for p in extra_polys:
        if p not in out.query_table:
            out.query_table[p] = len(out.tables)
            out.tables.append(polys[p])
    return out


def lookup_compressed_polys(lookups, polys: Sequence[ResidentScalars], challenges: Sequence[int], betas: Sequence[int]) -> List[List[ResidentScalars]]:
    """prover.rs:50-137: per lookup the compressed input and table polynomials, sum_i betas[i] * expression_i evaluated on
    every row (identity = row, Lagrange = one row, rotated queries by BooleanHypercube::rotate).  The expression tree is
    compiled like the zero check's and evaluated row-wise by k_expr_rows."""
    num_vars = polys[0].n.bit_length() - 1
    out = []
    for lookup in lookups:
        pair = []
        for exprs in ([inp for inp, _ in lookup], [tab for _, tab in lookup]):
            total = exprs[0] * betas[0]
            for e, b in zip(exprs[1:], betas[1:]):
                total = total + e * b
            compiled = compile_expression(total, challenges)
            if not compiled.terms:
                pair.append(fr_affine_table(num_vars, device=polys[0].device))      # identically zero
                continue
            st = build_tables(compiled, num_vars, polys, [])
            try:
                pair.append(fr_expression_table(st.tables, [(fr_to_montgomery(c), idx) for c, idx in compiled.terms], compiled.common))
            finally:
                st.release()
        out.append(pair)
    return out


def prove_sum_check(num_instance_poly: int, expression: Expression, claimed_sum: int, polys: Sequence[ResidentScalars], challenges: Sequence[int],
                    y: Sequence[int], transcript) -> Tuple[List[List[int]], List[Tuple[int, int, int]]]:
    """prover.rs:367-409: ClassicSumCheck<EvaluationsProver>::prove over the virtual polynomial, then the evaluations of
    every queried polynomial — at the sum check's point x, and for a rotated query at the 2^distance points of
    rotation_eval_points (evaluate_for_rotation, poly/multilinear.rs:191-263) — written to the transcript.
    Returns (points, [(poly, point index, value)])."""
    num_vars = polys[0].n.bit_length() - 1
    assert num_vars > 0 and expression.max_used_rotation_distance() <= num_vars                       # classic.rs:42
    compiled = compile_expression(expression, challenges)
    assert compiled.degree == expression.degree(), "the compiled terms must have the degree the reference sends messages for"
    queries = pcs_query(expression, num_instance_poly)
    st = build_tables(compiled, num_vars, polys, [y], extra_polys=sorted({q.poly for q in queries}))
    try:
        terms = [(fr_to_montgomery(c), idx) for c, idx in compiled.terms]
        # the zero check's eq_xy(0) = eq(x, y) is the common factor: its rounds run factored (one evaluation point fewer
        # per pair, the same messages); any other expression takes the plain rounds
        zc = list(y) if compiled.common_eq_xy() == 0 else None
        x, table_evals = sumcheck.prove_to_transcript(st.tables, terms, claimed_sum, transcript, common=compiled.common, zero_check_point=zc)
    finally:
        st.release()
    offsets = point_offset(queries)
    evals: List[Tuple[int, int, int]] = []
    for q in queries:                                                                               # prover.rs:392-403
        if q.rotation == 0:
            values = [table_evals[st.query_table[q.poly]]]
        else:
            pts = rotation_eval_points(x, q.rotation)
            got = fr_evaluate(polys[q.poly], np.stack([np.stack([fr_to_montgomery(v) for v in pt]) for pt in pts]))
            values = [sumcheck._to_int(row) for row in got]
        evals.extend((q.poly, offsets[q.rotation] + j, v) for j, v in enumerate(values))
    transcript.write_field_elements([v for _, _, v in evals])
    return points(queries, x), evals


def prove(pp: HyperPlonkProverParam, circuit, transcript, marks: Optional[list] = None) -> None:
    """HyperPlonk::prove (hyperplonk.rs:164-291).  circuit: .instances() -> canonical integers per instance column,
    .synthesize(round, challenges) -> witness polynomials of that round as [2^k, 4] Montgomery arrays (PlonkishCircuit,
    backend.rs:125-133).  The proof goes to `transcript` (Keccak256Transcript).  marks: optional list receiving
    (label, perf_counter()) at the reference's timer boundaries (hyperplonk.rs:192-288)."""
    import time

    def mark(label: str) -> None:
        if marks is not None:
            marks.append((label, time.perf_counter()))

    mark("start")
    pcs = _pcs_of(pp.pcs)
    instances = circuit.instances()
    assert len(instances) == len(pp.num_instances)
    for num, column in zip(pp.num_instances, instances):
        assert len(column) == num
        for v in column:
            transcript.common_field_element(v)
    device = pp.preprocess_polys[0].device if pp.preprocess_polys else 0
    inst_polys = instance_polys(pp.num_vars, instances, device)
    owned: List[ResidentScalars] = list(inst_polys)
    try:
        # Round 0..n (hyperplonk.rs:183-209)
        witness_polys: List[ResidentScalars] = []
        challenges: List[int] = []
        for rnd, (num_w, num_c) in enumerate(zip(pp.num_witness_polys, pp.num_challenges)):
            host = circuit.synthesize(rnd, challenges)
            assert len(host) == num_w
            comms, resident = pcs.batch_commit(pp.pcs, host, keep=True)          # Pcs::batch_commit_and_write
            owned.extend(resident)
            transcript.write_commitments(comms)
            witness_polys.extend(resident)
            challenges.extend(transcript.squeeze_challenges(num_c))
        mark("witness batch_commit")
        polys = inst_polys + pp.preprocess_polys + witness_polys
        # Round n (hyperplonk.rs:213-227)
        beta = transcript.squeeze_challenge()
        max_lookup_width = max([len(lookup) for lookup in pp.lookups] or [0])
        betas = [pow(beta, i, R) for i in range(max_lookup_width)]                # powers(beta).take(max_lookup_width)
        compressed = lookup_compressed_polys(pp.lookups, polys, challenges, betas)
        owned.extend(p for pair in compressed for p in pair)
        lookup_m_polys = [lookup_m_poly(inp, tab) for inp, tab in compressed]        # Err(InvalidSnark) -> PlonkishCudaError
        owned.extend(lookup_m_polys)
        transcript.write_commitments([pcs.commit(pp.pcs, m) for m in lookup_m_polys])
        mark("lookup m polys + commit")
        # Round n+1 (hyperplonk.rs:231-252)
        gamma = transcript.squeeze_challenge()
        lookup_h_polys = [lookup_h_poly(inp, tab, m, fr_to_montgomery(gamma)) for (inp, tab), m in zip(compressed, lookup_m_polys)]
        owned.extend(lookup_h_polys)
        z_polys: List[ResidentScalars] = []
        if pp.permutation_polys:
            z_polys = permutation_z_polys(pp.num_permutation_z_polys, [polys[idx] for idx, _ in pp.permutation_polys],
                                          [p for _, p in pp.permutation_polys], fr_to_montgomery(beta), fr_to_montgomery(gamma))
            owned.extend(z_polys)
        mark("lookup h polys + permutation_z_polys")
        transcript.write_commitments([pcs.commit(pp.pcs, p) for p in lookup_h_polys + z_polys])
        mark("h / z commit")
        # Round n+2 (hyperplonk.rs:256-273)
        alpha = transcript.squeeze_challenge()
        y = transcript.squeeze_challenges(pp.num_vars)
        polys = polys + [p for _, p in pp.permutation_polys] + lookup_m_polys + lookup_h_polys + z_polys
        challenges = challenges + [beta, gamma, alpha]
        pts, evals = prove_sum_check(len(pp.num_instances), pp.expression, 0, polys, challenges, y, transcript)   # prove_zero_check
        mark("zero check + evals")
        # PCS open (hyperplonk.rs:277-288)
        pcs.batch_open(pp.pcs, pp.num_vars, polys, pts, evals, transcript)
        mark("pcs_batch_open")
    finally:
        for p in owned:
            p.release()
        mark("release")
