"""CPU suite for the HyperPlonk mirror (plonkish_b200/expression.py, hyperplonk.py host logic) and its all-integer
reference (tests/hyperplonk_ref.py):
  * compose() against the expression the reference's own tests spell out (preprocessor.rs:216-303);
  * the expression compiler against the tree it came from, on random values;
  * rotation_eval_points against the reference's own property (poly/multilinear.rs:682-712);
  * permutation polynomials, row mapping, query bookkeeping against the integer restatement;
  * the integer prover against the integer verifier (accepts; rejects a flipped byte): the pair that checks the GPU proofs;
  * the affine-table kernel, compiled by g++ for the emulator, against Python integers."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import hyperplonk_ref as ref
from conftest import ROOT
from oracle import bigint_ref as br
from plonkish_b200 import expression as ex
from plonkish_b200.expression import BooleanHypercube, Expression, Query, compile_expression

R = br.R
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


def _hp():
    from plonkish_b200 import hyperplonk

    return hyperplonk


def _fe(rng):
    return int.from_bytes(rng.bytes(40), "little") % R


# ------------------------------------------------------------------------------------------------ compose
def _vanilla_expected(num_vars):
    """preprocessor.rs:216-252, test compose_vanilla_plonk."""
    pi, q_l, q_r, q_m, q_o, q_c, w_l, w_r, w_o, s_1, s_2, s_3 = (Expression.polynomial(p) for p in range(12))
    z, z_next = Expression.polynomial(12, 0), Expression.polynomial(12, 1)
    beta, gamma, alpha = (Expression.challenge(i) for i in range(3))
    id_1, id_2, id_3 = (Expression.constant(idx << num_vars) + Expression.identity() for idx in range(3))
    l_1, one = Expression.lagrange(1), Expression.one()
    constraints = [
        q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi,
        l_1 * (z - one),
        (z * ((w_l + beta * id_1 + gamma) * (w_r + beta * id_2 + gamma) * (w_o + beta * id_3 + gamma)))
        - (z_next * ((w_l + beta * s_1 + gamma) * (w_r + beta * s_2 + gamma) * (w_o + beta * s_3 + gamma))),
    ]
    return Expression.distribute_powers(constraints, alpha) * Expression.eq_xy(0)


def test_compose_vanilla_plonk_is_the_expression_the_reference_test_spells_out():
    hp = _hp()
    num_vars = 3
    info = hp.vanilla_plonk_circuit_info(num_vars, 0, [np.zeros((8, 4), dtype=np.uint64)] * 5, [[(6, 1)], [(7, 1)], [(8, 1)]])  # util.rs:51-61
    num_z, expression = hp.compose(info)
    assert num_z == 1
    assert expression == _vanilla_expected(num_vars)
    assert not (expression == _vanilla_expected(num_vars + 1))
    assert expression.degree() == 5
    assert expression.used_query() == [Query(p, 0) for p in range(13)] + [Query(12, 1)]
    assert expression.used_lagrange() == [1] and expression.used_rotation() == [0, 1] and expression.used_challenge() == [0, 1, 2]


def test_compose_vanilla_plonk_with_lookup_is_the_expression_the_reference_test_spells_out():
    """preprocessor.rs:254-303 (compose_vanilla_plonk_with_lookup); the circuit info of util.rs:63-86."""
    hp = _hp()
    num_vars = 3
    pi, q_l, q_r, q_m, q_o, q_c, q_lookup, t_l, t_r, t_o, w_l, w_r, w_o, s_1, s_2, s_3 = (Expression.polynomial(p) for p in range(16))
    info = hp.PlonkishCircuitInfo(
        k=num_vars, num_instances=[0], preprocess_polys=[np.zeros((8, 4), dtype=np.uint64)] * 9, num_witness_polys=[3], num_challenges=[0],
        constraints=[q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi],
        lookups=[[(q_lookup * w_l, t_l), (q_lookup * w_r, t_r), (q_lookup * w_o, t_o)]],
        permutations=[[(10, 1)], [(11, 1)], [(12, 1)]], max_degree=4)
    num_z, expression = hp.compose(info)
    lookup_m, lookup_h = Expression.polynomial(16), Expression.polynomial(17)
    perm_z, perm_z_next = Expression.polynomial(18, 0), Expression.polynomial(18, 1)
    beta, gamma, alpha = (Expression.challenge(i) for i in range(3))
    id_1, id_2, id_3 = (Expression.constant(idx << num_vars) + Expression.identity() for idx in range(3))
    l_1, one = Expression.lagrange(1), Expression.one()
    lookup_input = Expression.distribute_powers([q_lookup * w for w in (w_l, w_r, w_o)], beta)
    lookup_table = Expression.distribute_powers([t_l, t_r, t_o], beta)
    constraints = [
        q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi,
        lookup_h * (lookup_input + gamma) * (lookup_table + gamma) - (lookup_table + gamma) + lookup_m * (lookup_input + gamma),
        l_1 * (perm_z - one),
        (perm_z * ((w_l + beta * id_1 + gamma) * (w_r + beta * id_2 + gamma) * (w_o + beta * id_3 + gamma)))
        - (perm_z_next * ((w_l + beta * s_1 + gamma) * (w_r + beta * s_2 + gamma) * (w_o + beta * s_3 + gamma))),
    ]
    zero_check_on_every_row = Expression.distribute_powers(constraints, alpha) * Expression.eq_xy(0)
    assert num_z == 1
    assert expression == Expression.distribute_powers([lookup_h, zero_check_on_every_row], alpha)


# ------------------------------------------------------------------------------------------------ compiler
def _leaf_values(expression, rng):
    vals = {}

    def leaf(l):
        if l not in vals:
            vals[l] = _fe(rng)
        return vals[l]

    return leaf


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_compiled_vanilla_plonk_expression_equals_the_tree_and_keeps_linear_factors_whole(seed):
    hp = _hp()
    rng = np.random.default_rng(seed)
    num_vars = 5
    info = hp.vanilla_plonk_circuit_info(num_vars, 0, [np.zeros((32, 4), dtype=np.uint64)] * 5, [[(6, 1)], [(7, 1)], [(8, 1)]])
    _, expression = hp.compose(info)
    challenges = [_fe(rng) for _ in range(3)]
    compiled = compile_expression(expression, challenges)
    leaf = _leaf_values(expression, rng)
    want = expression.evaluate_field(lambda p: leaf(tuple(p)), lambda q: leaf(("poly", q.poly, q.rotation)), challenges)
    assert compiled.value(leaf) == want
    # shape: eq_xy is the common factor, the degree is the reference's, the permutation factors are six whole atoms
    assert compiled.common >= 0 and compiled.atoms[compiled.common].leaf() == ("eq_xy", 0)
    assert compiled.degree == expression.degree() == 5
    assert len(compiled.terms) == 9 and max(len(f) for _, f in compiled.terms) == 4
    assert len(compiled.atoms) <= 32 and len(compiled.terms) <= 32
    whole = [a for a in compiled.atoms if len(a.terms) == 2 and a.const]      # w + beta * (id | sigma) + gamma, scaled or not
    assert len(whole) == 6
    # no coefficient multiplication is left for the kernel on the degree-4 terms
    assert all(c == 1 for c, f in compiled.terms if len(f) == 4)


def test_compiler_on_random_expression_trees():
    rng = np.random.default_rng(11)

    def rand_expr(depth):
        if depth == 0 or rng.integers(5) == 0:
            kind = int(rng.integers(6))
            if kind == 0:
                return Expression.constant(_fe(rng) if rng.integers(2) else int(rng.integers(3)))
            if kind == 1:
                return Expression.polynomial(int(rng.integers(4)), int(rng.integers(-1, 2)))
            if kind == 2:
                return Expression.challenge(int(rng.integers(2)))
            if kind == 3:
                return Expression.identity()
            if kind == 4:
                return Expression.lagrange(int(rng.integers(-2, 3)))
            return Expression.eq_xy(0)
        op = int(rng.integers(6))
        a = rand_expr(depth - 1)
        if op == 0:
            return -a
        if op == 1:
            return a * _fe(rng)
        if op == 2:
            return Expression.distribute_powers([a] + [rand_expr(depth - 1) for _ in range(int(rng.integers(0, 3)))], rand_expr(0))
        b = rand_expr(depth - 1)
        return a + b if op == 3 else (a - b if op == 4 else a * b)

    for _ in range(60):
        e = rand_expr(4)
        challenges = [_fe(rng), _fe(rng)]
        leaf = _leaf_values(e, rng)
        want = e.evaluate_field(lambda p: leaf(tuple(p)), lambda q: leaf(("poly", q.poly, q.rotation)), challenges)
        compiled = compile_expression(e, challenges)
        assert compiled.value(leaf) == want
        assert compiled.degree <= e.degree()      # cancellation can only lower it


# ------------------------------------------------------------------------------------------------ host bookkeeping
def _rotation_eval(x, rotation, evals_for_rotation):
    """poly/multilinear.rs:433-476 with the coefficient pattern of :547-570 (the verifier's side)."""
    if rotation == 0:
        return evals_for_rotation[0]
    num_vars, distance = len(x), abs(rotation)
    bh = BooleanHypercube(num_vars)
    is_next = rotation > 0
    remainder = bh.primitive - (1 << num_vars) if is_next else bh.x_inv << distance
    pattern = [0] * (1 << (distance - 1))
    for depth in range(distance - 1):
        step = 1 << (distance - depth - 1)
        for e in range(0, len(pattern), step):
            rotated = pattern[e] << 1 if is_next else pattern[e] >> 1
            pattern[e + (step >> 1)] = rotated ^ remainder
            pattern[e] = rotated
    if is_next:
        nths, xs = list(range(num_vars - 1, num_vars - 1 + distance)), list(x[num_vars - distance:])
    else:
        nths, xs = list(range(distance, 0, -1)), list(reversed(x[:distance]))
    evals = list(evals_for_rotation)
    for idx, (x_i, nth) in enumerate(zip(xs, nths)):
        bits = [(pat >> nth) & 1 for pat in pattern[:: 1 << idx]]
        evals = [((e0 - e1) * x_i + e1) % R if bit else ((e1 - e0) * x_i + e0) % R for bit, (e0, e1) in zip(bits, zip(evals[0::2], evals[1::2]))]
    return evals[0]


def test_rotation_eval_points_satisfy_the_reference_property():
    """poly/multilinear.rs:682-712 (test evaluate_for_rotation): rotation_eval over the evaluations at the rotated points
    is the rotated polynomial at x."""
    hp = _hp()
    rng = np.random.default_rng(3)
    for num_vars in range(1, 8):
        bh = BooleanHypercube(num_vars)
        n = 1 << num_vars
        f = [_fe(rng) for _ in range(n)]
        fs = [f]
        for _ in range(num_vars - 1):
            fs.append([fs[-1][bh.rotate(b, 1)] for b in range(n)])
        x = [_fe(rng) for _ in range(num_vars)]
        for rotation in range(-num_vars + 1, num_vars):
            base, rotated = (fs[-1], fs[len(fs) - abs(rotation) - 1]) if rotation < 0 else (fs[0], fs[rotation])
            pts = hp.rotation_eval_points(x, rotation)
            assert len(pts) == 1 << abs(rotation)
            evals = [ref.evaluate_multilinear(base, pt) for pt in pts]
            assert _rotation_eval(x, rotation, evals) == ref.evaluate_multilinear(rotated, x)
        if num_vars >= 2:
            assert hp.rotation_eval_points(x, 1) == ref.rotation_eval_points_next(x, num_vars)


def test_boolean_hypercube_and_permutation_bookkeeping():
    hp = _hp()
    for k in range(1, 11):
        bh = BooleanHypercube(k)
        order = list(bh.iter())
        assert order == ref.bh_iter(k) and sorted(order) == list(range(1 << k))
        assert all(bh.nth(i) == order[i] for i in range(1 << k))
        assert all(bh.rotate(bh.rotate(b, 3), -3) == b for b in range(1 << k))
        assert hp.row_mapping(k) == order[1:] + [0]
    rng = np.random.default_rng(9)
    k = 6
    _, _, _, cycles = ref.rand_vanilla_plonk_circuit(k, rng)
    want = ref.permutation_polys(k, [6, 7, 8], cycles)
    got = hp.permutation_polys_canonical(k, [6, 7, 8], cycles)
    assert [list(map(int, g)) for g in got] == want
    assert any(w != list(range(i << k, (i + 1) << k)) for i, w in enumerate(want))
    info = hp.vanilla_plonk_circuit_info(k, k, [np.zeros((64, 4), dtype=np.uint64)] * 5, cycles)
    assert info.permutation_polys() == [6, 7, 8] and info.num_poly() == 9
    _, expression = hp.compose(info)
    queries = hp.pcs_query(expression, 1)
    assert queries == [Query(p, 0) for p in range(1, 13)] + [Query(12, 1)]
    assert hp.point_offset(queries) == {0: 0, 1: 1}
    x = [_fe(rng) for _ in range(k)]
    assert hp.points(queries, x) == [x] + ref.rotation_eval_points_next(x, k)


# ------------------------------------------------------------------------------------------------ reference prover vs reference verifier
@pytest.mark.parametrize("k", [2, 3, 4])
def test_integer_prover_is_accepted_by_the_integer_verifier(oracle, k):
    from plonkish_b200.transcript import Keccak256Transcript

    rng = np.random.default_rng(40 + k)
    ss = [_fe(rng) for _ in range(k)]
    eq_scalars = oracle.kzg_eq_scalars(ref.mont_rows(ss))
    eqs_host = [oracle.fixed_base_msm(oracle.generator(), e) for e in eq_scalars]
    commit = lambda f: oracle.variable_base_msm(ref.mont_rows(f), eqs_host[k])  # noqa: E731
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_circuit(k, rng)
    sigmas = ref.permutation_polys(k, [6, 7, 8], cycles)
    # the circuit is satisfied: gate on every row, copies equal
    for b in range(1 << k):
        pi_b = dict(zip(ref.bh_iter(k)[1:], instances)).get(b, 0)
        q_l, q_r, q_m, q_o, q_c = (p[b] for p in preprocess)
        w_l, w_r, w_o = (w[b] for w in witness)
        assert (q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi_b) % R == 0
    t = Keccak256Transcript()
    state = ref.prove_reference(commit, ref.oracle_batch_open(oracle, eqs_host, k), k, instances, preprocess, witness, sigmas, t)
    proof = t.into_proof()
    # 3 + 1 commitments, k messages of 6, 14 evaluations, batch_open: k messages of 3 and k quotient commitments
    assert len(proof) == 4 * 64 + k * 6 * 32 + 14 * 32 + k * 3 * 32 + k * 64
    affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
    pre_comms = [affine(commit(p)) for p in preprocess]
    perm_comms = [affine(commit(s)) for s in sigmas]
    ref.verify_reference(oracle.keccak256, ss, k, instances, pre_comms, perm_comms, proof)
    # the grand product closes: z(last row) * factor(last row) = 1, i.e. z_next at the first row equals 1 there
    order = ref.bh_iter(k)
    assert state["z"][order[1]] == 1 and state["z"][0] == 0
    # a flipped byte anywhere a field element lives is rejected
    for pos in (4 * 64 + 7, 4 * 64 + k * 6 * 32 + 40, len(proof) - k * 64 - 9):
        bad = bytearray(proof)
        bad[pos] ^= 1
        with pytest.raises(AssertionError):
            ref.verify_reference(oracle.keccak256, ss, k, instances, pre_comms, perm_comms, bytes(bad))
    # wrong public input
    with pytest.raises(AssertionError):
        ref.verify_reference(oracle.keccak256, ss, k, [(instances[0] + 1) % R] + instances[1:], pre_comms, perm_comms, proof)


# ------------------------------------------------------------------------------------------------ emulated kernel
@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", EMUL_DIR], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(EMUL_DIR, "libemul_msm.so"))
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    lib.emul_fr_affine.argtypes = [vp, vp, vp, u32, u32, vp, vp, vp, vp, u32, vp]
    return lib


def affine_table_python(k, polys, coeffs, rotations, constant, id_coeff, rows, values):
    bh = BooleanHypercube(k)
    out = []
    for b in range(1 << k):
        v = constant + id_coeff * b
        for p, c, r in zip(polys, coeffs, rotations):
            v += c * p[bh.rotate(b, r)]
        out.append(v % R)
    for r_, v in zip(rows, values):
        out[r_] = (out[r_] + v) % R
    return out


@pytest.mark.parametrize("k", [1, 4, 9])
def test_emulated_affine_table_kernel_matches_python_integers(emul, k):
    rng = np.random.default_rng(70 + k)
    n = 1 << k
    for count, use_const, use_id, nrows in [(0, True, True, 2), (1, False, False, 0), (3, True, False, 1), (8, True, True, 3)]:
        polys = [[_fe(rng) for _ in range(n)] for _ in range(count)]
        coeffs = [1 if i == 1 else _fe(rng) for i in range(count)]
        rotations = [int(rng.integers(-min(k, 3), min(k, 3) + 1)) for _ in range(count)]
        constant, id_coeff = (_fe(rng) if use_const else 0), (_fe(rng) if use_id else 0)
        rows = [int(r_) for r_ in rng.choice(n, size=min(nrows, n), replace=False)]
        values = [_fe(rng) for _ in rows]
        arrs = [ref.mont_rows(p) for p in polys]
        ptrs = (ctypes.c_void_p * max(count, 1))(*[a.ctypes.data for a in arrs])
        rot = np.array(rotations or [0], dtype=np.int32)
        cs = ref.mont_rows(coeffs) if count else np.zeros((1, 4), dtype=np.uint64)
        cm, im = ref.to_mont(constant), ref.to_mont(id_coeff)
        rw = np.array(rows or [0], dtype=np.uint64)
        vs = ref.mont_rows(values) if rows else np.zeros((1, 4), dtype=np.uint64)
        out = np.zeros((n, 4), dtype=np.uint64)
        emul.emul_fr_affine(ptrs, rot.ctypes.data, cs.ctypes.data, count, k, cm.ctypes.data if use_const else None, im.ctypes.data if use_id else None,
                            rw.ctypes.data, vs.ctypes.data, len(rows), out.ctypes.data)
        want = affine_table_python(k, polys, coeffs, rotations, constant, id_coeff, rows, values)
        assert out.tobytes() == ref.mont_rows(want).tobytes()


# ------------------------------------------------------------------------------------------------ lookups
def _lookup_info(hp, k, preprocess, cycles, num_instances):
    pi, q_l, q_r, q_m, q_o, q_c, q_lookup, t_l, t_r, t_o, w_l, w_r, w_o = (Expression.polynomial(p) for p in range(13))
    return hp.PlonkishCircuitInfo(
        k=k, num_instances=[num_instances], preprocess_polys=preprocess, num_witness_polys=[3], num_challenges=[0],
        constraints=[q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi],
        lookups=[[(q_lookup * w_l, t_l), (q_lookup * w_r, t_r), (q_lookup * w_o, t_o)]], permutations=cycles, max_degree=4)   # util.rs:63-86


def test_compiled_lookup_expression_equals_the_tree_and_fits_the_kernel_limits():
    hp = _hp()
    rng = np.random.default_rng(21)
    k = 4
    info = _lookup_info(hp, k, [np.zeros((16, 4), dtype=np.uint64)] * 9, [[(10, 1)], [(11, 1)], [(12, 1)]], 0)
    _, expression = hp.compose(info)
    challenges = [_fe(rng) for _ in range(3)]
    compiled = compile_expression(expression, challenges)
    leaf = _leaf_values(expression, rng)
    assert compiled.value(leaf) == expression.evaluate_field(lambda p: leaf(tuple(p)), lambda q: leaf(("poly", q.poly, q.rotation)), challenges)
    assert compiled.common == -1 and compiled.degree == expression.degree() == 5      # the h term carries no eq_xy
    plain = {a.leaf()[1] for a in compiled.atoms if a.is_leaf() and a.leaf()[0] == "poly" and a.leaf()[2] == 0}
    riders = {q.poly for q in hp.pcs_query(expression, 1)} - plain
    assert len(compiled.atoms) + len(riders) <= 48 and len(compiled.terms) <= 32 and max(len(f) for _, f in compiled.terms) <= 8


@pytest.mark.parametrize("k", [3, 4])
def test_integer_lookup_prover_is_accepted_by_the_integer_verifier(oracle, k):
    from plonkish_b200.transcript import Keccak256Transcript

    rng = np.random.default_rng(60 + k)
    ss = [_fe(rng) for _ in range(k)]
    eqs_host = [oracle.fixed_base_msm(oracle.generator(), e) for e in oracle.kzg_eq_scalars(ref.mont_rows(ss))]
    commit = lambda f: oracle.variable_base_msm(ref.mont_rows(f), eqs_host[k])  # noqa: E731
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_with_lookup_circuit(k, rng)
    sigmas = ref.permutation_polys(k, [10, 11, 12], cycles)
    t = Keccak256Transcript()
    state = ref.prove_reference_lookup(commit, ref.oracle_batch_open(oracle, eqs_host, k), k, instances, preprocess, witness, sigmas, t)
    proof = t.into_proof()
    assert sum(state["m"]) == 1 << k and state["m"][0] == 0     # every row looks something up; the repeated value 0 counts at row 1
    affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
    pre_comms = [affine(commit(p)) for p in preprocess]
    perm_comms = [affine(commit(s)) for s in sigmas]
    ref.verify_reference_lookup(oracle.keccak256, ss, k, instances, pre_comms, perm_comms, proof)
    bad = bytearray(proof)
    bad[6 * 64 + 11] ^= 1
    with pytest.raises(AssertionError):
        ref.verify_reference_lookup(oracle.keccak256, ss, k, instances, pre_comms, perm_comms, bytes(bad))


@pytest.fixture(scope="module")
def emul_lookup():
    subprocess.run(["make", "-C", EMUL_DIR], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(EMUL_DIR, "libemul_msm.so"))
    vp, u32, ci = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
    lib.emul_expr_rows.argtypes = [vp, u32, u32, vp, vp, vp, u32, ci, vp]
    lib.emul_lookup_m.argtypes = [vp, vp, u32, vp]
    lib.emul_lookup_m.restype = ci
    lib.emul_lookup_h.argtypes = [vp, vp, vp, vp, u32, vp]
    return lib


def test_emulated_lookup_kernels_match_python_integers(emul_lookup, oracle):
    rng = np.random.default_rng(77)
    for k in (1, 5, 9):
        n = 1 << k
        # an expression on every row: 3 * p0 * p1 + p2 - 7 (times p3)
        polys = [[_fe(rng) for _ in range(n)] for _ in range(4)]
        terms_int = [(3, [0, 1]), (1, [2]), (R - 7, [])]
        terms = [(ref.to_mont(c), idx) for c, idx in terms_int]
        coeffs, offsets, flat = oracle.flatten_terms(terms)
        arrs = [ref.mont_rows(p) for p in polys]
        ptrs = (ctypes.c_void_p * 4)(*[a.ctypes.data for a in arrs])
        for common in (-1, 3):
            out = np.zeros((n, 4), dtype=np.uint64)
            emul_lookup.emul_expr_rows(ptrs, 4, n, coeffs.ctypes.data, offsets.ctypes.data, flat.ctypes.data, len(terms), common, out.ctypes.data)
            want = [((3 * polys[0][b] * polys[1][b] + polys[2][b] - 7) * (polys[3][b] if common >= 0 else 1)) % R for b in range(n)]
            assert out.tobytes() == ref.mont_rows(want).tobytes()
        # multiplicities: a table with repeated values (0 twice, as util.rs:224-230 builds it), inputs drawn from it
        table = [0, 0] + [_fe(rng) for _ in range(n - 2)] if n > 2 else [0] * n
        if n > 4:
            table[n - 1] = table[3]                                    # another repeated value: its last row is n - 1
        inputs = [table[int(i)] for i in rng.integers(0, n, n)]
        index = {v: i for i, v in enumerate(table)}
        want_m = [0] * n
        for v in inputs:
            want_m[index[v]] += 1
        tab_a, in_a = ref.mont_rows(table), ref.mont_rows(inputs)
        m_out = np.zeros((n, 4), dtype=np.uint64)
        assert emul_lookup.emul_lookup_m(in_a.ctypes.data, tab_a.ctypes.data, n, m_out.ctypes.data) == 0
        assert m_out.tobytes() == ref.mont_rows(want_m).tobytes()
        gamma = _fe(rng)
        h_out = np.zeros((n, 4), dtype=np.uint64)
        emul_lookup.emul_lookup_h(in_a.ctypes.data, tab_a.ctypes.data, m_out.ctypes.data, ref.to_mont(gamma).ctypes.data, n, h_out.ctypes.data)
        want_h = ref.lookup_h_python(inputs, table, want_m, gamma)
        assert h_out.tobytes() == ref.mont_rows(want_h).tobytes()
        assert sum(want_h) % R == 0                                     # the argument's identity (prover.rs:245-247)
        # an input outside the table is reported
        bad = list(inputs)
        bad[n // 2] = (max(table) + 1) % R if (max(table) + 1) % R not in index else 12345
        assert emul_lookup.emul_lookup_m(ref.mont_rows(bad).ctypes.data, tab_a.ctypes.data, n, m_out.ctypes.data) == 1
