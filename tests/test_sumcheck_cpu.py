"""CPU suite for the sum-check rounds: the oracle's restatement of
piop/sum_check/classic/eval.rs:101-131 and poly/multilinear.rs:179-189 pinned against Python integers
and the protocol's own identities, and the product's kernels (sumcheck_kernels.cuh) run through the CPU
emulation build against the oracle.  No GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import bigint_ref as br

R = br.R
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


def _ints(limbs):
    rinv = pow(br.MONT, -1, R)
    arr = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
    return [int.from_bytes(row.tobytes(), "little") * rinv % R for row in arr]


def _mont(vals):
    return np.frombuffer(b"".join((int(v) % R * br.MONT % R).to_bytes(32, "little") for v in vals), dtype=np.uint64).reshape(-1, 4).copy()


def _expr_int(terms_int, common, vals):
    total = 0
    for c, idx in terms_int:
        prod = c
        for i in idx:
            prod = prod * vals[i] % R
        total = (total + prod) % R
    return total * vals[common] % R if common >= 0 else total


def _round_int(polys_int, terms_int, common, degree):
    """eval.rs:101-131 on integers: sum over pairs (2b, 2b+1) of expr at X = 1..degree."""
    n = len(polys_int[0])
    out = [0] * degree
    for b in range(n // 2):
        for x in range(1, degree + 1):
            vals = [(p[2 * b] + x * (p[2 * b + 1] - p[2 * b])) % R for p in polys_int]
            out[x - 1] = (out[x - 1] + _expr_int(terms_int, common, vals)) % R
    return out


def _random_case(oracle, num_polys, num_vars, num_terms, max_fac, common, seed):
    rng = np.random.default_rng(seed)
    polys = [oracle.random_scalars(1 << num_vars, seed * 100 + p) for p in range(num_polys)]
    coeffs = oracle.random_scalars(num_terms, seed * 100 + 99)
    one = _mont([1])[0]
    terms = []
    for t in range(num_terms):
        nf = int(rng.integers(0 if t == num_terms - 1 else 1, max_fac + 1))
        idx = [int(i) for i in rng.integers(0, num_polys, nf)]
        terms.append((one if t % 3 == 0 else coeffs[t], idx))
    return polys, terms


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", EMUL_DIR], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(EMUL_DIR, "libemul_msm.so"))
    vp, u32, ci = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
    lib.emul_sumcheck_round.argtypes = [vp, u32, u32, vp, vp, vp, u32, ci, u32, u32, vp]
    lib.emul_sumcheck_fold.argtypes = [vp, vp, u32, u32, vp, u32]
    lib.emul_sumcheck_round_factored.argtypes = [vp, u32, u32, vp, vp, vp, u32, ci, u32, u32, vp]
    return lib


def test_oracle_round_and_fix_var_match_python_integers(oracle):
    for num_polys, num_vars, num_terms, max_fac, common, seed in [(3, 1, 2, 2, -1, 1), (5, 4, 4, 3, 0, 2), (4, 3, 3, 4, 2, 3)]:
        polys, terms = _random_case(oracle, num_polys, num_vars, num_terms, max_fac, common, seed)
        polys_int = [_ints(p) for p in polys]
        terms_int = [(_ints(c)[0], idx) for c, idx in terms]
        got = _ints(oracle.sumcheck_round(polys, terms, common))
        degree = max(max(len(i) for _, i in terms) + (1 if common >= 0 else 0), 1)
        assert got == _round_int(polys_int, terms_int, common, degree)
        x = oracle.random_scalars(1, seed + 50)[0]
        xi = _ints(x)[0]
        for p, pi in zip(polys, polys_int):
            want = [(pi[2 * b] + (pi[2 * b + 1] - pi[2 * b]) * xi) % R for b in range(len(pi) // 2)]  # multilinear.rs:615
            assert _ints(oracle.fix_var(p, x)) == want


def test_oracle_protocol_identities(oracle):
    # A zero check the way HyperPlonk uses it (backend/hyperplonk.rs:262-277): eq(x, y) * (a*b - c) with c = a o b sums
    # to 0; every round message satisfies msg(0) + msg(1) = running claim and the last claim equals the
    # expression at the final evaluations (the verifier's checks, classic.rs:168-190, 242-262).
    from plonkish_b200.sumcheck import interpolate_at

    k = 6
    n = 1 << k
    a, b = oracle.random_scalars(n, 1), oracle.random_scalars(n, 2)
    ai, bi = _ints(a), _ints(b)
    c = _mont([x * y % R for x, y in zip(ai, bi)])
    y = _ints(oracle.random_scalars(k, 3))
    eq = [1]
    for y_i in y:  # eq(x, y) table, lowest variable first
        eq = [e * (1 - y_i) % R for e in eq] + [e * y_i % R for e in eq]
    polys = [_mont(eq), a, b, c]
    terms = [(_mont([1])[0], [1, 2]), (_mont([R - 1])[0], [3])]
    terms_int = [(1, [1, 2]), (R - 1, [3])]
    claim = 0
    rng = np.random.default_rng(9)
    for rnd in range(k):
        tail = _ints(oracle.sumcheck_round(polys, terms, 0))
        msg = [(claim - tail[0]) % R] + tail
        # evals[0] derived from the claim must equal the direct sum at X = 0
        direct0 = sum(_expr_int(terms_int, 0, [_ints(p[2 * j: 2 * j + 1])[0] for p in polys]) for j in range(len(polys[0]) // 2)) % R
        assert msg[0] == direct0, rnd
        ch = int.from_bytes(rng.bytes(32), "little") % R
        claim = interpolate_at(msg, ch)
        polys = [oracle.fix_var(p, _mont([ch])[0]) for p in polys]
    finals = [_ints(p)[0] for p in polys]
    assert claim == _expr_int(terms_int, 0, finals)


@pytest.mark.parametrize("num_polys,num_vars,num_terms,max_fac,common,sms", [
    (1, 1, 1, 1, -1, 1), (3, 1, 2, 2, 1, 1), (6, 5, 5, 3, 0, 1), (9, 9, 6, 4, 0, 1), (4, 10, 3, 7, 3, 2), (27, 6, 8, 3, -1, 1), (32, 4, 32, 2, 5, 1),
])
def test_emulated_round_and_fold_kernels(emul, oracle, num_polys, num_vars, num_terms, max_fac, common, sms):
    # sms = 1 or 2 caps the grid at 2 / 4 blocks: the grid-stride loop, the block tree sum and the partial sums all run;
    # 27 and 32 tables force 64-thread blocks (shared-memory fit)
    polys, terms = _random_case(oracle, num_polys, num_vars, num_terms, max_fac, common, 7 + num_polys)
    n = 1 << num_vars
    coeffs, offsets, flat = oracle.flatten_terms(terms)
    degree = max(max(len(i) for _, i in terms) + (1 if common >= 0 else 0), 1)
    ptrs = (ctypes.c_void_p * num_polys)(*[p.ctypes.data for p in polys])
    out = np.zeros((degree, 4), dtype=np.uint64)
    emul.emul_sumcheck_round(ctypes.cast(ptrs, ctypes.c_void_p), num_polys, n, coeffs.ctypes.data, offsets.ctypes.data, flat.ctypes.data,
                             num_terms, common, degree, sms, out.ctypes.data)
    assert out.tobytes() == oracle.sumcheck_round(polys, terms, common).tobytes()
    x = oracle.random_scalars(1, 77)[0]
    outs = [np.zeros((n // 2, 4), dtype=np.uint64) for _ in polys]
    optrs = (ctypes.c_void_p * num_polys)(*[o.ctypes.data for o in outs])
    emul.emul_sumcheck_fold(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(optrs, ctypes.c_void_p), num_polys, n, x.ctypes.data, sms)
    for p, o in zip(polys, outs):
        assert o.tobytes() == oracle.fix_var(p, x).tobytes()


@pytest.mark.parametrize("degree_terms", [4, 5])
def test_emulated_round_kernel_on_extreme_values(emul, oracle, degree_terms):
    """The walk X = 1..D of a non-leading factor runs on unreduced sums (fe_add_plain: e[2b+1] + x * step < 5r): tables of
    0, 1, r - 1, r - 2 make every step and every walked value as large as they get."""
    rng = np.random.default_rng(degree_terms)
    num_vars, num_polys = 7, 6
    n = 1 << num_vars
    pool = [0, 1, R - 1, R - 2]
    polys = [_mont([pool[int(v)] for v in rng.integers(0, 4, n)]) for _ in range(num_polys)]
    terms = [(_mont([R - 1])[0], list(range(1, 1 + degree_terms))), (_mont([1])[0], [5, 4, 3]), (_mont([R - 2])[0], [2])]
    coeffs, offsets, flat = oracle.flatten_terms(terms)
    degree = degree_terms + 1
    ptrs = (ctypes.c_void_p * num_polys)(*[p.ctypes.data for p in polys])
    out = np.zeros((degree, 4), dtype=np.uint64)
    emul.emul_sumcheck_round(ctypes.cast(ptrs, ctypes.c_void_p), num_polys, n, coeffs.ctypes.data, offsets.ctypes.data, flat.ctypes.data,
                             len(terms), 0, degree, 1, out.ctypes.data)
    terms_int = [(R - 1, list(range(1, 1 + degree_terms))), (1, [5, 4, 3]), (R - 2, [2])]
    assert _ints(out) == _round_int([_ints(p) for p in polys], terms_int, 0, degree)


@pytest.mark.parametrize("k,max_fac,num_terms,sms", [(1, 2, 2, 1), (5, 3, 4, 1), (8, 4, 5, 2), (6, 6, 3, 1)])
def test_factored_zero_check_rounds_give_the_same_messages(emul, oracle, k, max_fac, num_terms, sms):
    """The factored round (common factor eq(x, y) entering as its pair sum, one evaluation point fewer) plus
    sumcheck.zero_check_message rebuilds, round after round, exactly the message of the plain round
    (eval.rs:101-131) — including after folds, when the eq table carries the product of the earlier eq factors."""
    from plonkish_b200.sumcheck import interpolate_at, zero_check_message

    n = 1 << k
    num_polys = 5
    y = _ints(oracle.random_scalars(k, 31 + k))
    eq = [1]
    for y_i in y:  # eq(x, y), lowest variable first (poly/multilinear.rs:91-130)
        eq = [e * (1 - y_i) % R for e in eq] + [e * y_i % R for e in eq]
    polys, terms = _random_case(oracle, num_polys, k, num_terms, max_fac, 0, 40 + k)
    polys[0] = _mont(eq)
    coeffs, offsets, flat = oracle.flatten_terms(terms)
    degree = max(len(i) for _, i in terms) + 1
    rng = np.random.default_rng(k)
    claim = None
    for rnd in range(k):
        m = len(polys[0])
        ptrs = (ctypes.c_void_p * num_polys)(*[p.ctypes.data for p in polys])
        tail = _ints(oracle.sumcheck_round(polys, terms, 0))
        if claim is None:  # h(0) + h(1) of the first round, directly
            terms_int = [(_ints(c)[0], idx) for c, idx in terms]
            claim = sum(_expr_int(terms_int, 0, [_ints(p[j: j + 1])[0] for p in polys]) for j in range(m)) % R
        want = [(claim - tail[0]) % R] + tail
        g = np.zeros((degree - 1, 4), dtype=np.uint64)
        emul.emul_sumcheck_round_factored(ctypes.cast(ptrs, ctypes.c_void_p), num_polys, m, coeffs.ctypes.data, offsets.ctypes.data, flat.ctypes.data,
                                          len(terms), 0, degree, sms, g.ctypes.data)
        assert zero_check_message(_ints(g), claim, y[rnd]) == want, rnd
        ch = int.from_bytes(rng.bytes(32), "little") % R
        claim = interpolate_at(want, ch)
        polys = [oracle.fix_var(p, _mont([ch])[0]) for p in polys]
    assert zero_check_message([5, 7], 11, 1) is None  # 1 - y_r = 0: the caller runs the plain round
