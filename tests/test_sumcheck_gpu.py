"""GPU parity (-m gpu) for the sum-check rounds: round messages and table folds
against the oracle's restatement of piop/sum_check/classic/eval.rs:101-131 and
poly/multilinear.rs:179-189, bit-exact, plus the verifier's identities at a size the oracle skips."""
import numpy as np
import pytest

from oracle import bigint_ref as br
from test_sumcheck_cpu import _expr_int, _ints, _mont, _random_case

pytestmark = pytest.mark.gpu

R = br.R


@pytest.fixture(scope="module")
def pk():
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import plonkish_b200

    plonkish_b200._lib.lib()
    return plonkish_b200


@pytest.mark.parametrize("num_polys,num_vars,num_terms,max_fac,common", [
    (1, 1, 1, 1, -1), (3, 2, 2, 2, 1), (6, 7, 5, 3, 0), (9, 12, 6, 4, 0), (4, 13, 3, 7, 3), (27, 9, 8, 3, -1), (32, 8, 32, 2, 5),
])
def test_every_round_matches_the_oracle(pk, oracle, num_polys, num_vars, num_terms, max_fac, common):
    from plonkish_b200.sumcheck import SumCheckProver

    polys, terms = _random_case(oracle, num_polys, num_vars, num_terms, max_fac, common, 11 + num_polys)
    resident = [pk.ResidentScalars(p) for p in polys]
    prover = SumCheckProver(resident, terms, common)
    assert prover.degree == max(max(len(i) for _, i in terms) + (1 if common >= 0 else 0), 1)
    cur = polys
    for rnd in range(num_vars):
        assert prover.round_evals().tobytes() == oracle.sumcheck_round(cur, terms, common).tobytes(), rnd
        ch = oracle.random_scalars(1, 500 + rnd)[0]
        prover.fix_var(ch)
        cur = [oracle.fix_var(p, ch) for p in cur]
    assert prover.final_evals().tobytes() == np.stack([p[0] for p in cur]).tobytes()
    prover.free()
    # the resident polynomials are untouched (they are opened afterwards, backend/hyperplonk.rs:287)
    for r, p in zip(resident, polys):
        assert r.to_host().tobytes() == p.tobytes()
        r.release()


def test_vanilla_plonk_zero_check_satisfies_the_verifier(pk, oracle):
    # The shape of HyperPlonk's zero check for vanilla_plonk (backend/hyperplonk.rs:262-277 over the constraint of
    # backend/hyperplonk/util.rs:33-60): eq(x, y) * (q_l*w_l + q_r*w_r + q_m*w_l*w_r + q_o*w_o + q_c) with a
    # satisfying witness, so the sum is 0.  Checked the way the verifier does (classic.rs:168-190, 242-262):
    # msg(0) + msg(1) = claim in every round, and the last claim equals the expression at the final evaluations,
    # which in turn are the multilinear evaluations of the tables at the challenges.
    from plonkish_b200 import sumcheck

    k = 18
    n = 1 << k
    rng = np.random.default_rng(3)
    small = lambda seed: np.random.default_rng(seed).integers(0, 1 << 62, n, dtype=np.int64)
    w_l, w_r = small(1), small(2)
    q_l, q_r, q_m = small(3) % 5, small(4) % 7, small(5) % 3
    w_o = small(6) | 1
    # choose q_c so that every row satisfies the gate with q_o = -1:  q_c = w_o - (q_l w_l + q_r w_r + q_m w_l w_r)
    ints = lambda a: [int(v) for v in a]
    wl, wr, wo, ql, qr, qm = map(ints, (w_l, w_r, w_o, q_l, q_r, q_m))
    qc = [(o - (a * x + b * y + m * x * y)) % R for o, a, b, m, x, y in zip(wo, ql, qr, qm, wl, wr)]
    y = [int.from_bytes(rng.bytes(32), "little") % R for _ in range(k)]
    eq = np.array([1], dtype=object)
    for y_i in y:
        eq = np.concatenate([eq * (1 - y_i) % R, eq * y_i % R])
    tables = [list(eq), ql, qr, qm, [R - 1] * n, qc, wl, wr, wo]  # 0 eq, 1 q_l, 2 q_r, 3 q_m, 4 q_o, 5 q_c, 6 w_l, 7 w_r, 8 w_o
    host = [_mont(t) for t in tables]
    resident = [pk.ResidentScalars(h) for h in host]
    one = _mont([1])[0]
    terms = [(one, [1, 6]), (one, [2, 7]), (one, [3, 6, 7]), (one, [4, 8]), (one, [5])]
    terms_int = [(1, idx) for _, idx in terms]
    seen = []

    def squeeze(msg):
        seen.append(msg)
        return int.from_bytes(rng.bytes(32), "little") % R

    msgs, challenges, evals = sumcheck.prove(resident, terms, 0, squeeze, common=0)
    assert len(msgs) == k and all(len(m) == 5 for m in msgs)  # degree 3 (q_m*w_l*w_r) + 1 (eq): points 0..4
    claim = 0
    for msg, ch in zip(msgs, challenges):
        assert (msg[0] + msg[1]) % R == claim
        claim = sumcheck.interpolate_at(msg, ch)
    assert claim == _expr_int(terms_int, 0, evals)
    # final evaluations = the tables' multilinear extensions at the challenges (independent path: the oracle's
    # quotients fold, pcs/multilinear.rs:72-107, returns f(point))
    point = _mont(challenges)
    for idx in (0, 3, 5, 8):
        _, value = oracle.quotients(host[idx], point)
        assert _ints(value)[0] == evals[idx], idx
    # the first round against the oracle as well
    assert msgs[0][1:] == _ints(oracle.sumcheck_round(host, terms, 0))
    for r in resident:
        r.release()


def test_argument_errors(pk, oracle):
    from plonkish_b200 import _lib
    from plonkish_b200.sumcheck import SumCheckProver

    p = pk.ResidentScalars(oracle.random_scalars(8, 1))
    q = pk.ResidentScalars(oracle.random_scalars(4, 2))
    one = _mont([1])[0]
    with pytest.raises(AssertionError):
        SumCheckProver([p, q], [(one, [0, 1])])  # tables of different sizes (classic.rs:41)
    with pytest.raises(_lib.PlonkishCudaError):
        SumCheckProver([p], [(one, [0, 3])])  # unknown polynomial index
    with pytest.raises(_lib.PlonkishCudaError):
        SumCheckProver([p], [(one, [0] * 9)])  # too many factors
    prover = SumCheckProver([p], [(one, [0, 0])])
    with pytest.raises(_lib.PlonkishCudaError):
        prover.final_evals()  # rounds left (classic.rs:144)
    for rnd in range(3):
        prover.round_evals()
        prover.fix_var(one)
    with pytest.raises(_lib.PlonkishCudaError):
        prover.round_evals()
    prover.free()
    p.release()
    q.release()


@pytest.mark.parametrize("degree_terms", [4, 5])
def test_round_kernel_on_extreme_values(pk, oracle, degree_terms):
    """Non-leading factors walk X = 1..D as unreduced sums below 5r (fe_add_plain in csrc/sumcheck_kernels.cuh): tables of
    0, 1, r - 1, r - 2 push every step and walked value to its bound; degree 6 takes the reduced walk."""
    from plonkish_b200.sumcheck import SumCheckProver
    from test_sumcheck_cpu import _round_int

    rng = np.random.default_rng(degree_terms)
    num_vars, num_polys = 9, 6
    n = 1 << num_vars
    pool = [0, 1, R - 1, R - 2]
    polys = [_mont([pool[int(v)] for v in rng.integers(0, 4, n)]) for _ in range(num_polys)]
    terms = [(_mont([R - 1])[0], list(range(1, 1 + degree_terms))), (_mont([1])[0], [5, 4, 3]), (_mont([R - 2])[0], [2])]
    terms_int = [(R - 1, list(range(1, 1 + degree_terms))), (1, [5, 4, 3]), (R - 2, [2])]
    resident = [pk.ResidentScalars(p) for p in polys]
    prover = SumCheckProver(resident, terms, 0)
    assert prover.degree == degree_terms + 1
    cur = polys
    for rnd in range(3):
        assert _ints(prover.round_evals()) == _round_int([_ints(p) for p in cur], terms_int, 0, degree_terms + 1), rnd
        ch = _mont([R - 1 - rnd])[0]
        prover.fix_var(ch)
        cur = [oracle.fix_var(p, ch) for p in cur]
    prover.free()
    for r in resident:
        r.release()
