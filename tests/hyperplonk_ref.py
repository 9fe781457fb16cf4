"""Test helper: HyperPlonk for vanilla_plonk restated with Python integers — prover AND verifier — independent of the
product's expression compiler and of its GPU kernels.

  circuit            backend/hyperplonk/util.rs:100-169 (rand_vanilla_plonk_circuit; own RNG) and :378-405 (Permutation)
  preprocess         backend/hyperplonk/preprocessor.rs:172-203 (permutation_polys)
  prover             backend/hyperplonk.rs:164-291, prover.rs:252-409, piop/sum_check/classic.rs:208-240 with
                     classic/eval.rs:101-131 (round message = evaluations at 0..degree, evals[0] = sum - evals[1])
  the constraint     written out by hand as preprocessor.rs:216-252 (test compose_vanilla_plonk) spells it:
                     (gate + alpha * l_1 (z - 1) + alpha^2 (z prod(w + beta id + gamma) - z_next prod(w + beta s + gamma))) * eq
  verifier           backend/hyperplonk.rs:293-362, verifier.rs:19-145, piop/sum_check.rs:59-130,
                     pcs/multilinear.rs:236-278 (additive::batch_verify), pcs/multilinear/kzg.rs:315-362 with the pairing
                     equation checked in G1 through the setup's trapdoor: C - v G = sum_i (s_i - x_i) Q_i
Everything is a table walk over 2^k rows: for k <= 8."""
import numpy as np

from batch_open_ref import batch_open_reference, eq_table, fix_var, to_mont
from oracle import bigint_ref as br
from test_permutation_cpu import bh_iter, z_polys_python

R = br.R
PRIMITIVES = [1, 3, 7, 11, 19, 37, 67, 131, 285, 529, 1033, 2053, 4179, 8219, 16427, 32771, 65581, 131081, 262183, 524327, 1048585, 2097157, 4194307,
              8388641, 16777243, 33554441, 67108935, 134217767, 268435465, 536870917, 1073741907, 2147483657]  # bh.rs:5-37
DEGREE = 5            # 1 (eq) + 1 (z) + 3 (the permutation chunk)


# ------------------------------------------------------------------------------------------------ circuit
class Permutation:
    """util.rs:378-405."""

    def __init__(self):
        self.cycles, self.cycle_idx = [], {}

    def copy(self, lhs, rhs):
        if lhs in self.cycle_idx:
            idx = self.cycle_idx[lhs]
            self.cycles[idx].add(rhs)
            self.cycle_idx[rhs] = idx
        else:
            self.cycles.append({lhs, rhs})
            self.cycle_idx[lhs] = self.cycle_idx[rhs] = len(self.cycles) - 1

    def into_cycles(self):
        return [sorted(c) for c in self.cycles]


def rand_vanilla_plonk_circuit(k, rng):
    """util.rs:100-169 with numpy's generator in place of the two rngs: returns (instances, [q_l, q_r, q_m, q_o, q_c],
    [w_l, w_r, w_o], cycles), canonical integers; the last row stays zero; every gate is satisfied and every copy holds."""
    size = 1 << k
    fe = lambda: int.from_bytes(rng.bytes(40), "little") % R  # noqa: E731
    polys = [[0] * size for _ in range(9)]
    instances = [fe() for _ in range(k)]
    order = bh_iter(k)
    for i, v in enumerate(instances):                       # instance_polys, prover.rs:32-48
        polys[0][order[i + 1]] = v
    perm = Permutation()
    for p in (6, 7, 8):
        perm.copy((p, 1), (p, 1))
    for idx in range(size - 1):
        if rng.integers(2) == 0 and idx > 1:
            l_copy, r_copy = [(int(rng.integers(6, 9)), int(rng.integers(1, idx))) for _ in range(2)]
            perm.copy(l_copy, (6, idx))
            perm.copy(r_copy, (7, idx))
            w_l, w_r = polys[l_copy[0]][l_copy[1]], polys[r_copy[0]][r_copy[1]]
        else:
            w_l, w_r = fe(), fe()
        q_c = fe()
        if rng.integers(2) == 0:
            values = [(1, 1), (2, 1), (4, R - 1), (5, q_c), (6, w_l), (7, w_r), (8, (w_l + w_r + q_c + polys[0][idx]) % R)]
        else:
            values = [(3, 1), (4, R - 1), (5, q_c), (6, w_l), (7, w_r), (8, (w_l * w_r + q_c + polys[0][idx]) % R)]
        for p, v in values:
            polys[p][idx] = v
    return instances, polys[1:6], polys[6:9], perm.into_cycles()


def permutation_polys(k, perm_polys, cycles):
    """preprocessor.rs:172-203."""
    index = {p: i for i, p in enumerate(perm_polys)}
    perms = [[(i << k) + j for j in range(1 << k)] for i in range(len(perm_polys))]
    for cycle in cycles:
        i0, j0 = cycle[0]
        last = perms[index[i0]][j0]
        for i, j in cycle[1:] + cycle[:1]:
            assert j != 0
            perms[index[i]][j], last = last, perms[index[i]][j]
    return perms


# ------------------------------------------------------------------------------------------------ pieces shared by both sides
def bh_next(b, k):
    b <<= 1
    return b ^ ((b >> k) * PRIMITIVES[k])


def bh_prefix(k, count):
    """The first `count` rows of BooleanHypercube::iter (bh.rs:123-130): all the verifier needs (instances, l_1)."""
    out, b = [0], 1
    while len(out) < count:
        out.append(b)
        b = bh_next(b, k)
    return out


def constraint(v, k, beta, gamma, alpha):
    """One row (or one point) of the zero-check polynomial.  v: dict of the values of pi, q_l.., w_.., s_1..3, z, z_next, id,
    l_1, eq."""
    gate = (v["q_l"] * v["w_l"] + v["q_r"] * v["w_r"] + v["q_m"] * v["w_l"] * v["w_r"] + v["q_o"] * v["w_o"] + v["q_c"] + v["pi"]) % R
    first = v["l_1"] * (v["z"] - 1) % R
    lhs, rhs = v["z"], v["z_next"]
    for j, w in enumerate(("w_l", "w_r", "w_o")):
        lhs = lhs * (v[w] + beta * ((j << k) + v["id"]) + gamma) % R
        rhs = rhs * (v[w] + beta * v[f"s_{j + 1}"] + gamma) % R
    return (gate + alpha * first + alpha * alpha % R * (lhs - rhs)) % R * v["eq"] % R


def interpolate(evals, x):
    """Evaluations::evaluate (eval.rs:50-52, barycentric over the points 0..degree): Lagrange form, exact."""
    d = len(evals) - 1
    total = 0
    for j, e in enumerate(evals):
        num = den = 1
        for i in range(d + 1):
            if i != j:
                num = num * (x - i) % R
                den = den * (j - i) % R
        total = (total + e * num % R * pow(den, -1, R)) % R
    return total


def evaluate_multilinear(table, point):
    for x in point:
        table = fix_var(table, x)
    return table[0]


def rotation_eval_points_next(x, k):
    """rotation_eval_points for Rotation::next (poly/multilinear.rs:504-523 with the pattern of :526-545): distance 1,
    pattern = [0, primitive]: the points (0, x_0, .., x_{k-2}) and (bit 0 of primitive, x_0 or 1 - x_0 by bit 1, ...)."""
    pattern = [0, PRIMITIVES[k]]
    return [[pat & 1] + [((1 - x[i]) % R if (pat >> (i + 1)) & 1 else x[i]) for i in range(k - 1)] for pat in pattern]


NAMES = ["pi", "q_l", "q_r", "q_m", "q_o", "q_c", "w_l", "w_r", "w_o", "s_1", "s_2", "s_3", "z"]


# ------------------------------------------------------------------------------------------------ prover
def prove_reference(commit, batch_open, k, instances, preprocess, witness, sigmas, transcript):
    """hyperplonk.rs:164-291.  commit(list of ints) -> affine limbs; batch_open(polys, points, evals, transcript)."""
    n = 1 << k
    order = bh_iter(k)
    for v in instances:
        transcript.common_field_element(v)
    pi = [0] * n
    for i, v in enumerate(instances):
        pi[order[i + 1]] = v
    transcript.write_commitments([commit(w) for w in witness])
    beta = transcript.squeeze_challenge()
    gamma = transcript.squeeze_challenge()
    ((z,), _, _) = z_polys_python(1, witness, sigmas, beta, gamma, k)
    transcript.write_commitments([commit(z)])
    alpha = transcript.squeeze_challenge()
    y = transcript.squeeze_challenges(k)
    polys = [pi] + list(preprocess) + list(witness) + list(sigmas) + [z]
    tabs = dict(zip(NAMES, polys))
    tabs["z_next"] = [z[bh_next(b, k)] for b in range(n)]           # rotation_map, classic.rs:105-125
    tabs["id"] = list(range(n))
    tabs["l_1"] = [1 if b == order[1] else 0 for b in range(n)]      # classic.rs:44-55
    tabs["eq"] = eq_table(y)
    claim, x = 0, []
    for _ in range(k):
        size = len(tabs["eq"]) // 2
        msg = [0] * (DEGREE + 1)
        for t in range(1, DEGREE + 1):
            for b in range(size):
                v = {name: (tab[2 * b] + t * (tab[2 * b + 1] - tab[2 * b])) % R for name, tab in tabs.items()}
                msg[t] = (msg[t] + constraint(v, k, beta, gamma, alpha)) % R
        msg[0] = (claim - msg[1]) % R                                # eval.rs:128
        transcript.write_field_elements(msg)
        ch = transcript.squeeze_challenge()
        x.append(ch)
        claim = interpolate(msg, ch)
        tabs = {name: fix_var(tab, ch) for name, tab in tabs.items()}
    # prover.rs:388-408: queries ordered by (poly, rotation); points = [x] + the two points of Rotation::next
    pts = [x] + rotation_eval_points_next(x, k)
    evals = [(i, 0, tabs[NAMES[i]][0]) for i in range(1, 13)] + [(12, 1 + j, evaluate_multilinear(z, pt)) for j, pt in enumerate(pts[1:])]
    transcript.write_field_elements([v for _, _, v in evals])
    batch_open(polys, pts, evals, transcript)
    return {"beta": beta, "gamma": gamma, "alpha": alpha, "y": y, "x": x, "evals": evals, "z": z}


# ------------------------------------------------------------------------------------------------ verifier
class ProofReader:
    """The read side of Keccak256Transcript (util/transcript.rs:133-166, 183-214) over any keccak256(bytes) function."""

    def __init__(self, keccak256, proof: bytes):
        self.h, self.absorbed, self.proof, self.pos = keccak256, b"", proof, 0

    def common_field_element(self, v):
        self.absorbed += (v % R).to_bytes(32, "little")

    def read_field_element(self):
        v = int.from_bytes(self.proof[self.pos: self.pos + 32], "big")
        self.pos += 32
        assert v < R, "Invalid field element encoding in proof"
        self.common_field_element(v)
        return v

    def read_field_elements(self, n):
        return [self.read_field_element() for _ in range(n)]

    def read_commitment(self):
        x = int.from_bytes(self.proof[self.pos: self.pos + 32], "big")
        y = int.from_bytes(self.proof[self.pos + 32: self.pos + 64], "big")
        self.pos += 64
        assert br.is_on_curve((x, y)), "Invalid elliptic curve point encoding in proof"
        self.absorbed += x.to_bytes(32, "little") + y.to_bytes(32, "little")
        return (x, y)

    def read_commitments(self, n):
        return [self.read_commitment() for _ in range(n)]

    def squeeze_challenge(self):
        h = self.h(self.absorbed)
        self.absorbed = h
        return int.from_bytes(h, "little") % R

    def squeeze_challenges(self, n):
        return [self.squeeze_challenge() for _ in range(n)]


def lagrange_eval(x, b):
    out = 1
    for i, x_i in enumerate(x):
        out = out * (x_i if (b >> i) & 1 else 1 - x_i) % R
    return out


def eq_xy_eval(x, y):
    out = 1
    for a, b in zip(x, y):
        out = out * ((2 * a * b + 1 - a - b) % R) % R
    return out


def verify_reference(keccak256, ss, k, instances, preprocess_comms, permutation_comms, proof: bytes, pcs_verify=None) -> None:
    """hyperplonk.rs:293-362; raises AssertionError where the reference returns Err.  Commitments are affine integer
    pairs; ss: the setup's trapdoor (canonical integers).  pcs_verify(reader, g_prime_comm, point, g_prime_eval): another
    additive PCS's verify for the last step (Zeromorph: zeromorph_ref.verify_reader_in_g1); default MultilinearKzg::verify."""
    t = ProofReader(keccak256, proof)
    order = bh_prefix(k, max(len(instances), 1) + 1)
    for v in instances:
        t.common_field_element(v)
    witness_comms = t.read_commitments(3)
    beta = t.squeeze_challenge()
    gamma = t.squeeze_challenge()
    z_comms = t.read_commitments(1)
    alpha = t.squeeze_challenge()
    y = t.squeeze_challenges(k)
    # ClassicSumCheck::verify (classic.rs:242-262) + verify_consistency (:175-194)
    claim, x = 0, []
    for rnd in range(k):
        msg = t.read_field_elements(DEGREE + 1)
        assert (msg[0] + msg[1]) % R == claim, f"sum check: consistency failure at round {rnd}"
        ch = t.squeeze_challenge()
        x.append(ch)
        claim = interpolate(msg, ch)
    # verifier.rs:56-78: the evaluations, the rotated one folded by rotation_eval (poly/multilinear.rs:433-476, distance 1)
    flat = t.read_field_elements(12 + 2)
    v = {NAMES[i]: flat[i - 1] for i in range(1, 13)}
    e0, e1 = flat[12], flat[13]
    v["z_next"] = ((e1 - e0) * x[k - 1] + e0) % R
    v["pi"] = sum(inst * lagrange_eval(x, order[i + 1]) for i, inst in enumerate(instances)) % R   # instance_evals, verifier.rs:92-145
    v["id"] = sum(x_i << i for i, x_i in enumerate(x)) % R                                         # identity_eval
    v["l_1"] = lagrange_eval(x, order[1])
    v["eq"] = eq_xy_eval(x, y)
    assert constraint(v, k, beta, gamma, alpha) == claim, "Unmatched between sum_check output and query evaluation"
    pts = [x] + rotation_eval_points_next(x, k)
    evals = [(i, 0, flat[i - 1]) for i in range(1, 13)] + [(12, 1, e0), (12, 2, e1)]
    comms = [None] + list(preprocess_comms) + witness_comms + list(permutation_comms) + z_comms
    # additive::batch_verify (pcs/multilinear.rs:236-278)
    ell = max(len(evals) - 1, 0).bit_length()
    tt = t.squeeze_challenges(ell)
    eq_xt = eq_table(tt)
    claim2 = sum(val * w for (_, _, val), w in zip(evals, eq_xt)) % R
    ch2 = []
    for rnd in range(k):
        c = t.read_field_elements(3)                                   # coefficient form, coeff.rs:25-47
        assert (2 * c[0] + c[1] + c[2]) % R == claim2, f"batch_verify sum check: consistency failure at round {rnd}"
        r_ = t.squeeze_challenge()
        ch2.append(r_)
        claim2 = (c[0] + r_ * (c[1] + r_ * c[2])) % R
    g_prime_eval = claim2
    eq_evals = [eq_xy_eval(ch2, pt) for pt in pts]
    scalars = [eq_evals[pt] * w % R for (_, pt, _), w in zip(evals, eq_xt)]
    g_prime_comm = br.msm(scalars, [comms[p] for p, _, _ in evals])
    if pcs_verify is not None:
        pcs_verify(t, g_prime_comm, ch2, g_prime_eval)
        assert t.pos == len(proof), "trailing bytes in the proof"
        return
    # MultilinearKzg::verify (kzg.rs:315-362): e(C - v g1, g2) = prod e(Q_i, s_i g2 - x_i g2), here in G1 with the trapdoor
    quotients = t.read_commitments(k)
    lhs = br.add(g_prime_comm, br.neg(br.scalar_mul(g_prime_eval, br.G)))
    rhs = br.msm([(s - x_i) % R for s, x_i in zip(ss, ch2)], quotients)
    assert lhs == rhs, "Invalid multilinear KZG opening"
    assert t.pos == len(proof), "trailing bytes in the proof"


def oracle_batch_open(oracle, eqs_host, k):
    return lambda polys, points, evals, transcript: batch_open_reference(oracle, eqs_host, k, polys, points, evals, transcript)


def mont_rows(values):
    return np.stack([to_mont(v) for v in values])


# =================================================================================================
# vanilla_plonk_with_lookup (backend/hyperplonk/util.rs:63-98, 216-330): the same restatement with the lookup argument
#   lookup_compressed_polys / lookup_m_polys / lookup_h_polys      backend/hyperplonk/prover.rs:50-250
#   the constraint as preprocessor.rs:254-303 (compose_vanilla_plonk_with_lookup) spells it:
#     h + alpha * (gate + alpha * (h (inp + gamma)(tab + gamma) - (tab + gamma) + m (inp + gamma))
#                  + alpha^2 * l_1 (z - 1) + alpha^3 * (z prod(..id..) - z_next prod(..s..))) * eq
#   with inp = q_lookup w_l + beta q_lookup w_r + beta^2 q_lookup w_o, tab = t_l + beta t_r + beta^2 t_o
NAMES_LOOKUP = ["pi", "q_l", "q_r", "q_m", "q_o", "q_c", "q_lookup", "t_l", "t_r", "t_o", "w_l", "w_r", "w_o", "s_1", "s_2", "s_3", "m", "h", "z"]


def rand_vanilla_plonk_with_lookup_circuit(k, rng):
    """util.rs:216-330 with numpy's generator: returns (instances, the 9 preprocessed columns, [w_l, w_r, w_o], cycles)."""
    size = 1 << k
    fe = lambda: int.from_bytes(rng.bytes(40), "little") % R  # noqa: E731
    polys = [[0] * size for _ in range(13)]
    for p in (7, 8, 9):
        polys[p] = [0, 0] + [fe() for _ in range(size - 2)]
    instances = [fe() for _ in range(k)]
    order = bh_iter(k)
    for i, v in enumerate(instances):
        polys[0][order[i + 1]] = v
    instance_rows = set(order[: k + 1])
    perm = Permutation()
    for p in (10, 11, 12):
        perm.copy((p, 1), (p, 1))
    for idx in range(size - 1):
        use_copy = rng.integers(2) == 0 and idx > 1
        if use_copy:
            l_copy, r_copy = [(int(rng.integers(10, 13)), int(rng.integers(1, idx))) for _ in range(2)]
            perm.copy(l_copy, (10, idx))
            perm.copy(r_copy, (11, idx))
            w_l, w_r = polys[l_copy[0]][l_copy[1]], polys[r_copy[0]][r_copy[1]]
        else:
            w_l, w_r = fe(), fe()
        q_c = fe()
        arithmetic, add = (use_copy or idx in instance_rows), rng.integers(2) == 0
        if arithmetic and add:
            values = [(1, 1), (2, 1), (4, R - 1), (5, q_c), (10, w_l), (11, w_r), (12, (w_l + w_r + q_c + polys[0][idx]) % R)]
        elif arithmetic:
            values = [(3, 1), (4, R - 1), (5, q_c), (10, w_l), (11, w_r), (12, (w_l * w_r + q_c + polys[0][idx]) % R)]
        else:
            row = int(rng.integers(1, size))
            values = [(6, 1), (10, polys[7][row]), (11, polys[8][row]), (12, polys[9][row])]
        for p, v in values:
            polys[p][idx] = v
    return instances, polys[1:10], polys[10:13], perm.into_cycles()


def lookup_polys_python(q_lookup, tables, witness, beta, gamma):
    """prover.rs:50-250 on integers: (compressed input, compressed table, m, h); raises on an input outside the table."""
    n = len(q_lookup)
    inp = [(q_lookup[b] * witness[0][b] + beta * q_lookup[b] * witness[1][b] + beta * beta * q_lookup[b] * witness[2][b]) % R for b in range(n)]
    tab = [(tables[0][b] + beta * tables[1][b] + beta * beta * tables[2][b]) % R for b in range(n)]
    index = {v: i for i, v in enumerate(tab)}                        # the last row of a repeated value (HashMap, :151)
    m = [0] * n
    for v in inp:
        if v not in index:
            raise ValueError("Invalid lookup input")
        m[index[v]] += 1
    return inp, tab, m


def lookup_h_python(inp, tab, m, gamma):
    return [(pow((gamma + i) % R, -1, R) - m_ * pow((gamma + t) % R, -1, R)) % R for i, t, m_ in zip(inp, tab, m)]


def constraint_lookup(v, k, beta, gamma, alpha):
    gate = (v["q_l"] * v["w_l"] + v["q_r"] * v["w_r"] + v["q_m"] * v["w_l"] * v["w_r"] + v["q_o"] * v["w_o"] + v["q_c"] + v["pi"]) % R
    inp = (v["q_lookup"] * v["w_l"] + beta * v["q_lookup"] * v["w_r"] + beta * beta * v["q_lookup"] * v["w_o"]) % R
    tab = (v["t_l"] + beta * v["t_r"] + beta * beta * v["t_o"]) % R
    lookup = (v["h"] * (inp + gamma) * (tab + gamma) - (tab + gamma) + v["m"] * (inp + gamma)) % R
    first = v["l_1"] * (v["z"] - 1) % R
    lhs, rhs = v["z"], v["z_next"]
    for j, w in enumerate(("w_l", "w_r", "w_o")):
        lhs = lhs * (v[w] + beta * ((j << k) + v["id"]) + gamma) % R
        rhs = rhs * (v[w] + beta * v[f"s_{j + 1}"] + gamma) % R
    every_row = (gate + alpha * lookup + pow(alpha, 2, R) * first + pow(alpha, 3, R) * (lhs - rhs)) % R * v["eq"] % R
    return (v["h"] + alpha * every_row) % R


def prove_reference_lookup(commit, batch_open, k, instances, preprocess, witness, sigmas, transcript):
    """hyperplonk.rs:164-291 for vanilla_plonk_with_lookup."""
    n = 1 << k
    order = bh_iter(k)
    for v in instances:
        transcript.common_field_element(v)
    pi = [0] * n
    for i, v in enumerate(instances):
        pi[order[i + 1]] = v
    transcript.write_commitments([commit(w) for w in witness])
    beta = transcript.squeeze_challenge()
    inp, tab, m = lookup_polys_python(preprocess[5], preprocess[6:9], witness, beta, None)
    transcript.write_commitments([commit(m)])
    gamma = transcript.squeeze_challenge()
    h = lookup_h_python(inp, tab, m, gamma)
    assert sum(h) % R == 0                                          # sanity-check, prover.rs:245-247
    ((z,), _, _) = z_polys_python(1, witness, sigmas, beta, gamma, k)
    transcript.write_commitments([commit(h), commit(z)])
    alpha = transcript.squeeze_challenge()
    y = transcript.squeeze_challenges(k)
    polys = [pi] + list(preprocess) + list(witness) + list(sigmas) + [m, h, z]
    tabs = dict(zip(NAMES_LOOKUP, polys))
    tabs["z_next"] = [z[bh_next(b, k)] for b in range(n)]
    tabs["id"] = list(range(n))
    tabs["l_1"] = [1 if b == order[1] else 0 for b in range(n)]
    tabs["eq"] = eq_table(y)
    claim, x = 0, []
    for _ in range(k):
        size = len(tabs["eq"]) // 2
        msg = [0] * (DEGREE + 1)
        for t in range(1, DEGREE + 1):
            for b in range(size):
                v = {name: (tb[2 * b] + t * (tb[2 * b + 1] - tb[2 * b])) % R for name, tb in tabs.items()}
                msg[t] = (msg[t] + constraint_lookup(v, k, beta, gamma, alpha)) % R
        msg[0] = (claim - msg[1]) % R
        transcript.write_field_elements(msg)
        ch = transcript.squeeze_challenge()
        x.append(ch)
        claim = interpolate(msg, ch)
        tabs = {name: fix_var(tb, ch) for name, tb in tabs.items()}
    pts = [x] + rotation_eval_points_next(x, k)
    evals = [(i, 0, tabs[NAMES_LOOKUP[i]][0]) for i in range(1, 19)] + [(18, 1 + j, evaluate_multilinear(z, pt)) for j, pt in enumerate(pts[1:])]
    transcript.write_field_elements([v for _, _, v in evals])
    batch_open(polys, pts, evals, transcript)
    return {"m": m, "h": h, "z": z, "input": inp, "table": tab}


def verify_reference_lookup(keccak256, ss, k, instances, preprocess_comms, permutation_comms, proof: bytes) -> None:
    """hyperplonk.rs:293-362 for vanilla_plonk_with_lookup (one lookup, one permutation z polynomial)."""
    t = ProofReader(keccak256, proof)
    order = bh_prefix(k, max(len(instances), 1) + 1)
    for v in instances:
        t.common_field_element(v)
    witness_comms = t.read_commitments(3)
    beta = t.squeeze_challenge()
    m_comms = t.read_commitments(1)
    gamma = t.squeeze_challenge()
    h_z_comms = t.read_commitments(2)
    alpha = t.squeeze_challenge()
    y = t.squeeze_challenges(k)
    claim, x = 0, []
    for rnd in range(k):
        msg = t.read_field_elements(DEGREE + 1)
        assert (msg[0] + msg[1]) % R == claim, f"sum check: consistency failure at round {rnd}"
        ch = t.squeeze_challenge()
        x.append(ch)
        claim = interpolate(msg, ch)
    flat = t.read_field_elements(18 + 2)
    v = {NAMES_LOOKUP[i]: flat[i - 1] for i in range(1, 19)}
    e0, e1 = flat[18], flat[19]
    v["z_next"] = ((e1 - e0) * x[k - 1] + e0) % R
    v["pi"] = sum(inst * lagrange_eval(x, order[i + 1]) for i, inst in enumerate(instances)) % R
    v["id"] = sum(x_i << i for i, x_i in enumerate(x)) % R
    v["l_1"] = lagrange_eval(x, order[1])
    v["eq"] = eq_xy_eval(x, y)
    assert constraint_lookup(v, k, beta, gamma, alpha) == claim, "Unmatched between sum_check output and query evaluation"
    pts = [x] + rotation_eval_points_next(x, k)
    evals = [(i, 0, flat[i - 1]) for i in range(1, 19)] + [(18, 1, e0), (18, 2, e1)]
    comms = [None] + list(preprocess_comms) + witness_comms + list(permutation_comms) + m_comms + h_z_comms
    ell = max(len(evals) - 1, 0).bit_length()
    tt = t.squeeze_challenges(ell)
    eq_xt = eq_table(tt)
    claim2 = sum(val * w for (_, _, val), w in zip(evals, eq_xt)) % R
    ch2 = []
    for rnd in range(k):
        c = t.read_field_elements(3)
        assert (2 * c[0] + c[1] + c[2]) % R == claim2, f"batch_verify sum check: consistency failure at round {rnd}"
        r_ = t.squeeze_challenge()
        ch2.append(r_)
        claim2 = (c[0] + r_ * (c[1] + r_ * c[2])) % R
    eq_evals = [eq_xy_eval(ch2, pt) for pt in pts]
    scalars = [eq_evals[pt] * w % R for (_, pt, _), w in zip(evals, eq_xt)]
    g_prime_comm = br.msm(scalars, [comms[p] for p, _, _ in evals])
    quotients = t.read_commitments(k)
    lhs = br.add(g_prime_comm, br.neg(br.scalar_mul(claim2, br.G)))
    rhs = br.msm([(s - x_i) % R for s, x_i in zip(ss, ch2)], quotients)
    assert lhs == rhs, "Invalid multilinear KZG opening"
    assert t.pos == len(proof), "trailing bytes in the proof"
