"""Test helper: `additive::batch_open` (pcs/multilinear.rs:134-235) restated with Python integers and the oracle's
MSM — independent of the product mirror (plonkish_b200/kzg.py batch_open): the sum-check messages are computed the
way CoefficientsProver does (piop/sum_check/classic/coeff.rs:132-190: c0 = sum lhs0 * rhs0, c2 = sum (lhs1 - lhs0)
(rhs1 - rhs0), c1 = claimed - 2 c0 - c2), directly from the tables."""
import numpy as np

from oracle import bigint_ref as br

R = br.R
MONT = 1 << 256
RINV = pow(MONT, -1, R)


def to_int(limbs) -> int:
    return int.from_bytes(np.ascontiguousarray(limbs, dtype=np.uint64).tobytes(), "little") * RINV % R


def to_mont(v: int) -> np.ndarray:
    return np.frombuffer((v % R * MONT % R).to_bytes(32, "little"), dtype=np.uint64).copy()


def eq_table(y):
    evals = [1]
    for v in y:
        evals = [e * (1 - v) % R for e in evals] + [e * v % R for e in evals]
    return evals


def fix_var(table, x):
    return [(table[2 * b] + (table[2 * b + 1] - table[2 * b]) * x) % R for b in range(len(table) // 2)]


def batch_open_reference(oracle, eqs_host, num_vars, polys, points, evals, transcript, open_fn=None):
    """polys: lists of canonical integers (2^num_vars each); eqs_host[i]: the SRS slice of 2^i bases ([2^i, 8] limbs).
    open_fn(g_prime, challenges, transcript): another PCS's open for the last step (Zeromorph); default MultilinearKzg::open."""
    ell = max(len(evals) - 1, 0).bit_length()
    t = transcript.squeeze_challenges(ell)
    eq_xt = eq_table(t)
    n = 1 << num_vars
    merged = [(1, None)] * len(points)
    for (poly, point, _), w in zip(evals, eq_xt):              # multilinear.rs:150-167
        if merged[point][1] is None:
            merged[point] = (w, list(polys[poly]))
        else:
            coeff, m = merged[point]
            if coeff != 1:
                m = [coeff * v % R for v in m]
            merged[point] = (1, [(a + w * b) % R for a, b in zip(m, polys[poly])])
    claim = sum(v * w for (_, _, v), w in zip(evals, eq_xt)) % R
    eq_tabs = [eq_table(pt) for pt in points]
    tabs = [m for _, m in merged]
    challenges = []
    for _ in range(num_vars):
        c0 = c2 = 0
        for (scalar, _), lhs, rhs in zip(merged, eq_tabs, tabs):  # coeff.rs:136-143, karatsuba::<true>
            a0 = sum(lhs[2 * b] * rhs[2 * b] for b in range(len(lhs) // 2)) % R
            a2 = sum((lhs[2 * b + 1] - lhs[2 * b]) * (rhs[2 * b + 1] - rhs[2 * b]) for b in range(len(lhs) // 2)) % R
            c0 = (c0 + scalar * a0) % R
            c2 = (c2 + scalar * a2) % R
        c1 = (claim - 2 * c0 - c2) % R
        transcript.write_field_elements([c0, c1, c2])
        ch = transcript.squeeze_challenge()
        challenges.append(ch)
        claim = (c0 + ch * (c1 + ch * c2)) % R
        eq_tabs = [fix_var(tb, ch) for tb in eq_tabs]
        tabs = [fix_var(tb, ch) for tb in tabs]
    g_prime = [0] * n
    for (scalar, m), pt in zip(merged, points):                # multilinear.rs:203-213
        e = 1
        for a, b in zip(challenges, pt):
            e = e * ((a * b + (1 - a) * (1 - b)) % R) % R
        w = scalar * e % R
        g_prime = [(g + w * v) % R for g, v in zip(g_prime, m)]
    if open_fn is not None:
        return challenges, open_fn(g_prime, challenges, transcript)
    # MultilinearKzg::open (kzg.rs:276-302): quotients (multilinear.rs:72-107) and their commitments
    rem = g_prime
    comms = []
    for i in reversed(range(num_vars)):
        half = 1 << i
        lo, hi = rem[:half], rem[half:2 * half]
        q = [(h - l) % R for h, l in zip(hi, lo)]
        comms.append(oracle.variable_base_msm(np.stack([to_mont(v) for v in q]), eqs_host[i]))
        rem = [(l + (h - l) * challenges[i]) % R for h, l in zip(hi, lo)]
    comms.reverse()
    transcript.write_commitments(comms)
    return challenges, rem[0]
