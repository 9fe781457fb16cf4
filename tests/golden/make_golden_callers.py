"""Generates tests/golden/caller_vectors.json with Python integers only (oracle/bigint_ref.py for the curve):
known answers for the rows either side of the MSM — `quotients` + the commitments of `open`
(pcs/multilinear.rs:72-107, pcs/multilinear/kzg.rs:276-302), the g_prime merge (pcs/multilinear.rs:203-213), the
eq tables and fixed-base multiples of `setup` (kzg.rs:174-208, util/arithmetic/msm.rs:16-81) and the classic
sum-check rounds (piop/sum_check/classic/eval.rs:101-131, poly/multilinear.rs:179-189).
Independent of the C oracle and of the CUDA path.  Run from the repo root:
    python tests/golden/make_golden_callers.py
The reference holds no known-answer vectors for these functions either and cannot be run here;
the vectors pin the mathematical values its code computes."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bigint_ref as br  # noqa: E402

R = br.R
rnd = random.Random(20261019)
fr = lambda v: br.scalar_to_bytes(v % R).hex()


def eq_tables(ss):
    eqs = [[1]]
    for s in ss:  # kzg.rs:178-192: lo = last - s*last, hi = s*last
        last = eqs[-1]
        hi = [s * e % R for e in last]
        eqs.append([(e - h) % R for e, h in zip(last, hi)] + hi)
    return eqs


def quotients(evals, point):
    rem = list(evals)
    qs = []
    for i in reversed(range(len(point))):  # multilinear.rs:80-103
        half = 1 << i
        qs.append([(rem[half + j] - rem[j]) % R for j in range(half)])
        rem = [(rem[j] + (rem[half + j] - rem[j]) * point[i]) % R for j in range(half)]
    qs.reverse()
    return qs, rem[0]


def kzg_case(k):
    ss = [rnd.randrange(R) for _ in range(k)]
    eqs = eq_tables(ss)
    eq_pts = [[br.scalar_mul(e, br.G) for e in tab] for tab in eqs]  # fixed_base_msm + normalise
    evals = [rnd.randrange(R) for _ in range(1 << k)]
    point = [rnd.randrange(R) for _ in range(k)]
    qs, value = quotients(evals, point)
    return {
        "num_vars": k,
        "ss": [fr(s) for s in ss],
        "eq_scalars": [[fr(e) for e in tab] for tab in eqs],
        "eq_points": [[br.point_to_bytes(p).hex() for p in tab] for tab in eq_pts],
        "evals": [fr(e) for e in evals],
        "point": [fr(x) for x in point],
        "commitment": br.point_to_bytes(br.msm(evals, eq_pts[k])).hex(),
        "quotients": [[fr(q) for q in qi] for qi in qs],
        "quotient_commitments": [br.point_to_bytes(br.msm(qi, eq_pts[i])).hex() for i, qi in enumerate(qs)],
        "eval": fr(value),
    }


def merge_case(count, n):
    polys = [[rnd.randrange(R) for _ in range(n)] for _ in range(count)]
    coeffs = [rnd.randrange(R) for _ in range(count)]
    out = [sum(c * p[j] for c, p in zip(coeffs, polys)) % R for j in range(n)]
    return {"polys": [[fr(v) for v in p] for p in polys], "coeffs": [fr(c) for c in coeffs], "result": [fr(v) for v in out]}


def sumcheck_case(num_polys, k, terms, common):
    polys = [[rnd.randrange(R) for _ in range(1 << k)] for _ in range(num_polys)]
    coeffs = [1 if t % 2 == 0 else rnd.randrange(R) for t in range(len(terms))]
    degree = max(max(len(t) for t in terms) + (1 if common >= 0 else 0), 1)

    def expr(vals):
        total = 0
        for c, idx in zip(coeffs, terms):
            prod = c
            for i in idx:
                prod = prod * vals[i] % R
            total += prod
        total %= R
        return total * vals[common] % R if common >= 0 else total

    rounds = []
    cur = polys
    for _ in range(k):
        msg = []
        for x in range(1, degree + 1):  # eval.rs:101-131 over the pairs (2b, 2b+1)
            msg.append(sum(expr([(p[2 * b] + x * (p[2 * b + 1] - p[2 * b])) % R for p in cur]) for b in range(len(cur[0]) // 2)) % R)
        ch = rnd.randrange(R)
        rounds.append({"evals_1_to_degree": [fr(v) for v in msg], "challenge": fr(ch)})
        cur = [[(p[2 * b] + (p[2 * b + 1] - p[2 * b]) * ch) % R for b in range(len(p) // 2)] for p in cur]  # multilinear.rs:615
    return {"num_vars": k, "polys": [[fr(v) for v in p] for p in polys], "coeffs": [fr(c) for c in coeffs], "terms": terms, "common": common,
            "degree": degree, "rounds": rounds, "final_evals": [fr(p[0]) for p in cur]}


out = {
    "generator": "tests/golden/make_golden_callers.py (Python integers; curve arithmetic from oracle/bigint_ref.py)",
    "encoding": "field elements: 32-byte little-endian Montgomery form (bn256::Fr); points: x||y Montgomery Fq, 64 zero bytes = identity",
    "kzg": [kzg_case(k) for k in (0, 1, 3, 5)],
    "merge": [merge_case(1, 5), merge_case(3, 8), merge_case(14, 4)],
    "fixed_base": None,
    "sumcheck": [
        sumcheck_case(1, 1, [[0]], -1),
        sumcheck_case(4, 3, [[1, 2], [3]], 0),                       # eq * (a*b - c) shape
        sumcheck_case(9, 4, [[1, 6], [2, 7], [3, 6, 7], [4, 8], [5]], 0),  # vanilla_plonk gate shape
        sumcheck_case(3, 5, [[0, 1, 2, 0], [], [2, 2]], -1),
    ],
}
scalars = [0, 1, 2, R - 1, 0x8000, 0x8001, 0xFFFF, 0x10000, (1 << 253) + 5] + [rnd.randrange(R) for _ in range(7)]
base = br.scalar_mul(rnd.randrange(1, R), br.G)
out["fixed_base"] = {"base": br.point_to_bytes(base).hex(), "scalars": [fr(s) for s in scalars],
                     "points": [br.point_to_bytes(br.scalar_mul(s, base)).hex() for s in scalars]}
path = os.path.join(ROOT, "tests", "golden", "caller_vectors.json")
with open(path, "w") as f:
    json.dump(out, f, indent=0)
print("wrote", path, os.path.getsize(path), "bytes")
