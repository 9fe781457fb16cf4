"""Generates tests/golden/pcs_vectors.json with Python integers only (oracle/bigint_ref.py for the curve, the loop-for-loop
restatement of tests/zeromorph_ref.py for the scheme): known-answer proofs of Zeromorph<UnivariateKzg>::open
(pcs/multilinear/zeromorph.rs:126-186) in the shape of the reference's own PCS test (run_commit_open_verify,
pcs/multilinear.rs:293-335: commit, squeeze the point, write the evaluation, open), one with a setup longer than the
polynomial (open_pp = the last 2^n powers, zeromorph.rs:84-102).  Independent of the C oracle and of the CUDA path; every
proof is checked against Zeromorph::verify's equation before it is written.  Run from the repo root:
    python tests/golden/make_golden_pcs.py
The reference holds no known-answer vectors for this scheme (its tests are randomised round trips) and cannot be run
here; the vectors pin the bytes its code writes for these inputs."""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import bigint_ref as br  # noqa: E402
import zeromorph_ref as zr  # noqa: E402
from plonkish_b200.transcript import Keccak256Transcript  # noqa: E402  (the byte-exact mirror of util/transcript.rs, pinned by its own tests)

R = br.R
rnd = random.Random(20261019)


class IntMsm:
    """variable_base_msm over Montgomery limb arrays with Python integers (what zeromorph_ref.commit_coeffs calls)."""

    @staticmethod
    def variable_base_msm(scalars, bases):
        ks = [br.scalar_from_bytes(np.ascontiguousarray(row).tobytes()) for row in scalars]
        pts = [br.point_from_bytes(np.ascontiguousarray(row).tobytes()) for row in bases]
        return np.frombuffer(br.point_to_bytes(br.msm(ks, pts)), dtype=np.uint64).copy()


def limbs(pt):
    return np.frombuffer(br.point_to_bytes(pt), dtype=np.uint64).copy()


def case(num_vars, extra):
    n = 1 << num_vars
    s = rnd.randrange(2, R)
    powers = np.stack([limbs(br.scalar_mul(pow(s, i, R), br.G)) for i in range(n + extra)])
    evals = [rnd.randrange(R) for _ in range(n)]
    t = Keccak256Transcript()
    comm = zr.commit_coeffs(IntMsm, powers[:n], evals)
    t.write_commitment(comm)
    point = t.squeeze_challenges(num_vars)
    value = evals
    for x in point:                                        # MultilinearPolynomial::evaluate, lowest variable first
        value = [(value[2 * b] + (value[2 * b + 1] - value[2 * b]) * x) % R for b in range(len(value) // 2)]
    value = value[0]
    t.write_field_element(value)
    remainder, f_at_x = zr.open_reference(IntMsm, powers[:n], powers[extra:], evals, point, value, t)
    assert remainder == value and f_at_x == 0
    proof = t.into_proof()
    pts = [limbs((int.from_bytes(proof[i:i + 32], "big"), int.from_bytes(proof[i + 32:i + 64], "big"))) for i in range(96, len(proof), 64)]
    v = Keccak256Transcript()
    v.write_commitment(comm)
    v.squeeze_challenges(num_vars)
    v.write_field_element(value)
    zr.verify_in_g1(comm, point, value, pts[:num_vars], pts[num_vars], pts[num_vars + 1], v, s, extra)
    return {"num_vars": num_vars, "extra": extra, "s": hex(s), "evals": [br.scalar_to_bytes(e).hex() for e in evals],
            "powers_of_s_g1": [row.tobytes().hex() for row in powers], "point": [hex(x) for x in point], "eval": hex(value), "proof": proof.hex()}


if __name__ == "__main__":
    out = {"scheme": "Zeromorph<UnivariateKzg<Bn256>>, Keccak256Transcript", "cases": [case(3, 0), case(4, 5)]}
    path = os.path.join(ROOT, "tests", "golden", "pcs_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, [len(c["proof"]) // 2 for c in out["cases"]])
