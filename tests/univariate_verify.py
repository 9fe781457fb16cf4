"""Test helper: UnivariateKzg::batch_verify (pcs/univariate/kzg.rs:380-417) restated over G1 only.

The reference checks e(pi * z + f - eval * G, -G2) * e(pi, s * G2) = 1; with the setup's trapdoor s known to the test
that is the G1 equation (s - z) * pi = f - eval * G, evaluated here with the big-integer reference (no pairing, no
code shared with the prover under test except the set bookkeeping the reference's prover and verifier share too)."""
import numpy as np

from oracle import bigint_ref as br

R = br.R


def as_limbs(pt):
    return np.frombuffer(br.point_to_bytes(pt), dtype=np.uint64)


def _point(b):
    return br.point_from_bytes(np.ascontiguousarray(b, dtype=np.uint64).tobytes())


def batch_verify_in_g1(comms, points, evals, q_comm, pi, s):
    """comms: [8]-limb commitments as written to the transcript before batch_open; q_comm, pi: the two points of the proof."""
    from plonkish_b200 import univariate
    from plonkish_b200.transcript import Keccak256Transcript

    v = Keccak256Transcript()                                          # the verifier absorbs what it reads
    v.write_commitments(comms)
    sets, superset = univariate.eval_sets(evals)
    beta, gamma = v.squeeze_challenge(), v.squeeze_challenge()
    v.write_commitment(as_limbs(q_comm))
    zc = v.squeeze_challenge()
    _check_equation(sets, superset, beta, gamma, zc, [_point(c) for c in comms], q_comm, pi, points, s)
    return sets


def batch_verify_reader_in_g1(reader, comms, points, evals, s):
    """The same over a proof reader (tests/hyperplonk_ref.py ProofReader) positioned at batch_open's first challenge: what
    Gemini::verify ends with (pcs/multilinear/gemini.rs:196).  comms: affine integer pairs."""
    from plonkish_b200 import univariate

    sets, superset = univariate.eval_sets(evals)
    beta, gamma = reader.squeeze_challenge(), reader.squeeze_challenge()
    q_comm = reader.read_commitment()
    zc = reader.squeeze_challenge()
    pi = reader.read_commitment()
    _check_equation(sets, superset, beta, gamma, zc, comms, q_comm, pi, points, s)
    return sets


def _check_equation(sets, superset, beta, gamma, zc, comms, q_comm, pi, points, s):
    """kzg.rs:380-417 after the transcript reads; comms, q_comm, pi: affine integer pairs."""
    from plonkish_b200 import univariate

    pb = univariate._powers(beta, max(len(st.polys) for st in sets))
    pg = univariate._powers(gamma, len(sets))
    normalized, normalizer = univariate.set_scalars(sets, pg, points, zc)
    scalars = [0] * len(comms)
    for st, coeff in zip(sets, normalized):                            # comm_scalars, kzg.rs:541-552
        for poly, b in zip(st.polys, pb):
            scalars[poly] = coeff * b % R
    q_scalar = (-univariate.vanishing_eval([points[i] for i in superset], zc) * normalizer) % R
    f = None
    for sc, c in zip(scalars, comms):
        f = br.add(f, br.scalar_mul(sc, c))
    f = br.add(f, br.scalar_mul(q_scalar, q_comm))

    def r_eval(st):                                                    # kzg.rs:442-451: interpolate every poly's evals over the set's points at z
        xs = [points[i] for i in st.points]
        total = 0
        for vals, b in zip(st.evals, pb):
            acc = 0
            for k, xk in enumerate(xs):
                num, den = 1, 1
                for j, xj in enumerate(xs):
                    if j != k:
                        num = num * (zc - xj) % R
                        den = den * (xk - xj) % R
                acc = (acc + vals[k] * num % R * pow(den, -1, R)) % R
            total = (total + b * acc) % R
        return total

    ev = sum(ns * r_eval(st) for ns, st in zip(normalized, sets)) % R
    lhs = br.scalar_mul((s - zc) % R, pi)
    rhs = br.add(f, br.neg(br.scalar_mul(ev, br.G)))
    assert lhs == rhs, "the proof does not satisfy batch_verify's equation"
