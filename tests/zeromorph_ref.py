"""Test helper: Zeromorph<UnivariateKzg>::open / verify (pcs/multilinear/zeromorph.rs:126-186, 216-245) restated with
Python integers and the oracle's MSM — written from the Rust, loop for loop, and sharing no code with the product
mirror (plonkish_b200/zeromorph.py) or its kernels.

The verifier's pairing check e(c, -[s^offset]_2) * e(pi, [s]_2 - x [1]_2) = 1 becomes, with the setup's trapdoor s known
to the test, the G1 equation (s - x) * pi = s^offset * c (big-integer group law, no pairing)."""
import numpy as np

from oracle import bigint_ref as br

R = br.R
MONT = 1 << 256


def to_mont(v: int) -> np.ndarray:
    return np.frombuffer((v % R * MONT % R).to_bytes(32, "little"), dtype=np.uint64).copy()


def mont_rows(values):
    return np.stack([to_mont(v) for v in values]) if len(values) else np.zeros((0, 4), dtype=np.uint64)


def commit_coeffs(oracle, powers_host, coeffs):
    """UnivariateKzg::commit_coeffs (univariate/kzg.rs:24-30): variable_base_msm(coeffs, &powers_of_s_g1[..coeffs.len()])."""
    if len(coeffs) == 0:
        return np.zeros(8, dtype=np.uint64)
    return oracle.variable_base_msm(mont_rows(coeffs), powers_host[: len(coeffs)])


def quotients(evals, point):
    """pcs/multilinear.rs:72-107."""
    remainder = [v % R for v in evals]
    qs = []
    for num_vars in reversed(range(len(point))):
        x_i = point[num_vars]
        lo, hi = remainder[: 1 << num_vars], remainder[1 << num_vars: 2 << num_vars]
        qs.append([(h - l) % R for l, h in zip(lo, hi)])
        remainder = [(l + (h - l) * x_i) % R for l, h in zip(lo, hi)]
    qs.reverse()
    return qs, remainder[0]


def eval_and_quotient_scalars(y, x, z, u):
    """zeromorph.rs:259-294, statement by statement."""
    num_vars = len(u)
    squares_of_x = [pow(x, 1 << i, R) for i in range(num_vars + 1)]
    offsets_of_x = [pow(x, (1 << num_vars) - (1 << i), R) for i in range(num_vars)]   # prod of squares_of_x[i..num_vars)
    v_numer = (squares_of_x[num_vars] - 1) % R
    vs = [v_numer * pow((sq - 1) % R, -1, R) % R for sq in squares_of_x]
    q_scalars = [(-(pow(y, i, R) * offsets_of_x[i] + z * (squares_of_x[i] * vs[i + 1] - u[i] * vs[i]))) % R for i in range(num_vars)]
    return (-vs[0] * z) % R, q_scalars


def open_reference(oracle, commit_powers, open_powers, evals, point, eval_, transcript):
    """zeromorph.rs:126-186.  evals / point / eval_: canonical integers; *_powers: host [n, 8] limb arrays."""
    num_vars = len(point)
    n = 1 << num_vars
    assert len(evals) == n
    qs, remainder = quotients(evals, point)
    for q in qs:                                                                       # batch_commit_and_write, :150
        transcript.write_commitment(commit_coeffs(oracle, commit_powers, q))
    y = transcript.squeeze_challenge()
    q_hat = [0] * n                                                                    # :157-168
    for idx, q in enumerate(qs):
        power_of_y = pow(y, idx, R)
        offset = n - (1 << idx)
        for j, v in enumerate(q):
            q_hat[offset + j] = (q_hat[offset + j] + power_of_y * v) % R
    transcript.write_commitment(commit_coeffs(oracle, commit_powers, q_hat))           # :169
    x = transcript.squeeze_challenge()
    z = transcript.squeeze_challenge()
    eval_scalar, q_scalars = eval_and_quotient_scalars(y, x, z, point)
    f = [(z * v + h) % R for v, h in zip(evals, q_hat)]                                # :175-177
    f[0] = (f[0] + eval_scalar * eval_) % R                                            # :178
    for q, scalar in zip(qs, q_scalars):                                               # :179
        for j, v in enumerate(q):
            f[j] = (f[j] + scalar * v) % R
    # UnivariateKzg::open(&pp.open_pp, &f, .., &x, &ZERO) (univariate/kzg.rs:264-299): the quotient of f by (X - x)
    quotient, carry = [0] * (n - 1), 0
    for i in reversed(range(1, n)):
        carry = (f[i] + x * carry) % R
        quotient[i - 1] = carry
    f_at_x = (f[0] + x * carry) % R
    transcript.write_commitment(commit_coeffs(oracle, open_powers, quotient))
    return remainder, f_at_x


def _point(limbs):
    return br.point_from_bytes(np.ascontiguousarray(limbs, dtype=np.uint64).tobytes())


def verify_in_g1(comm, point, eval_, q_comms, q_hat_comm, pi, transcript, s, offset):
    """zeromorph.rs:216-245 over G1: the verifier's transcript has already absorbed what precedes the opening; the proof's
    points are passed in as limb arrays and absorbed here in the order the verifier reads them."""
    transcript.write_commitments(q_comms)
    y = transcript.squeeze_challenge()
    transcript.write_commitment(q_hat_comm)
    x = transcript.squeeze_challenge()
    z = transcript.squeeze_challenge()
    eval_scalar, q_scalars = eval_and_quotient_scalars(y, x, z, point)
    scalars = [1, z, eval_scalar * eval_ % R] + q_scalars
    bases = [_point(q_hat_comm), _point(comm), br.G] + [_point(c) for c in q_comms]
    c = None
    for sc, b in zip(scalars, bases):
        c = br.add(c, br.scalar_mul(sc, b))
    transcript.write_commitment(pi)
    lhs = br.scalar_mul((s - x) % R, _point(pi))
    rhs = br.scalar_mul(pow(s, offset, R), c) if c is not None else None
    assert lhs == rhs, "the proof does not satisfy Zeromorph's verification equation"


def verify_reader_in_g1(reader, comm, point, eval_, s, offset=0):
    """The same equation over a proof reader (tests/hyperplonk_ref.py ProofReader: read_commitment(s) return affine integer
    pairs): Zeromorph::verify as the last step of additive::batch_verify inside HyperPlonk::verify.  comm: affine pair."""
    q_comms = reader.read_commitments(len(point))
    y = reader.squeeze_challenge()
    q_hat_comm = reader.read_commitment()
    x = reader.squeeze_challenge()
    z = reader.squeeze_challenge()
    eval_scalar, q_scalars = eval_and_quotient_scalars(y, x, z, point)
    c = br.msm([1, z, eval_scalar * eval_ % R] + q_scalars, [q_hat_comm, comm, br.G] + list(q_comms))
    pi = reader.read_commitment()
    assert br.scalar_mul((s - x) % R, pi) == br.scalar_mul(pow(s, offset, R), c), "Invalid Zeromorph KZG open"
