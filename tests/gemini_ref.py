"""Test helper: Gemini<UnivariateKzg>::verify (pcs/multilinear/gemini.rs:168-197) restated over a proof reader, with the
final UnivariateKzg::batch_verify evaluated in G1 through the setup's trapdoor (tests/univariate_verify.py) — and the
folds / evaluations of Gemini::open (gemini.rs:98-138) with Python integers, independent of the product mirror."""
from oracle import bigint_ref as br
from univariate_verify import batch_verify_reader_in_g1

R = br.R


def folds(evals, point):
    """fs of gemini.rs:98-108: fs[0] = evals, fs[i] = merge_into(fs[i-1], point[i-1], 1, 0) (poly/multilinear.rs:599-618)."""
    fs = [[v % R for v in evals]]
    for x_i in point[: len(point) - 1]:
        prev = fs[-1]
        fs.append([((prev[2 * j + 1] - prev[2 * j]) * x_i + prev[2 * j]) % R for j in range(len(prev) // 2)])
    return fs


def horner(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def open_points_and_evals(fs, beta):
    """gemini.rs:130-137."""
    num_vars = len(fs)
    points = [beta % R] + [(-pow(beta, 1 << i, R)) % R for i in range(num_vars)]
    queries = [(0, 0), (0, 1)] + [(i, i + 1) for i in range(1, num_vars)]
    return points, [(idx, pt, horner(fs[idx], points[pt])) for idx, pt in queries]


def verify_reader_in_g1(reader, comm, point, eval_, s):
    """gemini.rs:168-197; comm: affine integer pair; raises AssertionError where the reference returns Err."""
    num_vars = len(point)
    comms = [comm] + reader.read_commitments(num_vars - 1)
    beta = reader.squeeze_challenge()
    squares_of_beta = [pow(beta, 1 << i, R) for i in range(num_vars)]
    evals = reader.read_field_elements(num_vars)
    eval_0 = eval_ % R
    for eval_neg, sq, x_i in reversed(list(zip(evals, squares_of_beta, point))):   # the fold of :184-190
        eval_0 = (2 * sq * eval_0 - ((1 - x_i) * sq - x_i) * eval_neg) * pow(((1 - x_i) * sq + x_i) % R, -1, R) % R
    queries = [(0, 0), (0, 1)] + [(i, i + 1) for i in range(1, num_vars)]
    full = [(idx, pt, v) for (idx, pt), v in zip(queries, [eval_0] + evals)]
    points = [beta] + [(-sq) % R for sq in squares_of_beta]
    batch_verify_reader_in_g1(reader, comms, points, full, s)
