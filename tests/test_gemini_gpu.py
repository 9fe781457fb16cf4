"""GPU parity suite (-m gpu): Gemini<UnivariateKzg<Bn256>> (pcs/multilinear/gemini.rs) through the C ABI — the packed
folds, slice handles and mixed-length sums against Python integers / the oracle, the proofs of open / batch_open byte for
byte against the same host logic driven through the oracle, and against Gemini::verify restated in G1 with the setup's
trapdoor (tests/gemini_ref.py)."""
import numpy as np
import pytest

from oracle import bigint_ref as br
import gemini_ref as gr
import zeromorph_ref as zr
from test_gemini_cpu import _oracle_ops

pytestmark = pytest.mark.gpu
R = br.R


@pytest.fixture(scope="module")
def pk():
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import plonkish_b200

    plonkish_b200._lib.lib()
    return plonkish_b200


def _mont(v):
    from plonkish_b200.sumcheck import _to_mont

    return _to_mont(v)


def _ints(arr):
    from plonkish_b200.sumcheck import _to_int

    return [_to_int(row) for row in np.asarray(arr).reshape(-1, 4)]


@pytest.mark.parametrize("num_vars", [1, 2, 5, 11, 14, 18])
def test_folds_slices_and_evaluations_match_the_oracle(pk, oracle, num_vars):
    # gemini.rs:98-108 packed (f_i at element offset 2^(n-i)); a slice is a polynomial of its own for div_linear
    n = 1 << num_vars
    poly_h = pk.random_scalars(n, seed=800 + num_vars)
    point = pk.random_scalars(num_vars, seed=801)
    poly = pk.ResidentScalars(poly_h)
    folds = pk.fr_gemini_folds(poly, point)
    got = folds.to_host()
    assert not got[: min(2, n)].any()
    cur = poly_h
    z = pk.random_scalars(1, seed=802)[0]
    for i in range(1, num_vars):
        cur = oracle.fix_var(cur, point[i - 1])
        size = 1 << (num_vars - i)
        assert got[size: 2 * size].tobytes() == cur.tobytes(), i
        if i in (1, num_vars - 1):
            view = pk.scalars_slice(folds, size, size)
            q, rem = pk.fr_div_linear(view, z)
            want_q, want_rem = oracle.fr_div_linear(cur, z)
            assert rem.tobytes() == want_rem.tobytes() and q.to_host()[: size - 1].tobytes() == want_q.tobytes()
            q.release()
            view.release()
    # a slice keeps the memory alive after its parent is released
    if num_vars >= 2:
        view = pk.scalars_slice(folds, 2, 2)
        folds.release()
        assert view.to_host().tobytes() == got[2:4].tobytes()
        view.release()
    else:
        folds.release()
    with pytest.raises(pk.PlonkishCudaError):
        pk.scalars_slice(poly, n - 1, 2)
    poly.release()


def test_mixed_length_sums_match_the_oracle(pk, oracle):
    lens = [1 << 12, 1, 1 << 11, 1 << 12, 7, 1 << 10, 2, 1 << 12, 33, 8, 4, 4000, 1 << 12, 5]
    hosts = [pk.random_scalars(m, seed=850 + i) for i, m in enumerate(lens)]
    coeffs = pk.random_scalars(len(lens), seed=870)
    res = [pk.ResidentScalars(h) for h in hosts]
    out = pk.fr_linear_combination_padded(res, coeffs)
    n = max(lens)
    padded = [h if len(h) == n else np.concatenate([h, np.zeros((n - len(h), 4), dtype=np.uint64)]) for h in hosts]
    assert out.to_host().tobytes() == oracle.fr_linear_combination(padded, coeffs).tobytes()
    with pytest.raises(pk.PlonkishCudaError):  # the unpadded entry still refuses short polynomials
        import ctypes

        hs = np.array([r.handle for r in res], dtype=np.uint64)
        h = ctypes.c_uint64(0)
        pk._lib.check(pk._lib.lib().plonkish_cuda_fr_linear_combination(hs.ctypes.data, coeffs.ctypes.data, len(res), n, ctypes.byref(h)), "lincomb")
    for r_ in res + [out]:
        r_.release()


def _open_both_ways(pk, oracle, num_vars, s, seed):
    from plonkish_b200 import gemini, kzg
    from plonkish_b200.transcript import Keccak256Transcript

    n = 1 << num_vars
    srs = kzg.univariate_setup(oracle.generator(), _mont(s), n)
    pp = gemini.GeminiKzgProverParam(srs)
    poly_h = pk.random_scalars(n, seed=seed)
    poly = pk.ResidentScalars(poly_h)
    t = Keccak256Transcript()
    comm = gemini.commit(pp, poly)
    t.write_commitment(comm)
    point = t.squeeze_challenges(num_vars)
    value = _ints(oracle.evaluate_multilinear(poly_h, zr.mont_rows(point), oracle.host_threads()))[0]
    t.write_field_element(value)
    gemini.open(pp, poly, point, t)
    poly.release()
    return pp, poly_h, comm, point, value, t.into_proof()


def _verify(oracle, proof, num_vars, point, value, s):
    from hyperplonk_ref import ProofReader

    reader = ProofReader(oracle.keccak256, proof)
    c = reader.read_commitment()
    assert reader.squeeze_challenges(num_vars) == point
    assert reader.read_field_element() == value
    gr.verify_reader_in_g1(reader, c, point, value, s)
    assert reader.pos == len(proof)


@pytest.mark.parametrize("num_vars", [2, 3, 6, 10, 13])
def test_open_writes_the_oracle_driven_proof_bytes_and_verifies(pk, oracle, num_vars):
    from plonkish_b200 import gemini
    from plonkish_b200.transcript import Keccak256Transcript

    s = 0xA5A5A5A5C3C3C3C3F0F0F0F0 % R
    pp, poly_h, comm, point, value, proof = _open_both_ways(pk, oracle, num_vars, s, 900 + num_vars)
    # the same host logic with every polynomial operation done by the oracle on the CPU
    srs_h = pp.powers_of_s_g1.to_host()
    t = Keccak256Transcript()
    ops = _oracle_ops(oracle)
    assert gemini.commit(gemini.GeminiKzgProverParam(srs_h), poly_h, ops).tobytes() == comm.tobytes()
    t.write_commitment(comm)
    assert t.squeeze_challenges(num_vars) == point
    t.write_field_element(value)
    gemini.open(gemini.GeminiKzgProverParam(srs_h), poly_h, point, t, ops)
    assert proof == t.into_proof()
    _verify(oracle, proof, num_vars, point, value, s)
    pp.release()


@pytest.mark.parametrize("num_vars", [16, 18])
def test_open_of_a_larger_polynomial_verifies(pk, oracle, num_vars):
    s = 0x31415926535897932384626433 % R
    pp, _, _, point, value, proof = _open_both_ways(pk, oracle, num_vars, s, 950 + num_vars)
    _verify(oracle, proof, num_vars, point, value, s)
    with pytest.raises(AssertionError):
        _verify(oracle, proof[:64] + ((value + 1) % R).to_bytes(32, "big") + proof[96:], num_vars, point, (value + 1) % R, s)
    pp.release()


@pytest.mark.parametrize("num_vars", [3, 6])
def test_batch_open_writes_the_reference_proof_bytes(pk, oracle, num_vars):
    # additive::batch_open (pcs/multilinear.rs:134-235) with Gemini's open on g_prime (gemini.rs:143-158): the sum check and
    # the merge against the integer restatement, the opening against the oracle-driven host logic
    from batch_open_ref import batch_open_reference, to_int
    from plonkish_b200 import gemini, kzg
    from plonkish_b200.transcript import Keccak256Transcript

    n, s = 1 << num_vars, 0x271828182845904523536 % R
    pp = gemini.GeminiKzgProverParam(kzg.univariate_setup(oracle.generator(), _mont(s), n))
    srs_h = pp.powers_of_s_g1.to_host()
    polys_h = [pk.random_scalars(n, seed=970 + i) for i in range(4)]
    polys_i = [[to_int(r) for r in p] for p in polys_h]
    rng = np.random.default_rng(num_vars)
    points = [[int(x) for x in rng.integers(1, 1 << 62, num_vars)] for _ in range(3)]
    pairs = [(0, 0), (0, 1), (1, 0), (2, 2), (3, 1), (3, 0)]

    def evaluate(poly, pt):
        cur = poly
        for x in pt:
            cur = [(cur[2 * b] + (cur[2 * b + 1] - cur[2 * b]) * x) % R for b in range(len(cur) // 2)]
        return cur[0]

    evals = [(p, x, evaluate(polys_i[p], points[x])) for p, x in pairs]
    resident = [pk.ResidentScalars(p) for p in polys_h]
    t_gpu, t_ref = Keccak256Transcript(), Keccak256Transcript()
    for tr in (t_gpu, t_ref):
        tr.write_commitments(gemini.batch_commit(pp, resident))
        tr.write_field_elements([v for _, _, v in evals])
    gemini.batch_open(pp, num_vars, resident, points, evals, t_gpu)
    ops = _oracle_ops(oracle)

    def open_ref(g_prime, challenges, transcript):
        gemini.open(gemini.GeminiKzgProverParam(srs_h), zr.mont_rows(g_prime), challenges, transcript, ops)

    batch_open_reference(oracle, None, num_vars, polys_i, points, evals, t_ref, open_fn=open_ref)
    assert t_gpu.into_proof() == t_ref.into_proof()
    for r_ in resident:
        r_.release()
    pp.release()
