"""CPU suite for the Gemini mirror (pcs/multilinear/gemini.rs): the fold and the mixed-length sum kernels in the emulator
against Python integers, and the host logic of plonkish_b200/gemini.py — driven through the oracle instead of the GPU —
against Gemini::verify restated over G1 (tests/gemini_ref.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import bigint_ref as br
import gemini_ref as gr
import zeromorph_ref as zr

R = br.R


def _ints(arr):
    from plonkish_b200.sumcheck import _to_int

    return [_to_int(row) for row in np.asarray(arr).reshape(-1, 4)]


@pytest.fixture(scope="module")
def emul():
    emul_dir = os.path.join(ROOT, "tests", "emul")
    subprocess.run(["make", "-C", emul_dir], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(emul_dir, "libemul_msm.so"))
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    lib.emul_gemini_folds.argtypes = [vp, u32, vp, vp]
    lib.emul_fr_lincomb_padded.argtypes = [vp, vp, vp, u32, u32, vp]
    return lib


@pytest.mark.parametrize("num_vars", [1, 2, 3, 6, 10])
def test_fold_kernel_matches_python_integers(emul, oracle, num_vars):
    # gemini.rs:98-108: f_i (2^(n-i) values) at element offset 2^(n-i) of the packed vector
    n = 1 << num_vars
    poly = oracle.random_scalars(n, 40 + num_vars)
    point = oracle.random_scalars(num_vars, 41)
    out = np.full((n, 4), 0xEE, dtype=np.uint64)
    emul.emul_gemini_folds(poly.ctypes.data, num_vars, point.ctypes.data, out.ctypes.data)
    want = gr.folds(_ints(poly), _ints(point))
    assert len(want) == num_vars
    for i in range(1, num_vars):
        size = 1 << (num_vars - i)
        assert out[size: 2 * size].tobytes() == zr.mont_rows(want[i]).tobytes(), i


def test_mixed_length_sums_match_python_integers(emul, oracle):
    # `f += (scalar, q)` over polynomials of different lengths (poly/univariate.rs), more terms than one launch holds
    lens = [64, 1, 32, 64, 7, 16, 2, 64, 33, 8, 4, 60, 64, 5]
    polys = [oracle.random_scalars(m, 70 + i) for i, m in enumerate(lens)]
    coeffs = oracle.random_scalars(len(lens), 99)
    ptrs = (ctypes.c_void_p * len(lens))(*[p.ctypes.data for p in polys])
    lens_a = np.array(lens, dtype=np.uint64)
    out = np.full((64, 4), 0x77, dtype=np.uint64)
    emul.emul_fr_lincomb_padded(ctypes.cast(ptrs, ctypes.c_void_p), lens_a.ctypes.data, coeffs.ctypes.data, len(lens), 64, out.ctypes.data)
    ci = _ints(coeffs)
    want = [0] * 64
    for c, p in zip(ci, polys):
        for j, v in enumerate(_ints(p)):
            want[j] = (want[j] + c * v) % R
    assert out.tobytes() == zr.mont_rows(want).tobytes()


def _oracle_ops(oracle):
    from plonkish_b200.sumcheck import _to_int, _to_mont

    def pad(p, n):
        return p if len(p) == n else np.concatenate([p, np.zeros((n - len(p), 4), dtype=np.uint64)])

    class OracleOps:
        @staticmethod
        def linear_combination(polys, coeffs):
            n = max(len(p) for p in polys)
            return oracle.fr_linear_combination([pad(p, n) for p in polys], np.stack([_to_mont(c) for c in coeffs]))

        @staticmethod
        def folds(poly, point):
            n = len(poly)
            packed = np.zeros((n, 4), dtype=np.uint64)
            cur = poly
            for x in point[:-1]:
                cur = oracle.fix_var(cur, _to_mont(x))          # (e1 - e0) * x + e0 on consecutive pairs
                packed[len(cur): 2 * len(cur)] = cur
            return packed

        @staticmethod
        def fold_views(folds, num_vars):
            return [folds[1 << (num_vars - i): 2 << (num_vars - i)] for i in range(1, num_vars)]

        @staticmethod
        def commit_folds(srs, folds, num_vars):
            out = [oracle.variable_base_msm(folds[1 << (num_vars - i): 2 << (num_vars - i)], srs[: 1 << (num_vars - i)]) for i in range(1, num_vars)]
            return np.stack(out) if out else np.zeros((0, 8), dtype=np.uint64)

        @staticmethod
        def evaluate(poly, x):
            return _to_int(oracle.fr_div_linear(poly, _to_mont(x))[1])

        @staticmethod
        def div_linear(poly, z):
            q, rem = oracle.fr_div_linear(poly, _to_mont(z))
            return np.concatenate([q, np.zeros((1, 4), dtype=np.uint64)]), _to_int(rem)

        @staticmethod
        def commit(srs, poly):
            return oracle.variable_base_msm(poly, srs[: len(poly)])

        @staticmethod
        def release(p):
            pass

    return OracleOps


@pytest.mark.parametrize("num_vars", [2, 3, 6])
def test_open_host_logic_satisfies_the_verifier(oracle, num_vars):
    # run_commit_open_verify (pcs/multilinear.rs:293-335) for Gemini with the oracle as `ops`
    from hyperplonk_ref import ProofReader
    from plonkish_b200 import gemini
    from plonkish_b200.sumcheck import _to_mont
    from plonkish_b200.transcript import Keccak256Transcript

    n, s = 1 << num_vars, 0xDEC0DE5EED1234
    srs = oracle.fixed_base_msm(oracle.generator(), np.stack([_to_mont(pow(s, i, R)) for i in range(n)]))
    pp = gemini.GeminiKzgProverParam(srs)
    ops = _oracle_ops(oracle)
    poly = oracle.random_scalars(n, 120 + num_vars)
    t = Keccak256Transcript()
    comm = gemini.commit(pp, poly, ops)
    t.write_commitment(comm)
    point = t.squeeze_challenges(num_vars)
    value = _ints(oracle.evaluate_multilinear(poly, zr.mont_rows(point)))[0]
    t.write_field_element(value)
    gemini.open(pp, poly, point, t, ops)
    proof = t.into_proof()
    assert len(proof) == 64 + 32 + (num_vars - 1) * 64 + num_vars * 32 + 2 * 64
    # the evaluations the prover wrote are those of the integer folds
    fs = gr.folds(_ints(poly), point)

    def verify(eval_):
        reader = ProofReader(oracle.keccak256, proof)
        c = reader.read_commitment()
        assert reader.squeeze_challenges(num_vars) == point
        assert reader.read_field_element() == value
        gr.verify_reader_in_g1(reader, c, point, eval_, s)
        assert reader.pos == len(proof)

    verify(value)
    with pytest.raises(AssertionError):
        verify((value + 1) % R)
    assert (fs[-1][0] * (1 - point[-1]) + fs[-1][1] * point[-1]) % R == value   # the sanity check of gemini.rs:110-117


def test_open_of_one_variable_fails_like_the_reference(oracle):
    # a polynomial of 2 coefficients divided by (X - beta)(X + beta) leaves a zero quotient, whose commitment is the
    # identity: write_commitment refuses it (util/transcript.rs:175-181) — the reference's tests start at 2 variables
    from plonkish_b200 import gemini
    from plonkish_b200.sumcheck import _to_mont
    from plonkish_b200.transcript import Keccak256Transcript

    srs = oracle.fixed_base_msm(oracle.generator(), np.stack([_to_mont(pow(5, i, R)) for i in range(2)]))
    with pytest.raises(ValueError, match="Invalid elliptic curve point encoding"):
        gemini.open(gemini.GeminiKzgProverParam(srs), oracle.random_scalars(2, 1), [12345], Keccak256Transcript(), _oracle_ops(oracle))
