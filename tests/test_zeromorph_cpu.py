"""CPU suite for the Zeromorph mirror (pcs/multilinear/zeromorph.rs): the two kernels in the emulator against Python
integers, and the host logic of plonkish_b200/zeromorph.py — driven through the oracle instead of the GPU — against the
loop-for-loop restatement in tests/zeromorph_ref.py (identical proof bytes) and the verifier's equation."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import bigint_ref as br
from univariate_verify import as_limbs
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
    lib.emul_zm_q_hat.argtypes = [vp, vp, u32, vp]
    lib.emul_zm_f.argtypes = [vp, vp, vp, vp, vp, vp, u32, vp]
    return lib


@pytest.mark.parametrize("num_vars", [0, 1, 2, 3, 5, 8, 11])
def test_q_hat_and_f_kernels_match_python_integers(emul, oracle, num_vars):
    # zeromorph.rs:157-168 and :175-180 from the packed quotient buffer (q_i at element offset 2^i)
    n = 1 << num_vars
    q = oracle.random_scalars(n, 500 + num_vars)
    q[0] = 0
    poly = oracle.random_scalars(n, 600 + num_vars)
    w = oracle.random_scalars(max(num_vars, 1), 700 + num_vars)[:num_vars]
    z, c0 = oracle.random_scalars(2, 800 + num_vars)
    qi, pi_, wi = _ints(q), _ints(poly), _ints(w)
    zi, ci = _ints(z)[0], _ints(c0)[0]
    want_hat = [0] * n
    for i in range(num_vars):
        for j in range(1 << i):
            want_hat[n - (1 << i) + j] = (want_hat[n - (1 << i) + j] + wi[i] * qi[(1 << i) + j]) % R
    got_hat = np.full((n, 4), 0xAB, dtype=np.uint64)
    wbuf = np.ascontiguousarray(w) if num_vars else np.zeros((1, 4), dtype=np.uint64)
    emul.emul_zm_q_hat(q.ctypes.data, wbuf.ctypes.data, num_vars, got_hat.ctypes.data)
    assert got_hat.tobytes() == zr.mont_rows(want_hat).tobytes()
    want_f = [(zi * a + b) % R for a, b in zip(pi_, want_hat)]
    want_f[0] = (want_f[0] + ci) % R
    for i in range(num_vars):
        for j in range(1 << i):
            want_f[j] = (want_f[j] + wi[i] * qi[(1 << i) + j]) % R
    got_f = np.full((n, 4), 0xCD, dtype=np.uint64)
    emul.emul_zm_f(poly.ctypes.data, got_hat.ctypes.data, q.ctypes.data, wbuf.ctypes.data, z.ctypes.data, c0.ctypes.data, num_vars, got_f.ctypes.data)
    assert got_f.tobytes() == zr.mont_rows(want_f).tobytes()


def test_eval_and_quotient_scalars_agree_with_the_restatement():
    from plonkish_b200 import zeromorph

    for num_vars in (0, 1, 4, 9):
        y, x, z = 0x1234567 + num_vars, 0xABCDEF0123 * 7 + num_vars, R - 5
        u = [(0x9999 * (i + 3)) ** 3 % R for i in range(num_vars)]
        assert zeromorph.eval_and_quotient_scalars(y, x, z, u) == zr.eval_and_quotient_scalars(y, x, z, u)


def _oracle_ops(oracle):
    from plonkish_b200.sumcheck import _to_int, _to_mont

    class OracleOps:
        @staticmethod
        def quotients(poly, point):
            n = len(poly)
            qs, value = oracle.quotients(poly, zr.mont_rows(point)) if point else ([], poly[0])
            packed = np.zeros((n, 4), dtype=np.uint64)
            for i, q in enumerate(qs):
                packed[1 << i: 2 << i] = q
            return packed, _to_int(value)

        @staticmethod
        def commit_quotients(srs, q, num_vars):
            out = [oracle.variable_base_msm(q[1 << i: 2 << i], srs[: 1 << i]) for i in range(num_vars)]
            return np.stack(out) if out else np.zeros((0, 8), dtype=np.uint64)

        @staticmethod
        def q_hat(q, weights):
            n = len(q)
            out = np.zeros((n, 4), dtype=np.uint64)
            for i, w in enumerate(weights):
                seg = oracle.fr_linear_combination([q[1 << i: 2 << i]], _to_mont(w).reshape(1, 4))
                lo = n - (1 << i)
                out[lo:] = oracle.fr_vec_op("add", out[lo:], np.concatenate([seg, np.zeros((n - lo - (1 << i), 4), dtype=np.uint64)]))
            return out

        @staticmethod
        def f(poly, q_hat, q, z, c0, q_scalars):
            out = oracle.fr_vec_op("add", oracle.fr_linear_combination([poly], _to_mont(z).reshape(1, 4)), q_hat)
            out[0] = oracle.fr_vec_op("add", out[:1], _to_mont(c0).reshape(1, 4))[0]
            for i, s in enumerate(q_scalars):
                seg = oracle.fr_linear_combination([q[1 << i: 2 << i]], _to_mont(s).reshape(1, 4))
                out[: 1 << i] = oracle.fr_vec_op("add", out[: 1 << i], seg)
            return out

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


def _srs(oracle, s, n):
    from plonkish_b200.sumcheck import _to_mont

    return oracle.fixed_base_msm(oracle.generator(), np.stack([_to_mont(pow(s, i, R)) for i in range(n)]))


def _proof_limbs(proof: bytes):
    assert len(proof) % 64 == 0
    return [as_limbs((int.from_bytes(proof[i:i + 32], "big"), int.from_bytes(proof[i + 32:i + 64], "big"))) for i in range(0, len(proof), 64)]


@pytest.mark.parametrize("num_vars,extra", [(1, 0), (3, 0), (6, 0), (5, 9)])
def test_open_host_logic_matches_the_restatement_and_verifies(oracle, num_vars, extra):
    # zeromorph.rs:126-186 through the product's host code with the oracle as `ops`; extra > 0: a longer setup, so that
    # open_pp = powers[offset..] differs from commit_pp (trim, :84-102)
    from plonkish_b200 import zeromorph
    from plonkish_b200.transcript import Keccak256Transcript

    n, s = 1 << num_vars, 0xFEEDFACE12345
    full = _srs(oracle, s, n + extra)
    pp = zeromorph.ZeromorphKzgProverParam(full[:n], full[extra:])
    ops = _oracle_ops(oracle)
    poly = oracle.random_scalars(n, 90 + num_vars)
    evals = _ints(poly)
    t = Keccak256Transcript()
    comm = zeromorph.commit(pp, poly, ops)
    t.write_commitment(comm)
    point = t.squeeze_challenges(num_vars)
    value = _ints(oracle.evaluate_multilinear(poly, zr.mont_rows(point)))[0] if num_vars else evals[0]
    t.write_field_element(value)
    t_ref = Keccak256Transcript()
    t_ref.write_commitment(comm)
    assert t_ref.squeeze_challenges(num_vars) == point
    t_ref.write_field_element(value)
    remainder = zeromorph.open(pp, poly, point, value, t, ops)
    want_remainder, f_at_x = zr.open_reference(oracle, full[:n], full[extra:], evals, point, value, t_ref)
    assert remainder == want_remainder == value and f_at_x == 0
    proof = t.into_proof()
    assert proof == t_ref.into_proof()
    pts = _proof_limbs(proof[64 + 32:])
    assert len(pts) == num_vars + 2
    v = Keccak256Transcript()
    v.write_commitment(comm)
    assert v.squeeze_challenges(num_vars) == point
    v.write_field_element(value)
    zr.verify_in_g1(comm, point, value, pts[:num_vars], pts[num_vars], pts[num_vars + 1], v, s, extra)
    with pytest.raises(AssertionError):  # a wrong evaluation must not verify
        v2 = Keccak256Transcript()
        v2.write_commitment(comm)
        v2.squeeze_challenges(num_vars)
        v2.write_field_element(value)
        zr.verify_in_g1(comm, point, (value + 1) % R, pts[:num_vars], pts[num_vars], pts[num_vars + 1], v2, s, extra)


def test_commit_and_open_reject_polynomials_larger_than_the_param(oracle):
    from plonkish_b200 import zeromorph
    from plonkish_b200.transcript import Keccak256Transcript

    srs = _srs(oracle, 77, 4)
    pp = zeromorph.ZeromorphKzgProverParam(srs, srs)
    poly = oracle.random_scalars(8, 1)
    with pytest.raises(ValueError, match="Too large degree of poly to commit"):
        zeromorph.commit(pp, poly, _oracle_ops(oracle))
    with pytest.raises(ValueError, match="Too large degree of poly to open"):
        zeromorph.open(pp, poly, [1, 2, 3], 0, Keccak256Transcript(), _oracle_ops(oracle))


def _golden_cases():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "pcs_vectors.json")) as f:
        return json.load(f)["cases"]


def golden_inputs(case):
    """(powers [n + extra, 8], evals [n, 4]) limb arrays and the integers of a committed known-answer case."""
    powers = np.frombuffer(bytes.fromhex("".join(case["powers_of_s_g1"])), dtype=np.uint64).reshape(-1, 8).copy()
    evals = np.frombuffer(bytes.fromhex("".join(case["evals"])), dtype=np.uint64).reshape(-1, 4).copy()
    return powers, evals, [int(x, 16) for x in case["point"]], int(case["eval"], 16), bytes.fromhex(case["proof"])


@pytest.mark.parametrize("idx", [0, 1])
def test_golden_proofs_through_the_host_logic_and_the_c_oracle(oracle, idx):
    # tests/golden/pcs_vectors.json was made with Python integers only (make_golden_pcs.py); here the product's host logic
    # runs over the C oracle's MSM and vector operations and must write the same bytes
    from plonkish_b200 import zeromorph
    from plonkish_b200.transcript import Keccak256Transcript

    case = _golden_cases()[idx]
    powers, evals, point, value, proof = golden_inputs(case)
    n, extra = 1 << case["num_vars"], case["extra"]
    pp = zeromorph.ZeromorphKzgProverParam(powers[:n], powers[extra:])
    ops = _oracle_ops(oracle)
    t = Keccak256Transcript()
    t.write_commitment(zeromorph.commit(pp, evals, ops))
    assert t.squeeze_challenges(case["num_vars"]) == point
    t.write_field_element(value)
    assert zeromorph.open(pp, evals, point, value, t, ops) == value
    assert t.into_proof() == proof
