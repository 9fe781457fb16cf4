"""oracle/pyoracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for oracle/liboracle_bn254.so (the C restatement of
/root/reference/plonkish_backend/src/util/arithmetic/msm.rs:84-181) with numpy
array plumbing.  Import only from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.

Array conventions (same bytes that cross the product's C ABI):
  scalars : np.uint64 [n, 4]  — Fr, Montgomery form, little-endian limbs
  bases   : np.uint64 [n, 8]  — G1Affine x||y, Montgomery Fq limbs, (0,0) = identity
  point   : np.uint64 [8]     — affine result
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_bn254.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bn254_oracle.c")
    hdr = os.path.join(_HERE, "bn254_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        _lib.oracle_fe_mul.argtypes = [ci, vp, vp, vp]
        _lib.oracle_fe_add.argtypes = [ci, vp, vp, vp]
        _lib.oracle_fe_sub.argtypes = [ci, vp, vp, vp]
        _lib.oracle_fe_inv.argtypes = [ci, vp, vp]
        _lib.oracle_fe_from_canonical.argtypes = [ci, vp, vp]
        _lib.oracle_fe_to_canonical.argtypes = [ci, vp, vp]
        _lib.oracle_fe_from_canonical_n.argtypes = [ci, vp, sz, vp]
        _lib.oracle_fr_div_linear.argtypes = [vp, sz, vp, vp, vp]
        _lib.oracle_bh_iter.argtypes = [sz, vp]
        _lib.oracle_permutation_z_polys.argtypes = [sz, vp, vp, sz, sz, vp, vp, vp]
        _lib.oracle_permutation_z_polys_mt.argtypes = [sz, vp, vp, sz, sz, vp, vp, ci, vp]
        _lib.oracle_fe_to_canonical_n.argtypes = [ci, vp, sz, vp]
        _lib.oracle_g1_generator.argtypes = [vp]
        _lib.oracle_g1_is_on_curve.argtypes = [vp]
        _lib.oracle_g1_is_on_curve.restype = ci
        _lib.oracle_g1_to_affine.argtypes = [vp, vp]
        _lib.oracle_g1_scalar_mul.argtypes = [vp, vp, vp]
        _lib.oracle_g1_transcript_bytes.argtypes = [vp, vp]
        _lib.oracle_g1_transcript_bytes.restype = ci
        _lib.oracle_window_size.argtypes = [sz]
        _lib.oracle_window_size.restype = sz
        _lib.oracle_windowed_scalar.argtypes = [sz, sz, sz, vp]
        _lib.oracle_windowed_scalar.restype = sz
        _lib.oracle_variable_base_msm.argtypes = [vp, vp, sz, ci, vp]
        _lib.oracle_msm_naive.argtypes = [vp, vp, sz, vp]
        _lib.oracle_known_dlog_bases.argtypes = [vp, vp, sz, ci, vp]
        _lib.oracle_known_dlog_answer.argtypes = [vp, vp, vp, sz, vp]
        _lib.oracle_quotients.argtypes = [vp, vp, sz, vp, vp]
        _lib.oracle_fr_linear_combination.argtypes = [vp, vp, sz, sz, vp]
        _lib.oracle_kzg_eq_scalars.argtypes = [vp, sz, vp]
        _lib.oracle_fixed_base_msm.argtypes = [vp, sz, vp, sz, ci, vp]
        _lib.oracle_sumcheck_round.argtypes = [vp, sz, sz, vp, vp, vp, sz, ci, sz, vp]
        _lib.oracle_fix_var.argtypes = [vp, sz, vp, vp]
        _lib.oracle_sumcheck_round_mt.argtypes = [vp, sz, sz, vp, vp, vp, sz, ci, sz, ci, vp]
        _lib.oracle_fix_var_mt.argtypes = [vp, sz, vp, ci, vp]
        _lib.oracle_fr_vec_op.argtypes = [ci, vp, vp, sz, ci, vp]
        _lib.oracle_fr_affine.argtypes = [vp, vp, vp, sz, vp, vp, sz, ci, vp]
        _lib.oracle_keccak256.argtypes = [ctypes.c_char_p, sz, vp]
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def int_to_limbs(v: int) -> np.ndarray:
    return np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint64).copy()


def limbs_to_int(a) -> int:
    return int.from_bytes(_u64(a).tobytes(), "little")


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def fe_op(op: str, which: int, a, b=None) -> np.ndarray:
    a = _u64(a)
    out = np.zeros(4, dtype=np.uint64)
    fn = getattr(lib(), f"oracle_fe_{op}")
    if b is None:
        fn(which, _ptr(a), _ptr(out))
    else:
        b = _u64(b)
        fn(which, _ptr(a), _ptr(b), _ptr(out))
    return out


def to_canonical(which: int, a) -> np.ndarray:
    a = _u64(a).reshape(-1, 4)
    out = np.zeros_like(a)
    lib().oracle_fe_to_canonical_n(which, _ptr(a), a.shape[0], _ptr(out))
    return out


def from_canonical(which: int, c) -> np.ndarray:
    c = _u64(c).reshape(-1, 4)
    out = np.zeros_like(c)
    lib().oracle_fe_from_canonical_n(which, _ptr(c), c.shape[0], _ptr(out))
    return out


def generator() -> np.ndarray:
    out = np.zeros(8, dtype=np.uint64)
    lib().oracle_g1_generator(_ptr(out))
    return out


def is_on_curve(pt) -> bool:
    pt = _u64(pt)
    return bool(lib().oracle_g1_is_on_curve(_ptr(pt)))


def scalar_mul(base, k_canonical: int) -> np.ndarray:
    base = _u64(base)
    k = int_to_limbs(k_canonical)
    jac = np.zeros(12, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    lib().oracle_g1_scalar_mul(_ptr(base), _ptr(k), _ptr(jac))
    lib().oracle_g1_to_affine(_ptr(jac), _ptr(out))
    return out


def transcript_bytes(pt) -> bytes:
    pt = _u64(pt)
    out = np.zeros(64, dtype=np.uint8)
    if lib().oracle_g1_transcript_bytes(_ptr(pt), _ptr(out)) != 0:
        raise ValueError("identity has no coordinates")
    return out.tobytes()


def variable_base_msm(scalars, bases, num_threads: int | None = None) -> np.ndarray:
    """msm.rs:84-115 followed by the callers' to_affine(); returns np.uint64[8]."""
    scalars = _u64(scalars).reshape(-1, 4)
    bases = _u64(bases).reshape(-1, 8)
    assert scalars.shape[0] == bases.shape[0]  # msm.rs:90
    if num_threads is None:
        num_threads = host_threads()
    jac = np.zeros(12, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    lib().oracle_variable_base_msm(_ptr(scalars), _ptr(bases), scalars.shape[0], int(num_threads), _ptr(jac))
    lib().oracle_g1_to_affine(_ptr(jac), _ptr(out))
    return out


def msm_naive(scalars, bases) -> np.ndarray:
    scalars = _u64(scalars).reshape(-1, 4)
    bases = _u64(bases).reshape(-1, 8)
    jac = np.zeros(12, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    lib().oracle_msm_naive(_ptr(scalars), _ptr(bases), scalars.shape[0], _ptr(jac))
    lib().oracle_g1_to_affine(_ptr(jac), _ptr(out))
    return out


def known_dlog_bases(a: int, d: int, n: int, num_threads: int | None = None) -> np.ndarray:
    if num_threads is None:
        num_threads = host_threads()
    out = np.zeros((n, 8), dtype=np.uint64)
    la, ld = int_to_limbs(a), int_to_limbs(d)
    lib().oracle_known_dlog_bases(_ptr(la), _ptr(ld), n, int(num_threads), _ptr(out))
    return out


def known_dlog_answer(a: int, d: int, scalars) -> np.ndarray:
    scalars = _u64(scalars).reshape(-1, 4)
    out = np.zeros(8, dtype=np.uint64)
    la, ld = int_to_limbs(a), int_to_limbs(d)
    lib().oracle_known_dlog_answer(_ptr(la), _ptr(ld), _ptr(scalars), scalars.shape[0], _ptr(out))
    return out


def random_scalars(n: int, seed: int) -> np.ndarray:
    """Uniform-digit Fr elements in Montgomery form: three uniform u64 limbs and a
    top limb below r's top limb, so every value is a valid representation < r."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    out[:, 3] = rng.integers(0, 0x30644E72E131A029, size=n, dtype=np.uint64)
    return out


def window_size(num_scalars: int) -> int:
    return int(lib().oracle_window_size(num_scalars))


def quotients(evals, point):
    """pcs/multilinear.rs:72-107: returns ([q_0, ..., q_{k-1}] with q_i of 2^i Montgomery
    scalars, f(point) as Montgomery limbs)."""
    evals = _u64(evals).reshape(-1, 4)
    point = _u64(point).reshape(-1, 4)
    k = point.shape[0]
    assert evals.shape[0] == 1 << k
    q = np.zeros((1 << k, 4), dtype=np.uint64)
    value = np.zeros(4, dtype=np.uint64)
    lib().oracle_quotients(_ptr(evals), _ptr(point), k, _ptr(q), _ptr(value))
    return [q[1 << i: 2 << i].copy() for i in range(k)], value


def fr_linear_combination(polys, coeffs) -> np.ndarray:
    """pcs/multilinear.rs:203-213: sum_i coeffs[i] * polys[i]."""
    polys = [_u64(p).reshape(-1, 4) for p in polys]
    coeffs = _u64(coeffs).reshape(-1, 4)
    n = polys[0].shape[0]
    assert all(p.shape[0] == n for p in polys) and coeffs.shape[0] == len(polys)
    ptrs = (ctypes.c_void_p * len(polys))(*[p.ctypes.data for p in polys])
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().oracle_fr_linear_combination(ctypes.cast(ptrs, ctypes.c_void_p), _ptr(coeffs), len(polys), n, _ptr(out))
    return out


def fr_vec_op(op: str, a, b, num_threads: int | None = None) -> np.ndarray:
    """Element-wise a + b / a - b / a * b over [n, 4] Montgomery Fr arrays on the host threads."""
    a, b = _u64(a).reshape(-1, 4), _u64(b).reshape(-1, 4)
    assert a.shape == b.shape
    out = np.zeros_like(a)
    lib().oracle_fr_vec_op({"add": 0, "sub": 1, "mul": 2}[op], _ptr(a), _ptr(b), a.shape[0], int(num_threads or host_threads()), _ptr(out))
    return out


def fr_affine(n: int, polys=(), coeffs=None, rows=None, constant=None, id_coeff=None, num_threads: int | None = None) -> np.ndarray:
    """out[j] = constant + id_coeff * j + sum_i coeffs[i] * polys[i][rows[i][j] if rows[i] is not None else j]: one
    linear factor of a zero-check expression as an explicit table (preprocessor.rs:153-165; classic.rs:92, 105-125).
    rows[i]: uint32 row map of a rotated query (BooleanHypercube::rotation_map) or None."""
    polys = [_u64(p).reshape(-1, 4) for p in polys]
    count = len(polys)
    assert all(p.shape[0] == n for p in polys)
    cs = _u64(coeffs).reshape(count, 4) if count else np.zeros((1, 4), dtype=np.uint64)
    ptrs = (ctypes.c_void_p * max(count, 1))(*[p.ctypes.data for p in polys])
    maps = [None if rows is None or rows[i] is None else np.ascontiguousarray(rows[i], dtype=np.uint32) for i in range(count)]
    rptrs = (ctypes.c_void_p * max(count, 1))(*[None if m is None else m.ctypes.data for m in maps])
    const = None if constant is None else _u64(constant).reshape(4)
    idc = None if id_coeff is None else _u64(id_coeff).reshape(4)
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().oracle_fr_affine(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(rptrs, ctypes.c_void_p), _ptr(cs), count,
                           None if const is None else _ptr(const), None if idc is None else _ptr(idc), n, int(num_threads or host_threads()), _ptr(out))
    return out


def kzg_eq_scalars(ss):
    """pcs/multilinear/kzg.rs:174-193: [eqs[0], ..., eqs[num_vars]] as Montgomery scalars."""
    ss = _u64(ss).reshape(-1, 4)
    k = ss.shape[0]
    out = np.zeros(((2 << k) - 1, 4), dtype=np.uint64)
    lib().oracle_kzg_eq_scalars(_ptr(ss), k, _ptr(out))
    return [out[(1 << i) - 1: (2 << i) - 1].copy() for i in range(k + 1)]


def fixed_base_msm(base, scalars, window: int | None = None, num_threads: int | None = None) -> np.ndarray:
    """msm.rs:16-31, 50-81 + batch_normalize: out[i] = scalars[i] * base (affine, [n, 8])."""
    base = _u64(base).reshape(8)
    scalars = _u64(scalars).reshape(-1, 4)
    n = scalars.shape[0]
    if window is None:
        window = window_size(n)
    if num_threads is None:
        num_threads = host_threads()
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().oracle_fixed_base_msm(_ptr(base), int(window), _ptr(scalars), n, int(num_threads), _ptr(out))
    return out


def flatten_terms(terms):
    """[(coeff limbs[4], [poly indices])] -> (coeffs [T,4] uint64, offsets [T+1] uint32, polys uint32)."""
    coeffs = np.stack([_u64(c).reshape(4) for c, _ in terms])
    offsets = np.zeros(len(terms) + 1, dtype=np.uint32)
    flat = []
    for t, (_, idx) in enumerate(terms):
        flat.extend(int(i) for i in idx)
        offsets[t + 1] = len(flat)
    return coeffs, offsets, np.array(flat if flat else [0], dtype=np.uint32)


def sumcheck_round(polys, terms, common: int = -1, num_threads: int = 1) -> np.ndarray:
    """piop/sum_check/classic/eval.rs:101-131: evaluations of the round polynomial at X = 1..degree."""
    polys = [_u64(p).reshape(-1, 4) for p in polys]
    n = polys[0].shape[0]
    assert n >= 2 and all(p.shape[0] == n for p in polys)
    coeffs, offsets, flat = flatten_terms(terms)
    degree = max(len(idx) for _, idx in terms) + (1 if common >= 0 else 0)
    degree = max(degree, 1)
    ptrs = (ctypes.c_void_p * len(polys))(*[p.ctypes.data for p in polys])
    out = np.zeros((degree, 4), dtype=np.uint64)
    if num_threads > 1:
        lib().oracle_sumcheck_round_mt(ctypes.cast(ptrs, ctypes.c_void_p), len(polys), n // 2, _ptr(coeffs), _ptr(offsets), _ptr(flat),
                                       len(terms), int(common), degree, int(num_threads), _ptr(out))
        return out
    lib().oracle_sumcheck_round(ctypes.cast(ptrs, ctypes.c_void_p), len(polys), n // 2, _ptr(coeffs), _ptr(offsets), _ptr(flat),
                                len(terms), int(common), degree, _ptr(out))
    return out


def fix_var(evals, x, num_threads: int = 1) -> np.ndarray:
    """MultilinearPolynomial::fix_var (poly/multilinear.rs:179-189)."""
    evals = _u64(evals).reshape(-1, 4)
    x = _u64(x).reshape(4)
    out = np.zeros((evals.shape[0] // 2, 4), dtype=np.uint64)
    if num_threads > 1:
        lib().oracle_fix_var_mt(_ptr(evals), evals.shape[0], _ptr(x), int(num_threads), _ptr(out))
    else:
        lib().oracle_fix_var(_ptr(evals), evals.shape[0], _ptr(x), _ptr(out))
    return out


def evaluate_multilinear(evals, point, num_threads: int = 1) -> np.ndarray:
    """MultilinearPolynomial::evaluate (poly/multilinear.rs:142-170 in effect): fix every variable in turn, lowest first.
    evals [2^k, 4], point [k, 4] -> [4] (Montgomery)."""
    cur = _u64(evals).reshape(-1, 4)
    pt = _u64(point).reshape(-1, 4)
    assert cur.shape[0] == 1 << pt.shape[0]
    for i in range(pt.shape[0]):
        cur = fix_var(cur, pt[i], num_threads if cur.shape[0] >= 1 << 14 else 1)
    return cur[0].copy()


def fr_div_linear(coeffs, z):
    """UnivariatePolynomial::div_rem by (X - z): returns (quotient [n-1, 4], remainder [4]) in Montgomery form."""
    coeffs = _u64(coeffs).reshape(-1, 4)
    z = _u64(z).reshape(4)
    n = coeffs.shape[0]
    q = np.zeros((max(n - 1, 0), 4), dtype=np.uint64)
    rem = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_div_linear(_ptr(coeffs), n, _ptr(z), _ptr(q) if n > 1 else None, _ptr(rem))
    return q, rem


def bh_iter(num_vars: int) -> np.ndarray:
    """BooleanHypercube::iter (util/arithmetic/bh.rs:118-125)."""
    out = np.zeros(1 << num_vars, dtype=np.uint32)
    lib().oracle_bh_iter(num_vars, _ptr(out))
    return out


def permutation_z_polys(num_chunks: int, values, sigmas, beta, gamma, num_threads: int = 1):
    """backend/hyperplonk/prover.rs:252-345 -> list of num_chunks [2^k, 4] arrays."""
    values = [_u64(v).reshape(-1, 4) for v in values]
    sigmas = [_u64(v).reshape(-1, 4) for v in sigmas]
    n = values[0].shape[0]
    k = n.bit_length() - 1
    vp_ = (ctypes.c_void_p * len(values))(*[v.ctypes.data for v in values])
    sp_ = (ctypes.c_void_p * len(sigmas))(*[v.ctypes.data for v in sigmas])
    out = np.zeros((num_chunks * n, 4), dtype=np.uint64)
    lib().oracle_permutation_z_polys_mt(num_chunks, ctypes.cast(vp_, ctypes.c_void_p), ctypes.cast(sp_, ctypes.c_void_p), len(values), k,
                                        _ptr(_u64(beta).reshape(4)), _ptr(_u64(gamma).reshape(4)), int(num_threads), _ptr(out))
    return [out[c * n:(c + 1) * n].copy() for c in range(num_chunks)]


def keccak256(data: bytes) -> bytes:
    """Keccak256 with the original padding (the hash of Keccak256Transcript, util/transcript.rs:100-131)."""
    out = np.zeros(32, dtype=np.uint8)
    lib().oracle_keccak256(bytes(data), len(data), _ptr(out))
    return out.tobytes()
