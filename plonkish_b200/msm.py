"""Host-side mirror of the reference's MSM interface, over the C ABI.

Mirrors ``plonkish_backend::util::arithmetic::variable_base_msm``
(/root/reference/plonkish_backend/src/util/arithmetic/msm.rs:84-115): same
argument meaning (scalars, bases of equal length), same error behaviour for a
length mismatch (the reference's ``assert_eq!`` at msm.rs:90 -> AssertionError),
and the affine value the reference's callers take with ``.into()`` /
``.to_affine()`` (pcs/multilinear/kzg.rs:255).

Array conventions (the bytes that cross the ABI, see include/plonkish_cuda.h):
  scalars  uint64 [n, 4]  bn256::Fr, Montgomery form, little-endian limbs
  bases    uint64 [n, 8]  bn256::G1Affine x||y (Montgomery Fq), (0,0) = identity
  result   uint64 [8]     G1Affine
Everything is computed on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np

from . import _lib

SCALAR_BYTES = 32
AFFINE_BYTES = 64
XYZZ_BYTES = 128


def _as_u64(a, width: int, what: str) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.uint64)
    if arr.ndim == 1 and arr.size % width == 0:
        arr = arr.reshape(-1, width)
    if arr.ndim != 2 or arr.shape[1] != width:
        raise ValueError(f"{what} must have shape [n, {width}] of uint64 limbs, got {arr.shape}")
    return arr


class G1Bases:
    """Bases resident on one GPU (the SRS slice a ProverParam holds:
    MultilinearKzgProverParam.eqs[k], pcs/multilinear/kzg.rs:55-77;
    UnivariateKzgProverParam.powers_of_s_g1, pcs/univariate/kzg.rs:24-30).

    `bases` is a host [n, 8] uint64 array or a torch CUDA tensor of the same shape.
    Registration expands the slice into a table of window multiples when it fits in
    free HBM (mode 0), keeps plain bases (mode 1) or insists on the table (mode 2)."""

    PLAIN, TABLE = 1, 2

    def __init__(self, bases, device: int = 0, mode: int = 0):
        handle = ctypes.c_uint64(0)
        if hasattr(bases, "is_cuda"):
            assert bases.is_cuda and bases.is_contiguous()
            self.n = bases.numel() * bases.element_size() // AFFINE_BYTES
            self.device = bases.device.index or 0
            self._keepalive = bases  # mode 1 borrows the tensor's memory
            rc = _lib.lib().plonkish_cuda_bases_register_device(self.device, bases.data_ptr(), self.n, mode, ctypes.byref(handle))
            _lib.check(rc, "plonkish_cuda_bases_register_device")
        else:
            arr = _as_u64(bases, 8, "bases")
            self.n = arr.shape[0]
            self.device = device
            if mode == 0:
                rc = _lib.lib().plonkish_cuda_bases_register(device, arr.ctypes.data, self.n, ctypes.byref(handle))
                _lib.check(rc, "plonkish_cuda_bases_register")
            else:
                import torch

                dev_t = torch.from_numpy(arr.view(np.int64)).to(torch.device("cuda", device))
                self._keepalive = dev_t
                rc = _lib.lib().plonkish_cuda_bases_register_device(device, dev_t.data_ptr(), self.n, mode, ctypes.byref(handle))
                _lib.check(rc, "plonkish_cuda_bases_register_device")
        self.handle = handle.value

    @classmethod
    def _adopt(cls, handle: int, n: int, device: int) -> "G1Bases":
        """Wrap a handle the library produced itself (plonkish_cuda_kzg_setup_eqs_bn254)."""
        self = cls.__new__(cls)
        self.handle, self.n, self.device = int(handle), int(n), int(device)
        return self

    def __len__(self) -> int:
        return self.n

    def to_host(self, offset: int = 0, n: Optional[int] = None) -> np.ndarray:
        """The resident bases back on the host ([n, 8] uint64)."""
        n = self.n - offset if n is None else n
        out = np.zeros((n, 8), dtype=np.uint64)
        _lib.check(_lib.lib().plonkish_cuda_bases_read(self.handle, offset, n, out.ctypes.data), "plonkish_cuda_bases_read")
        return out

    def release(self) -> None:
        if self.handle:
            _lib.check(_lib.lib().plonkish_cuda_bases_release(self.handle), "plonkish_cuda_bases_release")
            self.handle = 0


def cached_bases(bases, device: int = 0) -> "G1Bases":
    """The handle the Rust shim's variable_base_msm gets for a borrowed `&[G1Affine]` (msm.rs:84-87): looked up by
    address in the library's cache and trusted only after its content fingerprint matches; see
    plonkish_cuda_bases_cached.  `bases` must be a C-contiguous [n, 8] uint64 array (its address is the key).
    The returned object does not own the handle: the cache does (evict with cache_evict)."""
    arr = bases if isinstance(bases, np.ndarray) and bases.dtype == np.uint64 and bases.flags.c_contiguous else None
    assert arr is not None and arr.ndim == 2 and arr.shape[1] == 8, "cached_bases needs a C-contiguous [n, 8] uint64 array"
    handle = ctypes.c_uint64(0)
    _lib.check(_lib.lib().plonkish_cuda_bases_cached(device, arr.ctypes.data, arr.shape[0], ctypes.byref(handle)), "plonkish_cuda_bases_cached")
    return G1Bases._adopt(handle.value, arr.shape[0], device)


def cache_evict(bases) -> None:
    _lib.check(_lib.lib().plonkish_cuda_bases_cache_evict(bases.ctypes.data), "plonkish_cuda_bases_cache_evict")


def cache_limit(max_bytes: int) -> None:
    _lib.check(_lib.lib().plonkish_cuda_bases_cache_limit(max_bytes), "plonkish_cuda_bases_cache_limit")


def cache_stats() -> dict:
    out = (ctypes.c_size_t * 2)()
    _lib.check(_lib.lib().plonkish_cuda_bases_cache_stats(out), "plonkish_cuda_bases_cache_stats")
    return {"entries": int(out[0]), "bytes": int(out[1])}


def timer_config(mode: int, depth: int = 0) -> None:
    """0 off, 1 stderr, 2 stdout: the reference's `variable_base_msm-{n}` timer lines (msm.rs:92) in perf_trace format."""
    _lib.check(_lib.load().plonkish_cuda_timer_config(mode, depth), "plonkish_cuda_timer_config")


def staged_bytes() -> int:
    return int(_lib.load().plonkish_cuda_staged_bytes())


def staging_rate_gbps() -> float:
    """GB/s the pinned staging ring sustained over its recent large uploads (0.0 before the first)."""
    return float(_lib.load().plonkish_cuda_staging_rate_gbps())


class ResidentScalars:
    """A polynomial's evaluations (n x bn256::Fr, Montgomery) kept in HBM between its commit
    and its opening — poly.evals() of pcs/multilinear/kzg.rs:255,291 without the re-upload."""

    def __init__(self, scalars, device: int = 0):
        arr = _as_u64(scalars, 4, "scalars")
        handle = ctypes.c_uint64(0)
        _lib.check(_lib.lib().plonkish_cuda_scalars_register(device, arr.ctypes.data, arr.shape[0], ctypes.byref(handle)),
                   "plonkish_cuda_scalars_register")
        self.handle, self.n, self.device = handle.value, arr.shape[0], device

    @classmethod
    def _adopt(cls, handle: int, n: int, device: int) -> "ResidentScalars":
        self = cls.__new__(cls)
        self.handle, self.n, self.device = int(handle), int(n), int(device)
        return self

    def __len__(self) -> int:
        return self.n

    def to_host(self, offset: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.n - offset if n is None else n
        out = np.zeros((n, 4), dtype=np.uint64)
        _lib.check(_lib.lib().plonkish_cuda_scalars_read(self.handle, offset, n, out.ctypes.data), "plonkish_cuda_scalars_read")
        return out

    def release(self) -> None:
        if self.handle:
            _lib.check(_lib.lib().plonkish_cuda_scalars_release(self.handle), "plonkish_cuda_scalars_release")
            self.handle = 0


class ShardedG1Bases:
    """Bases split across n_gpus devices the way msm.rs:101-107 chunks them."""

    def __init__(self, bases, n_gpus: int):
        arr = _as_u64(bases, 8, "bases")
        self.n = arr.shape[0]
        self.n_gpus = n_gpus
        handle = ctypes.c_uint64(0)
        _lib.check(
            _lib.lib().plonkish_cuda_bases_register_sharded(n_gpus, arr.ctypes.data, self.n, ctypes.byref(handle)),
            "plonkish_cuda_bases_register_sharded",
        )
        self.handle = handle.value

    @classmethod
    def from_device(cls, shards, n: int, mode: int = 0) -> "ShardedG1Bases":
        """shards[g]: CUDA tensor on device g holding points [g*ceil(n/G), (g+1)*ceil(n/G)) of the slice."""
        self = cls.__new__(cls)
        self.n, self.n_gpus = int(n), len(shards)
        ptrs = (ctypes.c_void_p * len(shards))(*[t.data_ptr() if t is not None and t.numel() else None for t in shards])
        handle = ctypes.c_uint64(0)
        _lib.check(_lib.lib().plonkish_cuda_bases_register_sharded_device(len(shards), ctypes.cast(ptrs, ctypes.c_void_p), self.n, mode, ctypes.byref(handle)),
                   "plonkish_cuda_bases_register_sharded_device")
        self.handle = handle.value
        self._keepalive = list(shards) if mode == 1 else None  # plain mode borrows the tensors
        return self

    def release(self) -> None:
        if self.handle:
            _lib.check(_lib.lib().plonkish_cuda_bases_release(self.handle), "plonkish_cuda_bases_release")
            self.handle = 0


def variable_base_msm(scalars, bases, n_gpus: int = 1) -> np.ndarray:
    """sum_i scalars[i] * bases[i] on the GPU; returns the affine point, uint64[8].

    `bases` is an [n, 8] array, a G1Bases / ShardedG1Bases, or — like the
    reference's iterator-of-references callers (pcs/univariate/kzg.rs:346) — a
    sequence of separate 8-limb arrays, in which case `scalars` must be a sequence
    of 4-limb arrays too and the gather entry point is used.
    """
    lib = _lib.lib()
    out = np.zeros(8, dtype=np.uint64)
    if isinstance(scalars, ResidentScalars):
        assert isinstance(bases, G1Bases), "resident scalars need resident bases"
        assert scalars.n <= bases.n, "more scalars than registered bases"  # msm.rs:90
        _lib.check(lib.plonkish_cuda_msm_bn254_g1_resident(scalars.handle, bases.handle, scalars.n, out.ctypes.data),
                   "plonkish_cuda_msm_bn254_g1_resident")
        return out
    if isinstance(bases, (G1Bases, ShardedG1Bases)):
        sc = _as_u64(scalars, 4, "scalars")
        assert sc.shape[0] <= bases.n, "more scalars than registered bases"  # msm.rs:90
        if isinstance(bases, ShardedG1Bases):
            assert sc.shape[0] == bases.n, "scalars and sharded bases differ in length"  # msm.rs:90
            rc = lib.plonkish_cuda_msm_bn254_g1_multi(bases.n_gpus, sc.ctypes.data, None, bases.handle, sc.shape[0], out.ctypes.data)
            _lib.check(rc, "plonkish_cuda_msm_bn254_g1_multi")
        else:
            rc = lib.plonkish_cuda_msm_bn254_g1(sc.ctypes.data, None, bases.handle, sc.shape[0], out.ctypes.data)
            _lib.check(rc, "plonkish_cuda_msm_bn254_g1")
        return out
    if isinstance(bases, (list, tuple)) and not isinstance(scalars, np.ndarray):
        return _variable_base_msm_gather(scalars, bases)
    sc = _as_u64(scalars, 4, "scalars")
    bs = _as_u64(bases, 8, "bases")
    assert sc.shape[0] == bs.shape[0], "scalars and bases differ in length"  # msm.rs:90
    if n_gpus > 1:
        rc = lib.plonkish_cuda_msm_bn254_g1_multi(n_gpus, sc.ctypes.data, bs.ctypes.data, 0, sc.shape[0], out.ctypes.data)
        _lib.check(rc, "plonkish_cuda_msm_bn254_g1_multi")
    else:
        rc = lib.plonkish_cuda_msm_bn254_g1(sc.ctypes.data, bs.ctypes.data, 0, sc.shape[0], out.ctypes.data)
        _lib.check(rc, "plonkish_cuda_msm_bn254_g1")
    return out


def variable_base_msm_batch(scalars_list: Sequence, bases: "G1Bases") -> np.ndarray:
    """One MSM per entry of scalars_list (equal lengths) against one resident base slice ->
    [count, 8] affine points.  The batch_commit loop of pcs/multilinear/kzg.rs:259-274 with the
    upload of polynomial j+1 overlapped with the MSM of polynomial j."""
    arrs = [_as_u64(s, 4, "scalars") for s in scalars_list]
    count = len(arrs)
    out = np.zeros((count, 8), dtype=np.uint64)
    if count == 0:
        return out
    n = arrs[0].shape[0]
    assert all(a.shape[0] == n for a in arrs), "batched polynomials must have the same size"
    assert n <= bases.n, "more scalars than registered bases"  # msm.rs:90
    ptrs = (ctypes.c_void_p * count)(*[a.ctypes.data for a in arrs])
    rc = _lib.lib().plonkish_cuda_msm_bn254_g1_batch(ptrs, count, bases.handle, n, out.ctypes.data)
    _lib.check(rc, "plonkish_cuda_msm_bn254_g1_batch")
    return out


def variable_base_msm_batch_keep(scalars_list: Sequence, bases: "G1Bases"):
    """variable_base_msm_batch that also leaves every polynomial resident: returns
    ([count, 8] commitments, [ResidentScalars, ...])."""
    arrs = [_as_u64(s, 4, "scalars") for s in scalars_list]
    assert arrs, "empty batch"
    n = arrs[0].shape[0]
    assert n and all(a.shape[0] == n for a in arrs), "batch entries must have one non-zero length"
    assert n <= bases.n, "more scalars than registered bases"  # msm.rs:90
    ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    out = np.zeros((len(arrs), 8), dtype=np.uint64)
    handles = np.zeros(len(arrs), dtype=np.uint64)
    rc = _lib.lib().plonkish_cuda_msm_bn254_g1_batch_keep(ctypes.cast(ptrs, ctypes.c_void_p), len(arrs), bases.handle, n,
                                                          out.ctypes.data, handles.ctypes.data)
    _lib.check(rc, "plonkish_cuda_msm_bn254_g1_batch_keep")
    return out, [ResidentScalars._adopt(h, n, bases.device) for h in handles]


def fr_linear_combination(polys: Sequence["ResidentScalars"], coeffs) -> "ResidentScalars":
    """sum_i coeffs[i] * polys[i] as a new resident polynomial (the g_prime merge,
    pcs/multilinear.rs:203-213).  coeffs: [count, 4] Montgomery Fr."""
    cs = _as_u64(coeffs, 4, "coeffs")
    assert len(polys) == cs.shape[0] and len(polys) > 0
    n = min(p.n for p in polys)
    hs = np.array([p.handle for p in polys], dtype=np.uint64)
    out = ctypes.c_uint64(0)
    rc = _lib.lib().plonkish_cuda_fr_linear_combination(hs.ctypes.data, cs.ctypes.data, len(polys), n, ctypes.byref(out))
    _lib.check(rc, "plonkish_cuda_fr_linear_combination")
    return ResidentScalars._adopt(out.value, n, polys[0].device)


def permutation_z_polys(num_chunks: int, values: Sequence["ResidentScalars"], sigmas: Sequence["ResidentScalars"], beta, gamma):
    """permutation_z_polys (backend/hyperplonk/prover.rs:252-345) on resident polynomials: values[i] is the witness column of
    permutation polynomial i, sigmas[i] the permutation polynomial; beta, gamma Montgomery limbs.  Returns num_chunks
    ResidentScalars (the z polynomials committed at backend/hyperplonk.rs:251-252)."""
    assert len(values) == len(sigmas) and len(values) > 0
    n = values[0].n
    num_vars = n.bit_length() - 1
    vh = np.array([p.handle for p in values], dtype=np.uint64)
    sh = np.array([p.handle for p in sigmas], dtype=np.uint64)
    b = np.ascontiguousarray(beta, dtype=np.uint64).reshape(4)
    g = np.ascontiguousarray(gamma, dtype=np.uint64).reshape(4)
    out = np.zeros(num_chunks, dtype=np.uint64)
    rc = _lib.lib().plonkish_cuda_permutation_z_polys_bn254(vh.ctypes.data, sh.ctypes.data, len(values), num_chunks, num_vars, b.ctypes.data, g.ctypes.data,
                                                            out.ctypes.data)
    _lib.check(rc, "plonkish_cuda_permutation_z_polys_bn254")
    return [ResidentScalars._adopt(h, n, values[0].device) for h in out]


def fr_affine_table(num_vars: int, polys: Sequence["ResidentScalars"] = (), coeffs=None, rotations: Optional[Sequence[int]] = None, constant=None,
                    identity_coeff=None, sparse_rows: Sequence[int] = (), sparse_values=None, device: int = 0) -> "ResidentScalars":
    """One table of a compiled sum-check expression (plonkish_cuda_fr_affine_table):
    out[b] = constant + identity_coeff * b + sum_i coeffs[i] * polys[i][rotate(b, rotations[i])], then out[rows[j]] += values[j].
    coeffs / constant / identity_coeff / sparse_values are Montgomery limbs; rotate is BooleanHypercube::rotate
    (util/arithmetic/bh.rs:104-121).  What the reference's sum check keeps implicit (piop/sum_check/classic.rs:40-75, 104-126)."""
    count = len(polys)
    hs = np.array([p.handle for p in polys] or [0], dtype=np.uint64)
    cs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(count, 4) if count else np.zeros((1, 4), dtype=np.uint64)
    rot = np.array(list(rotations) if rotations is not None else [0] * count or [0], dtype=np.int32)
    assert not count or rot.shape[0] == count, "one rotation per polynomial"
    const = None if constant is None else np.ascontiguousarray(constant, dtype=np.uint64).reshape(4)
    idc = None if identity_coeff is None else np.ascontiguousarray(identity_coeff, dtype=np.uint64).reshape(4)
    rows = np.array(list(sparse_rows) or [0], dtype=np.uint64)
    vals = np.ascontiguousarray(sparse_values, dtype=np.uint64).reshape(len(sparse_rows), 4) if len(sparse_rows) else np.zeros((1, 4), dtype=np.uint64)
    out = ctypes.c_uint64(0)
    rc = _lib.lib().plonkish_cuda_fr_affine_table(device if not count else polys[0].device, num_vars, hs.ctypes.data, rot.ctypes.data, cs.ctypes.data, count,
                                                  None if const is None else const.ctypes.data, None if idc is None else idc.ctypes.data,
                                                  rows.ctypes.data, vals.ctypes.data, len(sparse_rows), ctypes.byref(out))
    _lib.check(rc, "plonkish_cuda_fr_affine_table")
    return ResidentScalars._adopt(out.value, 1 << num_vars, device if not count else polys[0].device)


def fr_expression_table(polys: Sequence["ResidentScalars"], terms, common: int = -1) -> "ResidentScalars":
    """A compiled expression on every row: out[b] = (sum_t coeff_t * prod_j polys[idx_t,j][b]) (* polys[common][b]);
    terms = [(coeff limbs, [poly indices])] as for the sum check.  lookup_compressed_poly (prover.rs:79-137) is one call."""
    n = polys[0].n
    num_vars = n.bit_length() - 1
    handles = np.array([p.handle for p in polys], dtype=np.uint64)
    coeffs = np.stack([_as_u64(c, 4, "coeff").reshape(4) for c, _ in terms])
    offsets = np.zeros(len(terms) + 1, dtype=np.uint32)
    flat: List[int] = []
    for t, (_, idx) in enumerate(terms):
        flat.extend(int(i) for i in idx)
        offsets[t + 1] = len(flat)
    flat_arr = np.array(flat if flat else [0], dtype=np.uint32)
    out = ctypes.c_uint64(0)
    rc = _lib.lib().plonkish_cuda_fr_expression_table(handles.ctypes.data, len(polys), num_vars, coeffs.ctypes.data, offsets.ctypes.data, flat_arr.ctypes.data,
                                                      len(terms), int(common), ctypes.byref(out))
    _lib.check(rc, "plonkish_cuda_fr_expression_table")
    return ResidentScalars._adopt(out.value, n, polys[0].device)


def lookup_m_poly(compressed_input: "ResidentScalars", compressed_table: "ResidentScalars") -> "ResidentScalars":
    """lookup_m_poly (backend/hyperplonk/prover.rs:145-192) on resident polynomials; PlonkishCudaError("Invalid lookup input")
    when an input value is not in the table."""
    out = ctypes.c_uint64(0)
    _lib.check(_lib.lib().plonkish_cuda_lookup_m_poly_bn254(compressed_input.handle, compressed_table.handle, ctypes.byref(out)), "plonkish_cuda_lookup_m_poly_bn254")
    return ResidentScalars._adopt(out.value, compressed_input.n, compressed_input.device)


def lookup_h_poly(compressed_input: "ResidentScalars", compressed_table: "ResidentScalars", m: "ResidentScalars", gamma) -> "ResidentScalars":
    """lookup_h_poly (backend/hyperplonk/prover.rs:206-250): 1 / (gamma + input) - m / (gamma + table); gamma Montgomery limbs."""
    g = np.ascontiguousarray(gamma, dtype=np.uint64).reshape(4)
    out = ctypes.c_uint64(0)
    rc = _lib.lib().plonkish_cuda_lookup_h_poly_bn254(compressed_input.handle, compressed_table.handle, m.handle, g.ctypes.data, ctypes.byref(out))
    _lib.check(rc, "plonkish_cuda_lookup_h_poly_bn254")
    return ResidentScalars._adopt(out.value, compressed_input.n, compressed_input.device)


def fr_evaluate(poly: "ResidentScalars", points) -> np.ndarray:
    """MultilinearPolynomial::evaluate (poly/multilinear.rs:137-156) of a resident polynomial at each of `points`
    ([count, num_vars, 4] Montgomery Fr); returns [count, 4] Montgomery limbs."""
    num_vars = poly.n.bit_length() - 1
    pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, max(num_vars, 1), 4) if num_vars else np.zeros((len(points), 1, 4), dtype=np.uint64)
    count = pts.shape[0]
    out = np.zeros((count, 4), dtype=np.uint64)
    _lib.check(_lib.lib().plonkish_cuda_fr_evaluate(poly.handle, pts.ctypes.data, num_vars, count, out.ctypes.data), "plonkish_cuda_fr_evaluate")
    return out


def fr_div_linear(poly: "ResidentScalars", z):
    """poly / (X - z) on resident coefficients (poly/univariate.rs:144-168 for a linear divisor): returns
    (quotient as ResidentScalars of the same length, top coefficient zero; remainder = poly(z) as Montgomery limbs [4])."""
    zz = np.ascontiguousarray(z, dtype=np.uint64).reshape(4)
    out = ctypes.c_uint64(0)
    rem = np.zeros(4, dtype=np.uint64)
    _lib.check(_lib.lib().plonkish_cuda_fr_div_linear(poly.handle, zz.ctypes.data, ctypes.byref(out), rem.ctypes.data), "plonkish_cuda_fr_div_linear")
    return ResidentScalars._adopt(out.value, poly.n, poly.device), rem


def fr_linear_combination_padded(polys: Sequence["ResidentScalars"], coeffs) -> "ResidentScalars":
    """sum_i coeffs[i] * polys[i] for univariate polynomials of different lengths (shorter ones count as zero past their
    last coefficient; `f += (scalar, q)`, poly/univariate.rs): a new resident polynomial of the longest length."""
    cs = _as_u64(coeffs, 4, "coeffs")
    assert len(polys) == cs.shape[0] and len(polys) > 0
    n = max(p.n for p in polys)
    hs = np.array([p.handle for p in polys], dtype=np.uint64)
    out = ctypes.c_uint64(0)
    rc = _lib.lib().plonkish_cuda_fr_linear_combination_padded(hs.ctypes.data, cs.ctypes.data, len(polys), n, ctypes.byref(out))
    _lib.check(rc, "plonkish_cuda_fr_linear_combination_padded")
    return ResidentScalars._adopt(out.value, n, polys[0].device)


def scalars_slice(scalars: "ResidentScalars", offset: int, n: int) -> "ResidentScalars":
    """A ResidentScalars on scalars[offset : offset + n] sharing the memory (kept alive until the slice is released)."""
    out = ctypes.c_uint64(0)
    _lib.check(_lib.lib().plonkish_cuda_scalars_slice(scalars.handle, offset, n, ctypes.byref(out)), "plonkish_cuda_scalars_slice")
    return ResidentScalars._adopt(out.value, n, scalars.device)


def fr_gemini_folds(poly: "ResidentScalars", point) -> "ResidentScalars":
    """The folds f_1 .. f_(k-1) of Gemini::open (pcs/multilinear/gemini.rs:98-108), packed: f_i (2^(k-i) values) at
    element offset 2^(k-i) of a new resident vector of 2^k scalars.  point: [k, 4] Montgomery Fr."""
    pt = _as_u64(point, 4, "point")
    k = pt.shape[0]
    assert poly.n == 1 << k, "point / polynomial size mismatch"
    out = ctypes.c_uint64(0)
    _lib.check(_lib.lib().plonkish_cuda_fr_gemini_folds(poly.handle, pt.ctypes.data, k, ctypes.byref(out)), "plonkish_cuda_fr_gemini_folds")
    return ResidentScalars._adopt(out.value, poly.n, poly.device)


def fr_quotients(poly: "ResidentScalars", point):
    """`quotients` (pcs/multilinear.rs:72-107) on a resident polynomial, kept in HBM: returns (packed quotients as
    ResidentScalars of 2^k scalars — quotient i, 2^i values, at element offset 2^i, element 0 zero —, f(point) as
    Montgomery limbs [4]).  point: [k, 4] Montgomery Fr."""
    pt = _as_u64(point, 4, "point") if len(point) else np.zeros((0, 4), dtype=np.uint64)
    k = pt.shape[0]
    assert poly.n == 1 << k, "point / polynomial size mismatch"  # multilinear.rs:77
    out = ctypes.c_uint64(0)
    value = np.zeros(4, dtype=np.uint64)
    rc = _lib.lib().plonkish_cuda_fr_quotients(poly.handle, pt.ctypes.data if k else None, k, ctypes.byref(out), value.ctypes.data)
    _lib.check(rc, "plonkish_cuda_fr_quotients")
    return ResidentScalars._adopt(out.value, poly.n, poly.device), value


def variable_base_msm_many_resident(scalars: "ResidentScalars", offsets: Sequence[int], bases_list: Sequence["G1Bases"], ns: Sequence[int]) -> np.ndarray:
    """len(ns) independent MSMs over sub-ranges of one resident vector: MSM j = scalars[offsets[j] : offsets[j] + ns[j]]
    against the first ns[j] bases of bases_list[j] (UnivariateKzg::batch_commit_and_write over resident quotients,
    pcs/multilinear/zeromorph.rs:150).  Returns [count, 8] affine points."""
    count = len(ns)
    assert len(offsets) == count and len(bases_list) == count
    out = np.zeros((count, 8), dtype=np.uint64)
    if count == 0:
        return out
    offs = (ctypes.c_size_t * count)(*[int(o) for o in offsets])
    nn = (ctypes.c_size_t * count)(*[int(n) for n in ns])
    hs = np.array([b.handle for b in bases_list], dtype=np.uint64)
    rc = _lib.lib().plonkish_cuda_msm_bn254_g1_many_resident(scalars.handle, ctypes.cast(offs, ctypes.c_void_p), hs.ctypes.data,
                                                             ctypes.cast(nn, ctypes.c_void_p), count, out.ctypes.data)
    _lib.check(rc, "plonkish_cuda_msm_bn254_g1_many_resident")
    return out


def zeromorph_q_hat(q: "ResidentScalars", weights) -> "ResidentScalars":
    """q_hat of Zeromorph::open (pcs/multilinear/zeromorph.rs:157-168) from the packed quotients of fr_quotients:
    q_hat[2^n - 2^i + j] += weights[i] * q_i[j]; weights: [n, 4] Montgomery Fr (the powers of y)."""
    w = _as_u64(weights, 4, "weights") if len(weights) else np.zeros((0, 4), dtype=np.uint64)
    k = w.shape[0]
    assert q.n == 1 << k, "weights / quotient buffer size mismatch"
    out = ctypes.c_uint64(0)
    _lib.check(_lib.lib().plonkish_cuda_zeromorph_q_hat_bn254(q.handle, w.ctypes.data if k else None, k, ctypes.byref(out)), "plonkish_cuda_zeromorph_q_hat_bn254")
    return ResidentScalars._adopt(out.value, q.n, q.device)


def zeromorph_f(poly: "ResidentScalars", q_hat: "ResidentScalars", q: "ResidentScalars", z, c0, q_scalars) -> "ResidentScalars":
    """f of Zeromorph::open (zeromorph.rs:175-180): z * poly + q_hat, f[0] += c0, f[j] += q_scalars[i] * q_i[j] (j < 2^i).
    z, c0: Montgomery limbs [4]; q_scalars: [n, 4]."""
    w = _as_u64(q_scalars, 4, "q_scalars") if len(q_scalars) else np.zeros((0, 4), dtype=np.uint64)
    k = w.shape[0]
    assert poly.n == 1 << k and q_hat.n == poly.n and q.n == poly.n, "polynomial / quotient sizes mismatch"
    zz = np.ascontiguousarray(z, dtype=np.uint64).reshape(4)
    cc = np.ascontiguousarray(c0, dtype=np.uint64).reshape(4)
    out = ctypes.c_uint64(0)
    rc = _lib.lib().plonkish_cuda_zeromorph_f_bn254(poly.handle, q_hat.handle, q.handle, zz.ctypes.data, cc.ctypes.data, w.ctypes.data if k else None, k,
                                                    ctypes.byref(out))
    _lib.check(rc, "plonkish_cuda_zeromorph_f_bn254")
    return ResidentScalars._adopt(out.value, poly.n, poly.device)


def eq_table(y, device: int = 0) -> "ResidentScalars":
    """eq(x, y) over the boolean hypercube as a resident polynomial (MultilinearPolynomial::eq_xy, the zero-check factor
    of piop/sum_check/classic.rs:57-61).  y: [k, 4] Montgomery Fr."""
    ys = _as_u64(y, 4, "y") if len(y) else np.zeros((0, 4), dtype=np.uint64)
    k = ys.shape[0]
    handle = ctypes.c_uint64(0)
    _lib.check(_lib.lib().plonkish_cuda_eq_table(device, ys.ctypes.data if k else None, k, ctypes.byref(handle)), "plonkish_cuda_eq_table")
    return ResidentScalars._adopt(handle.value, 1 << k, device)


def kzg_open_resident(poly: "ResidentScalars", eqs: Sequence["G1Bases"], point):
    """MultilinearKzg::open on a resident polynomial (kzg.rs:276-302): returns the num_vars
    quotient commitments ([k, 8]) and f(point) (Montgomery limbs [4])."""
    pt = _as_u64(point, 4, "point") if len(point) else np.zeros((0, 4), dtype=np.uint64)
    k = pt.shape[0]
    assert poly.n == 1 << k, "point / polynomial size mismatch"  # multilinear.rs:77
    assert len(eqs) >= k
    hs = np.array([e.handle for e in eqs[:k]], dtype=np.uint64)
    comms = np.zeros((k, 8), dtype=np.uint64)
    value = np.zeros(4, dtype=np.uint64)
    rc = _lib.lib().plonkish_cuda_kzg_open_bn254(poly.handle, hs.ctypes.data if k else None, pt.ctypes.data if k else None, k,
                                                 comms.ctypes.data if k else None, value.ctypes.data)
    _lib.check(rc, "plonkish_cuda_kzg_open_bn254")
    return comms, value


def fixed_base_msm(base, scalars, device: int = 0) -> np.ndarray:
    """fixed_base_msm + batch_normalize (msm.rs:16-31, 50-81; kzg.rs:204-207):
    [n, 8] affine points scalars[i] * base."""
    b = np.ascontiguousarray(base, dtype=np.uint64).reshape(8)
    sc = _as_u64(scalars, 4, "scalars")
    out = np.zeros((sc.shape[0], 8), dtype=np.uint64)
    rc = _lib.lib().plonkish_cuda_fixed_base_msm_bn254_g1(device, b.ctypes.data, sc.ctypes.data, sc.shape[0], out.ctypes.data)
    _lib.check(rc, "plonkish_cuda_fixed_base_msm_bn254_g1")
    return out


def kzg_setup_eqs(g1, ss, device: int = 0):
    """The prover half of MultilinearKzg::setup (kzg.rs:167-212) on the GPU: returns
    [G1Bases(eqs[0]), ..., G1Bases(eqs[num_vars])], resident, never in host memory."""
    g = np.ascontiguousarray(g1, dtype=np.uint64).reshape(8)
    s = _as_u64(ss, 4, "ss") if len(ss) else np.zeros((0, 4), dtype=np.uint64)
    k = s.shape[0]
    handles = np.zeros(k + 1, dtype=np.uint64)
    rc = _lib.lib().plonkish_cuda_kzg_setup_eqs_bn254(device, g.ctypes.data, s.ctypes.data if k else None, k, handles.ctypes.data)
    _lib.check(rc, "plonkish_cuda_kzg_setup_eqs_bn254")
    return [G1Bases._adopt(h, 1 << i, device) for i, h in enumerate(handles)]


def kzg_setup_powers(g1, s, n: int, device: int = 0) -> "G1Bases":
    """The G1 half of UnivariateKzg::setup (pcs/univariate/kzg.rs:175-195) on the GPU: powers_of_s_g1[i] = s^i * g1
    for i < n, resident."""
    g = np.ascontiguousarray(g1, dtype=np.uint64).reshape(8)
    sv = np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
    handle = ctypes.c_uint64(0)
    _lib.check(_lib.lib().plonkish_cuda_kzg_setup_powers_bn254(device, g.ctypes.data, sv.ctypes.data, n, ctypes.byref(handle)),
               "plonkish_cuda_kzg_setup_powers_bn254")
    return G1Bases._adopt(handle.value, n, device)


def variable_base_msm_many(scalars_list: Sequence, bases_list: Sequence["G1Bases"]) -> np.ndarray:
    """Independent MSMs of different sizes, MSM j against the resident slice bases_list[j] ->
    [count, 8] affine points.  The quotient commitments of MultilinearKzg::open
    (pcs/multilinear/kzg.rs:291-293) in one call: small MSMs run concurrently."""
    arrs = [_as_u64(s, 4, "scalars") for s in scalars_list]
    count = len(arrs)
    assert count == len(bases_list), "scalars and bases lists differ in length"
    out = np.zeros((count, 8), dtype=np.uint64)
    if count == 0:
        return out
    for a, b in zip(arrs, bases_list):
        assert a.shape[0] <= b.n, "more scalars than registered bases"  # msm.rs:90
    ptrs = (ctypes.c_void_p * count)(*[a.ctypes.data if a.shape[0] else None for a in arrs])
    handles = (ctypes.c_uint64 * count)(*[b.handle for b in bases_list])
    ns = (ctypes.c_size_t * count)(*[a.shape[0] for a in arrs])
    rc = _lib.lib().plonkish_cuda_msm_bn254_g1_many(ptrs, handles, ns, count, out.ctypes.data)
    _lib.check(rc, "plonkish_cuda_msm_bn254_g1_many")
    return out


def _variable_base_msm_gather(scalars: Sequence, bases: Sequence) -> np.ndarray:
    assert len(scalars) == len(bases), "scalars and bases differ in length"  # msm.rs:90
    n = len(scalars)
    keep_s = [np.ascontiguousarray(s, dtype=np.uint64).reshape(4) for s in scalars]
    keep_b = [np.ascontiguousarray(b, dtype=np.uint64).reshape(8) for b in bases]
    sp = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in keep_s])
    bp = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in keep_b])
    out = np.zeros(8, dtype=np.uint64)
    rc = _lib.lib().plonkish_cuda_msm_bn254_g1_gather(sp, bp, n, out.ctypes.data)
    _lib.check(rc, "plonkish_cuda_msm_bn254_g1_gather")
    return out


# ------------------------------------------------------------ device-resident API
def _torch():
    import torch

    return torch


def variable_base_msm_device(scalars, bases, out=None, *, window_bits: int = 0, partial: bool = False):
    """Device-resident MSM on torch CUDA tensors, enqueued on the current stream.

    scalars: [n, 4] int64/uint64 CUDA tensor (raw limbs); bases: an [n, 8] CUDA tensor
    or a G1Bases registered on the same device.  Returns a CUDA tensor holding the
    affine point ([8] limbs), or with partial=True the projective XYZZ partial
    ([16] limbs) a rank contributes before the gather.
    """
    torch = _torch()
    assert scalars.is_cuda and scalars.is_contiguous()
    n = scalars.numel() * scalars.element_size() // SCALAR_BYTES
    dev = scalars.device.index if scalars.device.index is not None else torch.cuda.current_device()
    words = 16 if partial else 8
    if out is None:
        out = torch.empty(words, dtype=torch.int64, device=scalars.device)
    stream = torch.cuda.current_stream(scalars.device).cuda_stream
    o_aff, o_xyzz = (None, out.data_ptr()) if partial else (out.data_ptr(), None)
    if isinstance(bases, G1Bases):
        assert n <= bases.n, "more scalars than registered bases"  # msm.rs:90
        assert bases.device == dev, "bases are registered on another device"
        rc = _lib.lib().plonkish_cuda_msm_bn254_g1_device_resident(scalars.data_ptr(), bases.handle, n, o_aff, o_xyzz, stream)
        _lib.check(rc, "plonkish_cuda_msm_bn254_g1_device_resident")
        return out
    assert bases.is_cuda and bases.is_contiguous()
    assert bases.numel() * bases.element_size() // AFFINE_BYTES >= n, "fewer bases than scalars"  # msm.rs:90
    rc = _lib.lib().plonkish_cuda_msm_bn254_g1_device(dev, scalars.data_ptr(), bases.data_ptr(), n, window_bits, o_aff, o_xyzz, stream)
    _lib.check(rc, "plonkish_cuda_msm_bn254_g1_device")
    return out


def host_partial(scalars, bases: "G1Bases"):
    """Host scalars against a resident slice -> this rank's projective partial, a [16]-limb int64
    CUDA tensor on the slice's device (blocking; the upload is pipelined with the compute)."""
    torch = _torch()
    sc = _as_u64(scalars, 4, "scalars")
    assert sc.shape[0] <= bases.n, "more scalars than registered bases"  # msm.rs:90
    out = torch.empty(16, dtype=torch.int64, device=torch.device("cuda", bases.device))
    rc = _lib.lib().plonkish_cuda_msm_bn254_g1_host_partial(sc.ctypes.data, bases.handle, sc.shape[0], out.data_ptr())
    _lib.check(rc, "plonkish_cuda_msm_bn254_g1_host_partial")
    return out


def sum_partials_device(partials, out=None):
    """Adds [k, 16]-limb projective partials and normalises to an affine [8] tensor."""
    torch = _torch()
    assert partials.is_cuda and partials.is_contiguous()
    count = partials.numel() * partials.element_size() // XYZZ_BYTES
    dev = partials.device.index if partials.device.index is not None else torch.cuda.current_device()
    if out is None:
        out = torch.empty(8, dtype=torch.int64, device=partials.device)
    stream = torch.cuda.current_stream(partials.device).cuda_stream
    rc = _lib.lib().plonkish_cuda_g1_sum_partials_device(dev, partials.data_ptr(), count, out.data_ptr(), stream)
    _lib.check(rc, "plonkish_cuda_g1_sum_partials_device")
    return out


def synth_bases_device(n: int, a: int, step: int, device=None, first: int = 0):
    """bases[i] = (a + (first + i)*step) * G on the GPU -> [n, 8] int64 CUDA tensor."""
    torch = _torch()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((n, 8), dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    rc = _lib.lib().plonkish_cuda_synth_bases_device(device.index or 0, out.data_ptr(), first, n, a, step, stream)
    _lib.check(rc, "plonkish_cuda_synth_bases_device")
    return out


def msm_plan(n: int, window_bits: int = 0, device: int = 0, bases: Optional["G1Bases"] = None) -> dict:
    out = (ctypes.c_uint32 * 8)()
    handle = bases.handle if bases is not None else 0
    _lib.check(_lib.lib().plonkish_cuda_msm_plan(device, n, window_bits, handle, out), "plonkish_cuda_msm_plan")
    keys = ("window_bits", "windows", "hi_bits", "lo_bits", "idx_bits", "tile", "run_length", "accumulate_threads")
    return dict(zip(keys, [int(v) for v in out]))


STAGES = ("decompose", "scans", "bin_scatter", "bin_sort", "accumulate", "item_levels", "bucket_reduce", "window_combine", "finalize")


def profile_stages_device(scalars, bases, *, window_bits: int = 0) -> dict:
    """One synchronous MSM with CUDA events between the stages -> {stage: ms}."""
    torch = _torch()
    n = scalars.numel() * scalars.element_size() // SCALAR_BYTES
    dev = scalars.device.index if scalars.device.index is not None else torch.cuda.current_device()
    torch.cuda.synchronize(scalars.device)
    out = (ctypes.c_double * 9)()
    if isinstance(bases, G1Bases):
        rc = _lib.lib().plonkish_cuda_msm_profile_device(dev, scalars.data_ptr(), None, bases.handle, n, 0, None, out)
    else:
        rc = _lib.lib().plonkish_cuda_msm_profile_device(dev, scalars.data_ptr(), bases.data_ptr(), 0, n, window_bits, None, out)
    _lib.check(rc, "plonkish_cuda_msm_profile_device")
    return dict(zip(STAGES, [float(v) for v in out]))


def launch_count() -> int:
    return int(_lib.load().plonkish_cuda_launch_count())


def bench_integer_pipe(device: int = 0) -> dict:
    out = (ctypes.c_double * 6)()
    _lib.check(_lib.lib().plonkish_cuda_bench_integer_pipe(device, out), "plonkish_cuda_bench_integer_pipe")
    return {"imad_wide_per_s": out[0], "fq_mul_per_s": out[1], "sm_max_mhz": out[2], "sm_count": int(out[3]),
            "imad32_per_s": out[4], "imad_wide_chain_per_s": out[5]}


def bench_fp64_pipe(device: int = 0) -> dict:
    out = (ctypes.c_double * 3)()
    _lib.check(_lib.lib().plonkish_cuda_bench_fp64_pipe(device, out), "plonkish_cuda_bench_fp64_pipe")
    return {"dfma_per_s": out[0], "dfma_per_s_mixed": out[1], "imad_wide_per_s_mixed": out[2]}


def bench_issue_mix(device: int = 0) -> dict:
    out = (ctypes.c_double * 4)()
    _lib.check(_lib.lib().plonkish_cuda_bench_issue_mix(device, out), "plonkish_cuda_bench_issue_mix")
    return {f"imad_wide_per_s_with_{r}_adds": out[r] for r in range(4)}


def bench_dp_madd(dp_blocks_per_sm: int = 1, int_blocks_per_sm: int = 2, device: int = 0) -> dict:
    out = (ctypes.c_double * 5)()
    _lib.check(_lib.lib().plonkish_cuda_bench_dp_madd(device, dp_blocks_per_sm, int_blocks_per_sm, out), "plonkish_cuda_bench_dp_madd")
    return {"dp_madd_per_s_alone": out[0], "int_madd_per_s_alone": out[1], "dp_madd_per_s_together": out[2], "int_madd_per_s_together": out[3],
            "together_ms": out[4]}


def bench_madd(device: int = 0) -> dict:
    """Field products per second of register-resident mixed-addition streams (see the header)."""
    out = (ctypes.c_double * 5)()
    _lib.check(_lib.lib().plonkish_cuda_bench_madd(device, out), "plonkish_cuda_bench_madd")
    keys = ("madd_1acc_128regs", "madd_2acc_128regs", "fq_mul_1chain", "madd_1acc_uncapped", "madd_2acc_uncapped")
    return dict(zip(keys, [float(v) for v in out]))


def random_scalars(n: int, seed: int) -> np.ndarray:
    """Synthetic Fr elements (Montgomery limbs): three uniform u64 limbs and a top
    limb drawn below r's top limb, so every value is a valid representation < r."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    out[:, 3] = rng.integers(0, 0x30644E72E131A029, size=n, dtype=np.uint64)
    return out
