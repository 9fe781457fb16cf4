"""ctypes binding of libplonkish_cuda.so (include/plonkish_cuda.h).

The library is built in-tree by ``__graft_entry__.build()`` /
``plonkish_b200.build.build_library()``.  There is no CPU fallback: if the shared
object is missing, or no CUDA device is usable, every call raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PLONKISH_CUDA_LIB overrides the path (A/B builds of the same ABI).
LIB_PATH = os.environ.get("PLONKISH_CUDA_LIB") or os.path.join(_HERE, "libplonkish_cuda.so")

# Every symbol include/plonkish_cuda.h declares (checked by tests/test_abi.py).
EXPORTS = (
    "plonkish_cuda_init",
    "plonkish_cuda_device_count",
    "plonkish_cuda_shutdown",
    "plonkish_cuda_last_error",
    "plonkish_cuda_bases_register",
    "plonkish_cuda_bases_release",
    "plonkish_cuda_bases_register_device",
    "plonkish_cuda_bases_cached",
    "plonkish_cuda_bases_cached_sharded",
    "plonkish_cuda_bases_cache_evict",
    "plonkish_cuda_bases_cache_limit",
    "plonkish_cuda_bases_cache_stats",
    "plonkish_cuda_timer_config",
    "plonkish_cuda_timer_emit",
    "plonkish_cuda_bases_register_sharded_device",
    "plonkish_cuda_staged_bytes",
    "plonkish_cuda_staging_rate_gbps",
    "plonkish_cuda_bench_fp64_pipe",
    "plonkish_cuda_bench_dp_madd",
    "plonkish_cuda_bench_issue_mix",
    "plonkish_cuda_msm_bn254_g1",
    "plonkish_cuda_msm_bn254_g1_batch",
    "plonkish_cuda_msm_bn254_g1_many",
    "plonkish_cuda_scalars_register",
    "plonkish_cuda_scalars_release",
    "plonkish_cuda_scalars_read",
    "plonkish_cuda_bases_read",
    "plonkish_cuda_msm_bn254_g1_resident",
    "plonkish_cuda_msm_bn254_g1_batch_keep",
    "plonkish_cuda_fr_linear_combination",
    "plonkish_cuda_fr_div_linear",
    "plonkish_cuda_fr_quotients",
    "plonkish_cuda_scalars_slice",
    "plonkish_cuda_fr_linear_combination_padded",
    "plonkish_cuda_fr_gemini_folds",
    "plonkish_cuda_msm_bn254_g1_many_resident",
    "plonkish_cuda_zeromorph_q_hat_bn254",
    "plonkish_cuda_zeromorph_f_bn254",
    "plonkish_cuda_permutation_z_polys_bn254",
    "plonkish_cuda_fr_affine_table",
    "plonkish_cuda_fr_evaluate",
    "plonkish_cuda_fr_expression_table",
    "plonkish_cuda_lookup_m_poly_bn254",
    "plonkish_cuda_lookup_h_poly_bn254",
    "plonkish_cuda_kzg_open_bn254",
    "plonkish_cuda_fixed_base_msm_bn254_g1",
    "plonkish_cuda_kzg_setup_eqs_bn254",
    "plonkish_cuda_kzg_setup_powers_bn254",
    "plonkish_cuda_eq_table",
    "plonkish_cuda_keccak_f1600",
    "plonkish_cuda_sumcheck_new",
    "plonkish_cuda_sumcheck_degree",
    "plonkish_cuda_sumcheck_round",
    "plonkish_cuda_sumcheck_round_factored",
    "plonkish_cuda_sumcheck_fix_var",
    "plonkish_cuda_sumcheck_final_evals",
    "plonkish_cuda_sumcheck_free",
    "plonkish_cuda_msm_bn254_g1_gather",
    "plonkish_cuda_bases_register_sharded",
    "plonkish_cuda_msm_bn254_g1_multi",
    "plonkish_cuda_msm_bn254_g1_device",
    "plonkish_cuda_msm_bn254_g1_device_resident",
    "plonkish_cuda_msm_bn254_g1_host_partial",
    "plonkish_cuda_g1_sum_partials_device",
    "plonkish_cuda_msm_plan",
    "plonkish_cuda_msm_profile_device",
    "plonkish_cuda_launch_count",
    "plonkish_cuda_bench_integer_pipe",
    "plonkish_cuda_bench_fq_mul_occupancy",
    "plonkish_cuda_bench_inversion",
    "plonkish_cuda_bench_madd",
    "plonkish_cuda_bench_row_forms",
    "plonkish_cuda_synth_bases_device",
    "plonkish_cuda_debug_field_op",
    "plonkish_cuda_debug_point_op",
)


class PlonkishCudaError(RuntimeError):
    """A non-zero return code from the C ABI (the Rust shim panics on these)."""


_lib = None
_initialised = False


def load() -> ctypes.CDLL:
    """dlopen the library and declare the prototypes; does not touch the GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlonkishCudaError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). plonkish_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, u64, u32, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    lib.plonkish_cuda_init.argtypes = [ci]
    lib.plonkish_cuda_device_count.argtypes = []
    lib.plonkish_cuda_shutdown.argtypes = []
    lib.plonkish_cuda_shutdown.restype = None
    lib.plonkish_cuda_last_error.argtypes = []
    lib.plonkish_cuda_last_error.restype = ctypes.c_char_p
    lib.plonkish_cuda_bases_register.argtypes = [ci, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_bases_release.argtypes = [u64]
    lib.plonkish_cuda_bases_register_device.argtypes = [ci, vp, sz, ci, ctypes.POINTER(u64)]
    lib.plonkish_cuda_bases_cached.argtypes = [ci, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_bases_cached_sharded.argtypes = [ci, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_bases_cache_evict.argtypes = [vp]
    lib.plonkish_cuda_bases_cache_limit.argtypes = [sz]
    lib.plonkish_cuda_bases_cache_stats.argtypes = [ctypes.POINTER(sz)]
    lib.plonkish_cuda_timer_config.argtypes = [ci, ci]
    lib.plonkish_cuda_timer_emit.argtypes = [sz, ctypes.c_double]
    lib.plonkish_cuda_timer_emit.restype = None
    lib.plonkish_cuda_bases_register_sharded_device.argtypes = [ci, vp, sz, ci, ctypes.POINTER(u64)]
    lib.plonkish_cuda_staged_bytes.argtypes = []
    lib.plonkish_cuda_staged_bytes.restype = u64
    lib.plonkish_cuda_staging_rate_gbps.argtypes = []
    lib.plonkish_cuda_staging_rate_gbps.restype = ctypes.c_double
    lib.plonkish_cuda_bench_fp64_pipe.argtypes = [ci, ctypes.POINTER(ctypes.c_double)]
    lib.plonkish_cuda_bench_issue_mix.argtypes = [ci, ctypes.POINTER(ctypes.c_double)]
    lib.plonkish_cuda_bench_dp_madd.argtypes = [ci, ci, ci, ctypes.POINTER(ctypes.c_double)]
    lib.plonkish_cuda_msm_bn254_g1.argtypes = [vp, vp, u64, sz, vp]
    lib.plonkish_cuda_msm_bn254_g1_batch.argtypes = [vp, sz, u64, sz, vp]
    lib.plonkish_cuda_msm_bn254_g1_many.argtypes = [vp, vp, vp, sz, vp]
    lib.plonkish_cuda_msm_bn254_g1_gather.argtypes = [vp, vp, sz, vp]
    lib.plonkish_cuda_kzg_setup_powers_bn254.argtypes = [ci, vp, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_keccak_f1600.argtypes = [vp]
    lib.plonkish_cuda_keccak_f1600.restype = None
    lib.plonkish_cuda_eq_table.argtypes = [ci, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_sumcheck_new.argtypes = [vp, sz, sz, vp, vp, vp, sz, ci, ctypes.POINTER(u64)]
    lib.plonkish_cuda_sumcheck_degree.argtypes = [u64]
    lib.plonkish_cuda_sumcheck_round.argtypes = [u64, vp]
    lib.plonkish_cuda_sumcheck_round_factored.argtypes = [u64, vp]
    lib.plonkish_cuda_sumcheck_fix_var.argtypes = [u64, vp]
    lib.plonkish_cuda_sumcheck_final_evals.argtypes = [u64, vp]
    lib.plonkish_cuda_sumcheck_free.argtypes = [u64]
    lib.plonkish_cuda_scalars_register.argtypes = [ci, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_scalars_release.argtypes = [u64]
    lib.plonkish_cuda_scalars_read.argtypes = [u64, sz, sz, vp]
    lib.plonkish_cuda_bases_read.argtypes = [u64, sz, sz, vp]
    lib.plonkish_cuda_msm_bn254_g1_resident.argtypes = [u64, u64, sz, vp]
    lib.plonkish_cuda_msm_bn254_g1_batch_keep.argtypes = [vp, sz, u64, sz, vp, vp]
    lib.plonkish_cuda_fr_linear_combination.argtypes = [vp, vp, sz, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_permutation_z_polys_bn254.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
    lib.plonkish_cuda_fr_div_linear.argtypes = [u64, vp, ctypes.POINTER(u64), vp]
    lib.plonkish_cuda_fr_quotients.argtypes = [u64, vp, sz, ctypes.POINTER(u64), vp]
    lib.plonkish_cuda_scalars_slice.argtypes = [u64, sz, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_fr_linear_combination_padded.argtypes = [vp, vp, sz, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_fr_gemini_folds.argtypes = [u64, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_msm_bn254_g1_many_resident.argtypes = [u64, vp, vp, vp, sz, vp]
    lib.plonkish_cuda_zeromorph_q_hat_bn254.argtypes = [u64, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_zeromorph_f_bn254.argtypes = [u64, u64, u64, vp, vp, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_fr_affine_table.argtypes = [ci, sz, vp, vp, vp, sz, vp, vp, vp, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_fr_evaluate.argtypes = [u64, vp, sz, sz, vp]
    lib.plonkish_cuda_fr_expression_table.argtypes = [vp, sz, sz, vp, vp, vp, sz, ci, ctypes.POINTER(u64)]
    lib.plonkish_cuda_lookup_m_poly_bn254.argtypes = [u64, u64, ctypes.POINTER(u64)]
    lib.plonkish_cuda_lookup_h_poly_bn254.argtypes = [u64, u64, u64, vp, ctypes.POINTER(u64)]
    lib.plonkish_cuda_kzg_open_bn254.argtypes = [u64, vp, vp, sz, vp, vp]
    lib.plonkish_cuda_fixed_base_msm_bn254_g1.argtypes = [ci, vp, vp, sz, vp]
    lib.plonkish_cuda_kzg_setup_eqs_bn254.argtypes = [ci, vp, vp, sz, vp]
    lib.plonkish_cuda_bases_register_sharded.argtypes = [ci, vp, sz, ctypes.POINTER(u64)]
    lib.plonkish_cuda_msm_bn254_g1_multi.argtypes = [ci, vp, vp, u64, sz, vp]
    lib.plonkish_cuda_msm_bn254_g1_device.argtypes = [ci, vp, vp, sz, u32, vp, vp, vp]
    lib.plonkish_cuda_msm_bn254_g1_device_resident.argtypes = [vp, u64, sz, vp, vp, vp]
    lib.plonkish_cuda_msm_bn254_g1_host_partial.argtypes = [vp, u64, sz, vp]
    lib.plonkish_cuda_g1_sum_partials_device.argtypes = [ci, vp, sz, vp, vp]
    lib.plonkish_cuda_msm_plan.argtypes = [ci, sz, u32, u64, ctypes.POINTER(u32)]
    lib.plonkish_cuda_msm_profile_device.argtypes = [ci, vp, vp, u64, sz, u32, vp, ctypes.POINTER(ctypes.c_double)]
    lib.plonkish_cuda_launch_count.argtypes = []
    lib.plonkish_cuda_launch_count.restype = u64
    lib.plonkish_cuda_bench_integer_pipe.argtypes = [ci, ctypes.POINTER(ctypes.c_double)]
    lib.plonkish_cuda_bench_fq_mul_occupancy.argtypes = [ci, ci, ctypes.POINTER(ctypes.c_double)]
    lib.plonkish_cuda_bench_inversion.argtypes = [ci, ctypes.POINTER(ctypes.c_double)]
    lib.plonkish_cuda_bench_madd.argtypes = [ci, ctypes.POINTER(ctypes.c_double)]
    lib.plonkish_cuda_bench_row_forms.argtypes = [ci, ctypes.POINTER(ctypes.c_double)]
    lib.plonkish_cuda_synth_bases_device.argtypes = [ci, vp, sz, sz, u64, u64, vp]
    lib.plonkish_cuda_debug_field_op.argtypes = [ci, ci, vp, vp, vp, sz]
    lib.plonkish_cuda_debug_point_op.argtypes = [ci, ci, vp, vp, vp, sz]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().plonkish_cuda_last_error().decode("utf-8", "replace")
        raise PlonkishCudaError(f"{what} failed (code {rc}): {msg}")


def lib() -> ctypes.CDLL:
    """The loaded library with its device contexts created (plonkish_cuda_init)."""
    global _initialised
    l = load()
    if not _initialised:
        check(l.plonkish_cuda_init(0), "plonkish_cuda_init")
        _initialised = True
    return l
