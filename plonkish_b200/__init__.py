"""plonkish_b200 — BN254 G1 variable-base MSM for amit0365/plonkish on NVIDIA B200.

Only the hot path of SURVEY.md §8 lives here: hand-written sm_100a kernels and the
C ABI (csrc/, include/plonkish_cuda.h) plus the host-side mirror of the reference
interface (msm.py, kzg.py, distributed.py).
"""
from ._lib import PlonkishCudaError, LIB_PATH  # noqa: F401
from .msm import (  # noqa: F401
    G1Bases,
    ResidentScalars,
    ShardedG1Bases,
    variable_base_msm_batch_keep,
    fr_linear_combination,
    fr_div_linear,
    fr_quotients,
    variable_base_msm_many_resident,
    zeromorph_q_hat,
    zeromorph_f,
    fr_affine_table,
    fr_evaluate,
    fr_expression_table,
    lookup_m_poly,
    lookup_h_poly,
    permutation_z_polys,
    kzg_open_resident,
    eq_table,
    fixed_base_msm,
    kzg_setup_eqs,
    kzg_setup_powers,
    variable_base_msm,
    variable_base_msm_batch,
    variable_base_msm_many,
    variable_base_msm_device,
    sum_partials_device,
    host_partial,
    synth_bases_device,
    msm_plan,
    profile_stages_device,
    launch_count,
    bench_integer_pipe,
    bench_madd,
    bench_fp64_pipe,
    bench_dp_madd,
    bench_issue_mix,
    cached_bases,
    cache_evict,
    cache_limit,
    cache_stats,
    timer_config,
    staged_bytes,
    staging_rate_gbps,
    random_scalars,
)
from .distributed import shard_bounds, variable_base_msm_sharded, variable_base_msm_sharded_host  # noqa: F401
