"""One process per GPU: point-sharded MSM with an NCCL gather of the partials.

The reference shards the MSM across CPU threads — contiguous chunks of
ceil(n/T) points, an independent Pippenger per chunk, then a fold of the T
projective partials (/root/reference/plonkish_backend/src/util/arithmetic/msm.rs:101-114,
util/parallel.rs:9-25).  Here the threads are ranks: every rank runs the whole
pipeline on its slice, the only exchange is one 128-byte projective point per
rank (torch.distributed all_gather -> NCCL over NVLink on GPUs, gloo in the CPU
tests), and the fold + to_affine runs on every rank's GPU (so all ranks hold the
commitment, as every rayon thread's caller would).
"""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of rank's points: msm.rs:101 `chunk_size = div_ceil(n, T)`,
    chunks(chunk_size) — trailing ranks may get a short or empty slice."""
    chunk = -(-n // world_size) if n else 0
    begin = min(rank * chunk, n)
    end = min(begin + chunk, n)
    return begin, end


def gather_partials(partial, group=None):
    """all_gather of one [16]-limb int64 partial per rank -> [world, 16] tensor."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    flat = torch.empty(world * partial.numel(), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(flat, partial.contiguous().view(-1), group=group)
    return flat.view(world, partial.numel())


def variable_base_msm_sharded(local_scalars, local_bases, group=None, *, window_bits: int = 0):
    """Each rank passes its own slice (CUDA scalars; bases as a CUDA tensor or a
    G1Bases registered on this rank's GPU); returns the affine sum of all ranks'
    slices as an [8]-limb CUDA tensor on every rank."""
    from . import msm

    partial = msm.variable_base_msm_device(local_scalars, local_bases, window_bits=window_bits, partial=True)
    allp = gather_partials(partial, group)
    return msm.sum_partials_device(allp)


def variable_base_msm_sharded_host(local_scalars, local_bases, group=None):
    """Same with this rank's scalars in host memory (numpy [n, 4] uint64) and its bases resident
    (G1Bases): the C-ABI host call uploads the scalars in chunks overlapped with the compute and
    leaves the partial on the GPU; all_gather + fold follow.  Returns the affine sum ([8] CUDA tensor)."""
    from . import msm

    partial = msm.host_partial(local_scalars, local_bases)
    allp = gather_partials(partial, group)
    return msm.sum_partials_device(allp)
