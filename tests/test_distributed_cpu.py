"""CPU suite, world_size 2 over gloo: the point sharding and the gather of the
per-rank partials (plonkish_b200/distributed.py).  The partial itself comes from
the CUDA kernels on a GPU box; here each rank produces it with the oracle so that
the N>1 plumbing — shard bounds, all_gather layout, fold — is what is tested."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as po
        from plonkish_b200.distributed import gather_partials, shard_bounds

        sc = po.random_scalars(n, 42)
        bs = po.known_dlog_bases(3, 5, n, 2)
        b, e = shard_bounds(n, world, rank)
        local = po.variable_base_msm(sc[b:e], bs[b:e], 1) if e > b else np.zeros(8, np.uint64)
        # Stand-in for the device partial: affine point padded to the 16-limb XYZZ slot.
        partial = torch.zeros(16, dtype=torch.int64)
        partial[:8] = torch.from_numpy(local.view(np.int64))
        allp = gather_partials(partial).numpy().view(np.uint64)
        assert allp.shape == (world, 16)
        assert (allp[rank, :8] == local).all()
        # fold (msm.rs:112-114) with the oracle: sum of 1 * partial_r
        ones = po.from_canonical(1, np.tile(np.array([1, 0, 0, 0], np.uint64), (world, 1)))
        total = po.variable_base_msm(ones, allp[:, :8].copy(), 1)
        want = po.known_dlog_answer(3, 5, sc)
        ret[rank] = bool((total == want).all())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 1001])
def test_two_rank_shard_and_gather(n):
    world = 2
    port = 29500 + (os.getpid() % 2000) + n % 7
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
