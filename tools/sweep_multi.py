"""Strong-scaling sweep (BASELINE.json config 5): one MSM of 2^k points, k = 16..26, point-sharded over the ranks of a
torchrun launch (ceil(n / G) points per rank, msm.rs:101), device-resident scalars and bases, NCCL gather of the
partials, every result checked against the known-discrete-log answer.  Time = max over ranks of the CUDA-event time.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/sweep_multi.py [lo hi]
Rank 0 prints one JSON line per size and writes gpurun_out/sweep_multi_G.json."""
import json, os, sys
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonkish_b200 as pk
from plonkish_b200.distributed import shard_bounds, variable_base_msm_sharded
from oracle import pyoracle as po

lo = int(sys.argv[1]) if len(sys.argv) > 1 else 16
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 26
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
else:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29577", rank=0, world_size=1)
rows = []
for lg in range(lo, hi + 1):
    n = 1 << lg
    beg, end = shard_bounds(n, world, rank)
    cnt = end - beg
    sc_all = pk.random_scalars(n, seed=lg) if rank == 0 else None   # rank 0 needs every scalar for the known answer
    sc = pk.random_scalars(n, seed=lg)[beg:end]
    d_sc = torch.from_numpy(np.ascontiguousarray(sc).view(np.int64)).to(dev)
    d_bs = pk.synth_bases_device(cnt, 3, 5, device=dev, first=beg)
    torch.cuda.synchronize()
    reg = pk.G1Bases(d_bs)
    out = variable_base_msm_sharded(d_sc, reg)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        ok = out.cpu().numpy().view(np.uint64).tobytes() == po.known_dlog_answer(3, 5, sc_all).tobytes()
    ts = []
    for _ in range(5 if lg <= 22 else 3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); variable_base_msm_sharded(d_sc, reg); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    if rank == 0:
        row = {"log_n": lg, "gpus": world, "points_per_gpu": cnt, "ms": round(min(ts), 3), "mpoints_per_s": round(n / min(ts) / 1e3, 1), "parity": bool(ok)}
        print(json.dumps(row), flush=True)
        rows.append(row)
    reg.release()
    del d_sc, d_bs
    torch.cuda.empty_cache()
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"sweep_multi_{world}.json"), "w"), indent=1)
dist.destroy_process_group()
