python -m pytest tests -x -q -m gpu -k "multi or shard or pageable or staging" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s4_bench_2gpu_b.json 2> gpurun_out/s4_bench_2gpu_b.err
tail -c 300 gpurun_out/s4_bench_2gpu_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s4_bench_2gpu_b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('staging_rate_gbps_rank0'), 'pinned', d['e2e_pinned']['ms_per_step'], 'strong', d['strong_2p24']['ms_per_step'], d['strong_2p24']['e2e_ms_per_step'], d['strong_2p24']['efficiency_vs_one_gpu_same_run'], 'sp', d['single_process'])
PY
