N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/s4_bench_${N}gpu.json 2> gpurun_out/s4_bench_${N}gpu.err
tail -c 300 gpurun_out/s4_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s4_bench_${N}gpu.json').read().strip().splitlines()[-1])
s=d['strong_2p24']
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('staging_rate_gbps_rank0'), 'pinned', d['e2e_pinned']['ms_per_step'], 'strong', s['ms_per_step'], s['e2e_ms_per_step'], s['efficiency_vs_one_gpu_same_run'], 'sp', d['single_process'].get('e2e_pageable_ms'), d['single_process'].get('e2e_pinned_ms'), d['single_process'].get('error'))
PY
nproc
