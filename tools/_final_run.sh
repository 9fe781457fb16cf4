N=$1
if [ "$N" = "1" ]; then
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py > gpurun_out/r02_bench_line_k.json 2> gpurun_out/r02_bench_line_k.err ) 2> gpurun_out/r02_bench_line_k.time
tail -3 gpurun_out/r02_bench_line_k.time; tail -c 300 gpurun_out/r02_bench_line_k.err
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_line_${N}gpu_k.json 2> gpurun_out/r02_bench_line_${N}gpu_k.err
tail -c 300 gpurun_out/r02_bench_line_${N}gpu_k.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_line_${N}gpu_k.json').read().strip().splitlines()[-1])
s=d['strong_2p24']
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('staging_rate_gbps_rank0'), 'pinned', d['e2e_pinned']['ms_per_step'], 'strong', s['ms_per_step'], s['e2e_ms_per_step'], s['efficiency_vs_one_gpu_same_run'], 'sp', d['single_process'].get('e2e_pageable_ms'), d['single_process'].get('e2e_pinned_ms'), d['single_process'].get('error'))
PY
fi
