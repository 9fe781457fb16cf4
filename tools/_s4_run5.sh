export PLONKISH_CUDA_STAGE_NT=1
i=0
for cuts in "" "0.0625,0.16,0.32,0.58,1" "0.04,0.12,0.28,0.55,1" "0.0625,0.1875,0.45,1" "0.03,0.08,0.17,0.32,0.57,1"; do
i=$((i+1))
if [ -n "$cuts" ]; then export PLONKISH_CUDA_HOST_CUTS=$cuts; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 --no-single-process > gpurun_out/s4_cuts_$i.json 2> gpurun_out/s4_cuts_$i.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s4_cuts_$i.json').read().strip().splitlines()[-1])
print("cuts '$cuts':", 'dev', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'pinned', round(d['e2e_pinned']['ms_per_step'],2), 'strong e2e', round(d['strong_2p24']['e2e_ms_per_step'],2))
PY
done
