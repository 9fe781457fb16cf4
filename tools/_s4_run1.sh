set -x
python tools/acc_run_length.py > gpurun_out/s4_acc.log 2>&1
for p in 0 1; do PLONKISH_CUDA_LANE_PRIORITY=$p python tools/open_timing.py 24 default >> gpurun_out/s4_open.log 2>&1; PLONKISH_CUDA_LANE_PRIORITY=$p python tools/open_timing.py 20 default >> gpurun_out/s4_open.log 2>&1; done
for nt in 0 1; do PLONKISH_CUDA_STAGE_NT=$nt python tools/e2e_pageable.py 24 >> gpurun_out/s4_e2e.log 2>&1; done
tail -n 80 gpurun_out/s4_acc.log; cat gpurun_out/s4_open.log gpurun_out/s4_e2e.log
