ACC_OUT=acc_run_length_b.json python tools/acc_run_length.py 16,17,18,19,20,21,22 default,8,10,12,16,24,32 > gpurun_out/s4_acc_b.log 2>&1
ACC_OUT=acc_run_length_c.json python tools/acc_run_length.py 22,23,24 default,L64,L96,L128,L160,L192,L224 > gpurun_out/s4_acc_c.log 2>&1
cat gpurun_out/s4_acc_b.log gpurun_out/s4_acc_c.log
