ACC_OUT=acc_run_length_d.json python tools/acc_run_length.py 14,16,18,19,20,21,22,23,24 default,T1,T3,T4,T5,T6,T8,T10,L64,L128,L256 > gpurun_out/s4_acc_d.log 2>&1
cat gpurun_out/s4_acc_d.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
