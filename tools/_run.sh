python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python tools/prove_bench.py 20 3 0 > gpurun_out/s5_prove20.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/s5_prove20.json')); print(20, d['gpu_ms'], d['phases_ms'][0])"
python tools/prove_bench.py 24 3 0 > gpurun_out/s5_prove24.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/s5_prove24.json')); print(24, d['gpu_ms'], d['phases_ms'][0])"
