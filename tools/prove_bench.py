"""The HyperPlonk prove leg of bench.py on its own (k and repetitions from the command line), with the phase split:
python tools/prove_bench.py [k] [reps] [cpu: 0|1]."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import plonkish_b200 as pk  # noqa: E402

if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cpu = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
    torch.cuda.init()
    print(json.dumps(bench.hyperplonk_prove_bench(pk, torch, np, k, cpu=cpu, reps=reps), indent=1))
