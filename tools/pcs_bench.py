"""The pcs_schemes leg of bench.py on its own — Zeromorph and Gemini over the univariate KZG SRS (pcs/multilinear/
zeromorph.rs, gemini.rs): commit + open of one 2^k-evaluation polynomial and HyperPlonk::prove over the scheme, every
proof checked against the scheme's verifier equation:
python tools/pcs_bench.py [k] [reps] [prove: 0|1] [zeromorph|gemini|both] [prefix tables: 0|1]."""
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
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with_prove = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
    which = sys.argv[4] if len(sys.argv) > 4 else "both"
    prefix_tables = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
    torch.cuda.init()
    print(json.dumps(bench.pcs_schemes_bench(pk, torch, np, k, reps, with_prove, which, prefix_tables), indent=1))
