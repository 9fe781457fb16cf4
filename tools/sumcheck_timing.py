"""Vanilla-plonk-shaped zero check (9 tables, degree 4) at 2^k rows: time per round and in total."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import plonkish_b200 as pk
from plonkish_b200 import sumcheck

k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << k
tables = [pk.ResidentScalars(pk.random_scalars(n, seed=100 + i)) for i in range(9)]
one = sumcheck._to_mont(1)
terms = [(one, [1, 6]), (one, [2, 7]), (one, [3, 6, 7]), (one, [4, 8]), (one, [5])]
for rep in range(2):
    prover = sumcheck.SumCheckProver(tables, terms, common=0)
    torch.cuda.synchronize()
    t_round, t_fold = [], []
    t0 = time.perf_counter()
    for rnd in range(k):
        a = time.perf_counter(); prover.round_evals(); b = time.perf_counter()
        prover.fix_var(one); c = time.perf_counter()
        t_round.append((b - a) * 1e3); t_fold.append((c - b) * 1e3)
    total = (time.perf_counter() - t0) * 1e3
    prover.free()
    print(f"k={k} rep {rep}: total {total:.2f} ms; rounds {sum(t_round):.2f} ms (first {t_round[0]:.2f}, second {t_round[1]:.2f}); "
          f"folds {sum(t_fold):.2f} ms (first {t_fold[0]:.2f}); last 10 rounds+folds {sum(t_round[-10:]) + sum(t_fold[-10:]):.2f} ms", flush=True)
