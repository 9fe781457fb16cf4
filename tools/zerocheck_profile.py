"""Where the zero check of the full vanilla_plonk proof spends its time (plonkish_b200/hyperplonk.py prove_sum_check):
tables of the compiled expression, every round (round polynomial / fold), the evaluations at the rotated points.
Random tables (the arithmetic does not depend on the circuit being satisfied).  python tools/zerocheck_profile.py [k]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import plonkish_b200 as pk  # noqa: E402
from plonkish_b200 import hyperplonk as hp, sumcheck  # noqa: E402
from plonkish_b200.expression import compile_expression  # noqa: E402
from plonkish_b200.transcript import fr_to_montgomery  # noqa: E402

if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << k
    info = hp.vanilla_plonk_circuit_info(k, k, [np.zeros((1, 4), dtype=np.uint64)] * 5, [[(6, 1)], [(7, 1)], [(8, 1)]])
    _, expression = hp.compose(info)
    polys = [pk.ResidentScalars(pk.random_scalars(n, seed=900 + i)) for i in range(13)]
    rng = np.random.default_rng(1)
    fe = lambda: int.from_bytes(rng.bytes(40), "little") % hp.R  # noqa: E731
    challenges, y = [fe() for _ in range(3)], [fe() for _ in range(k)]
    out = {"k": k}
    for rep in range(2):
        t0 = time.perf_counter()
        compiled = compile_expression(expression, challenges)
        t1 = time.perf_counter()
        queries = hp.pcs_query(expression, 1)
        st = hp.build_tables(compiled, k, polys, [y], extra_polys=sorted({q.poly for q in queries}))
        t2 = time.perf_counter()
        terms = [(fr_to_montgomery(c), idx) for c, idx in compiled.terms]
        prover = sumcheck.SumCheckProver(st.tables, terms, compiled.common)
        t3 = time.perf_counter()
        rounds, folds, x = [], [], []
        for _ in range(k):
            a = time.perf_counter()
            prover.round_evals()
            b = time.perf_counter()
            ch = fe()
            x.append(ch)
            prover.fix_var(fr_to_montgomery(ch))
            c = time.perf_counter()
            rounds.append(round((b - a) * 1e3, 3)); folds.append(round((c - b) * 1e3, 3))
        prover.final_evals()
        prover.free()
        st.release()
        t4 = time.perf_counter()
        pts = hp.rotation_eval_points(x, 1)
        pk.fr_evaluate(polys[12], np.stack([np.stack([fr_to_montgomery(v) for v in pt]) for pt in pts]))
        t5 = time.perf_counter()
        out = {"k": k, "tables": len(st.tables), "terms": len(terms), "degree": compiled.degree, "compile_ms": (t1 - t0) * 1e3, "build_tables_ms": (t2 - t1) * 1e3,
               "state_new_ms": (t3 - t2) * 1e3, "round_ms": rounds, "fold_ms": folds, "rounds_total_ms": sum(rounds), "folds_total_ms": sum(folds),
               "rotated_evals_ms": (t5 - t4) * 1e3, "total_ms": (t5 - t0) * 1e3}
    print(json.dumps(out, indent=1))
