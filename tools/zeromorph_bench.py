"""Zeromorph<UnivariateKzg> on one GPU: commit + open of one 2^k-evaluation polynomial with the phase split, checked
against the verifier's equation (tests/zeromorph_ref.py, G1 with the setup's trapdoor), and HyperPlonk::prove for
vanilla_plonk over this PCS on bench.py's synthetic circuit, accepted by the verifier restatement:
python tools/zeromorph_bench.py [k] [reps] [prove: 0|1]."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import plonkish_b200 as pk  # noqa: E402


def main():
    import hyperplonk_ref as ref
    import zeromorph_ref as zr
    from oracle import bigint_ref as br
    from oracle import pyoracle as po
    from plonkish_b200 import hyperplonk, kzg, zeromorph
    from plonkish_b200.sumcheck import _to_int, _to_mont
    from plonkish_b200.transcript import Keccak256Transcript
    from univariate_verify import as_limbs

    k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with_prove = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
    torch.cuda.init()
    n, s = 1 << k, 0x2468ACE13579BDF2468ACE13579BDF % br.R
    t0 = time.perf_counter()
    pp = zeromorph.trim(kzg.univariate_setup(bench.g1_generator(np), _to_mont(s), n), n)
    out = {"k": k, "reps": reps, "srs_setup_on_device_s": time.perf_counter() - t0}
    poly_h = pk.random_scalars(n, seed=1234)
    poly = pk.ResidentScalars(poly_h)

    class Timed(zeromorph.GpuOps):
        """GpuOps with a wall clock per operation (every entry point synchronises before it returns)."""
        spans = {}

    def timed(name):
        fn = getattr(zeromorph.GpuOps, name)

        def wrapper(*a):
            t = time.perf_counter()
            r = fn(*a)
            Timed.spans[name] = Timed.spans.get(name, 0.0) + (time.perf_counter() - t) * 1e3
            return r

        return staticmethod(wrapper)

    for name in ("quotients", "commit_quotients", "q_hat", "f", "div_linear", "commit"):
        setattr(Timed, name, timed(name))

    def run():
        Timed.spans = {}
        t = Keccak256Transcript()
        comm = zeromorph.commit(pp, poly)
        t.write_commitment(comm)
        point = t.squeeze_challenges(k)
        t.write_field_element(0)  # stands for the evaluation (written before open; open does not depend on it)
        t1 = time.perf_counter()
        value = zeromorph.open(pp, poly, point, 0, t, Timed)
        return comm, point, value, t.into_proof(), (time.perf_counter() - t1) * 1e3, dict(Timed.spans)

    (comm, point, value, proof, _, _), tm = bench.timed_reps(run, reps)
    opens = [run() for _ in range(reps)]
    best = min(opens, key=lambda r: r[4])
    out["commit_plus_open_ms"] = {"min": tm["ms_min"], "median": tm["ms_median"]}
    out["open_ms_min"] = best[4]
    out["open_phases_ms"] = {a: round(b, 3) for a, b in best[5].items()}
    # the opening against the verifier's equation; the evaluation the verifier uses is the remainder the prover returned,
    # cross-checked with the oracle's multilinear evaluation
    want = _to_int(po.evaluate_multilinear(poly_h, zr.mont_rows(point), po.host_threads()))
    assert value == want, "quotients' remainder differs from the oracle's evaluation"
    pts = [as_limbs((int.from_bytes(proof[i:i + 32], "big"), int.from_bytes(proof[i + 32:i + 64], "big"))) for i in range(96, len(proof), 64)]
    v = Keccak256Transcript()
    v.write_commitment(comm)
    v.squeeze_challenges(k)
    v.write_field_element(0)
    zr.verify_in_g1(comm, point, value, pts[:k], pts[k], pts[k + 1], v, s, 0)
    out["open_parity_checked"] = True
    poly.release()
    if with_prove:
        instances, preprocess, witness, sigma = bench.synth_vanilla_plonk_circuit(pk, po, np, k, seed=610)
        info = hyperplonk.vanilla_plonk_circuit_info(k, k, preprocess, [[(6, 1)], [(7, 1)], [(8, 1)]])
        hpp, hvp = hyperplonk.preprocess(pp, info, permutation_columns=sigma)

        class Circuit:
            def instances(self):
                return [instances]

            def synthesize(self, rnd, challenges):
                return witness

        phases = []

        def prove():
            t, marks = Keccak256Transcript(), []
            hyperplonk.prove(hpp, Circuit(), t, marks)
            phases.append({b_[0]: round((b_[1] - a_[1]) * 1e3, 2) for a_, b_ in zip(marks, marks[1:])})
            return t.into_proof()

        proof, tp = bench.timed_reps(prove, reps)
        affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
        pcs_verify = lambda reader, c, pt, val: zr.verify_reader_in_g1(reader, c, pt, val, s)  # noqa: E731
        ref.verify_reference(po.keccak256, None, k, instances, [affine(c) for c in hvp.preprocess_comms], [affine(c) for _, c in hvp.permutation_comms],
                             proof, pcs_verify=pcs_verify)
        out["hyperplonk_prove_zeromorph"] = {"gpu_ms": tp["ms_min"], "gpu_ms_median": tp["ms_median"], "proof_bytes": len(proof), "phases_ms": phases[-1],
                                             "verifier_accepts": True}
        hpp.release()
    pp.release()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
