"""The two multilinear PCSs over the univariate KZG SRS on one GPU — Zeromorph and Gemini (pcs/multilinear/zeromorph.rs,
gemini.rs): commit + open of one 2^k-evaluation polynomial, checked against the scheme's verifier equation (tests/
zeromorph_ref.py / gemini_ref.py: G1 with the setup's trapdoor), and HyperPlonk::prove for vanilla_plonk over the scheme
on bench.py's synthetic circuit, accepted by the verifier restatement:
python tools/pcs_bench.py [k] [reps] [prove: 0|1] [zeromorph|gemini|both]."""
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


def timed_ops(base, names):
    """`base` (an ops class) with a wall clock per operation: every entry point synchronises before it returns."""

    class Timed(base):
        spans = {}

    def wrap(name):
        fn = getattr(base, name)

        def wrapper(*a):
            t = time.perf_counter()
            r = fn(*a)
            Timed.spans[name] = Timed.spans.get(name, 0.0) + (time.perf_counter() - t) * 1e3
            return r

        return staticmethod(wrapper)

    for name in names:
        setattr(Timed, name, wrap(name))
    return Timed


def main():
    import gemini_ref as gr
    import hyperplonk_ref as ref
    import zeromorph_ref as zr
    from oracle import bigint_ref as br
    from oracle import pyoracle as po
    from plonkish_b200 import gemini, hyperplonk, kzg, zeromorph
    from plonkish_b200.sumcheck import _to_int, _to_mont
    from plonkish_b200.transcript import Keccak256Transcript

    k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with_prove = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
    which = sys.argv[4] if len(sys.argv) > 4 else "both"
    torch.cuda.init()
    n, s = 1 << k, 0x2468ACE13579BDF2468ACE13579BDF % br.R
    t0 = time.perf_counter()
    powers = kzg.univariate_setup(bench.g1_generator(np), _to_mont(s), n)
    out = {"k": k, "reps": reps, "srs_setup_on_device_s": time.perf_counter() - t0}
    poly_h = pk.random_scalars(n, seed=1234)
    poly = pk.ResidentScalars(poly_h)
    circuit_parts = bench.synth_vanilla_plonk_circuit(pk, po, np, k, seed=610) if with_prove else None
    schemes = {
        "zeromorph": (zeromorph, zeromorph.trim(powers, n), ("quotients", "commit_quotients", "q_hat", "f", "div_linear", "commit"),
                      lambda reader, c, pt, val: zr.verify_reader_in_g1(reader, c, pt, val, s)),
        "gemini": (gemini, gemini.GeminiKzgProverParam(powers), ("folds", "commit_folds", "evaluate", "linear_combination", "div_linear", "commit"),
                   lambda reader, c, pt, val: gr.verify_reader_in_g1(reader, c, pt, val, s)),
    }
    for name, (mod, pp, op_names, pcs_verify) in schemes.items():
        if which not in ("both", name):
            continue
        ops = timed_ops(mod.GpuOps, op_names)

        def run():
            ops.spans = {}
            t = Keccak256Transcript()
            comm = mod.commit(pp, poly)
            t.write_commitment(comm)
            point = t.squeeze_challenges(k)
            value = None
            t1 = time.perf_counter()
            if mod is zeromorph:
                t.write_field_element(0)       # stands for the evaluation (written before open; the proof does not depend on it)
                value = mod.open(pp, poly, point, 0, t, ops)
            else:
                t.write_field_element(0)
                mod.open(pp, poly, point, t, ops)
            return point, value, t.into_proof(), (time.perf_counter() - t1) * 1e3, dict(ops.spans)

        runs = []
        _, tm = bench.timed_reps(lambda: runs.append(run()), reps)
        best = min(runs[1:], key=lambda r_: r_[3])
        point, value, proof = best[0], best[1], best[2]
        res = {"commit_plus_open_ms": {"min": tm["ms_min"], "median": tm["ms_median"]}, "open_ms_min": best[3],
               "open_ops_ms": {a: round(b, 3) for a, b in best[4].items()}, "proof_bytes": len(proof) - 96}
        # the opening against the verifier's equation, with the oracle's multilinear evaluation as the claimed value
        want = _to_int(po.evaluate_multilinear(poly_h, zr.mont_rows(point), po.host_threads()))
        assert value is None or value == want, "quotients' remainder differs from the oracle's evaluation"
        reader = ref.ProofReader(po.keccak256, proof)
        c = reader.read_commitment()
        assert reader.squeeze_challenges(k) == point
        reader.read_field_element()
        pcs_verify(reader, c, point, want)
        assert reader.pos == len(proof)
        res["open_parity_checked"] = True
        if with_prove:
            instances, preprocess, witness, sigma = circuit_parts
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

            hp_proof, tp = bench.timed_reps(prove, reps)
            affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
            ref.verify_reference(po.keccak256, None, k, instances, [affine(c_) for c_ in hvp.preprocess_comms],
                                 [affine(c_) for _, c_ in hvp.permutation_comms], hp_proof, pcs_verify=pcs_verify)
            res["hyperplonk_prove"] = {"gpu_ms": tp["ms_min"], "gpu_ms_median": tp["ms_median"], "proof_bytes": len(hp_proof), "phases_ms": phases[-1],
                                       "verifier_accepts": True}
            hpp.release()
        out[name] = res
    poly.release()
    powers.release()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
