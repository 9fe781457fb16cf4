"""cProfile of HyperPlonk::prove at a small k (GPU work is negligible there: what is left is the host side — ctypes calls,
launch + sync round trips, Python integer arithmetic, Keccak).  python tools/prove_host_profile.py [k]"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
import plonkish_b200 as pk
from oracle import pyoracle as po
from plonkish_b200 import hyperplonk, kzg
from plonkish_b200.transcript import Keccak256Transcript

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
torch.cuda.init()
pp = kzg.setup(bench.g1_generator(np), pk.random_scalars(k, seed=601))
instances, preprocess, witness, sigma = bench.synth_vanilla_plonk_circuit(pk, po, np, k, seed=610)
info = hyperplonk.vanilla_plonk_circuit_info(k, k, preprocess, [[(6, 1)], [(7, 1)], [(8, 1)]])
hpp, hvp = hyperplonk.preprocess(pp, info, permutation_columns=sigma)


class Circuit:
    def instances(self):
        return [instances]

    def synthesize(self, rnd, challenges):
        return witness


def run():
    t = Keccak256Transcript()
    hyperplonk.prove(hpp, Circuit(), t, [])
    return t.into_proof()


run(); run()
t0 = time.perf_counter(); run(); print(f"k={k}: prove {1e3 * (time.perf_counter() - t0):.2f} ms")
pr = cProfile.Profile(); pr.enable(); run(); run(); run(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
