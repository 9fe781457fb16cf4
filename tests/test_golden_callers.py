"""The committed known-answer vectors for the rows either side of the MSM (tests/golden/caller_vectors.json, made
with Python integers by tests/golden/make_golden_callers.py): the oracle on CPU, the CUDA path on the GPU."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def vec():
    with open(os.path.join(ROOT, "tests", "golden", "caller_vectors.json")) as f:
        return json.load(f)


def fe(hexes):
    if isinstance(hexes, str):
        return np.frombuffer(bytes.fromhex(hexes), dtype=np.uint64).copy()
    return np.frombuffer(bytes.fromhex("".join(hexes)), dtype=np.uint64).reshape(len(hexes), -1).copy()


def terms_of(case):
    return [(c, idx) for c, idx in zip(fe(case["coeffs"]), case["terms"])]


# ---------------------------------------------------------------------- oracle
def test_oracle_matches_the_kzg_vectors(oracle, vec):
    g = oracle.generator()
    for c in vec["kzg"]:
        k = c["num_vars"]
        ss = fe(c["ss"]) if k else np.zeros((0, 4), dtype=np.uint64)
        eqs = oracle.kzg_eq_scalars(ss) if k else [fe(c["eq_scalars"][0])]
        for i in range(k + 1):
            assert np.ascontiguousarray(eqs[i]).tobytes() == fe(c["eq_scalars"][i]).tobytes()
            assert oracle.fixed_base_msm(g, eqs[i]).tobytes() == fe(c["eq_points"][i]).tobytes()
        assert oracle.variable_base_msm(fe(c["evals"]), fe(c["eq_points"][k])).tobytes() == fe(c["commitment"]).tobytes()
        if k:
            qs, value = oracle.quotients(fe(c["evals"]), fe(c["point"]))
            assert value.tobytes() == fe(c["eval"]).tobytes()
            for i in range(k):
                assert qs[i].tobytes() == fe(c["quotients"][i]).tobytes()
                assert oracle.variable_base_msm(qs[i], fe(c["eq_points"][i])).tobytes() == fe(c["quotient_commitments"][i]).tobytes()
    for c in vec["merge"]:
        assert oracle.fr_linear_combination([fe(p) for p in c["polys"]], fe(c["coeffs"])).tobytes() == fe(c["result"]).tobytes()
    fb = vec["fixed_base"]
    for window in (3, 7):
        assert oracle.fixed_base_msm(fe(fb["base"]), fe(fb["scalars"]), window=window).tobytes() == fe(fb["points"]).tobytes()


def test_oracle_matches_the_sumcheck_vectors(oracle, vec):
    for c in vec["sumcheck"]:
        cur = [fe(p) for p in c["polys"]]
        for r in c["rounds"]:
            assert oracle.sumcheck_round(cur, terms_of(c), c["common"]).tobytes() == fe(r["evals_1_to_degree"]).tobytes()
            cur = [oracle.fix_var(p, fe(r["challenge"])) for p in cur]
        assert np.stack([p[0] for p in cur]).tobytes() == fe(c["final_evals"]).tobytes()


# ------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def pk():
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import plonkish_b200

    plonkish_b200._lib.lib()
    return plonkish_b200


@pytest.mark.gpu
def test_cuda_matches_the_kzg_vectors(pk, vec):
    from plonkish_b200 import kzg

    g = np.frombuffer(b"".join((v * (1 << 256) % kzg_p()).to_bytes(32, "little") for v in (1, 2)), dtype=np.uint64).copy()
    for c in vec["kzg"]:
        k = c["num_vars"]
        ss = fe(c["ss"]) if k else np.zeros((0, 4), dtype=np.uint64)
        pp = kzg.setup(g, ss)
        for i in range(k + 1):
            assert pp.eq(i).to_host().tobytes() == fe(c["eq_points"][i]).tobytes()
        poly = pk.ResidentScalars(fe(c["evals"]))
        assert kzg.commit(pp, poly).tobytes() == fe(c["commitment"]).tobytes()
        assert kzg.commit(pp, fe(c["evals"])).tobytes() == fe(c["commitment"]).tobytes()
        comms, value = kzg.open_resident(pp, poly, fe(c["point"]) if k else np.zeros((0, 4), dtype=np.uint64))
        assert value.tobytes() == fe(c["eval"]).tobytes()
        assert [x.tobytes() for x in comms] == [fe(q).tobytes() for q in c["quotient_commitments"]]
        poly.release()
        pp.release()
    for c in vec["merge"]:
        res = [pk.ResidentScalars(fe(p)) for p in c["polys"]]
        merged = pk.fr_linear_combination(res, fe(c["coeffs"]))
        assert merged.to_host().tobytes() == fe(c["result"]).tobytes()
        for r in res + [merged]:
            r.release()
    fb = vec["fixed_base"]
    assert pk.fixed_base_msm(fe(fb["base"]), fe(fb["scalars"])).tobytes() == fe(fb["points"]).tobytes()


def kzg_p():
    return 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


@pytest.mark.gpu
def test_cuda_matches_the_sumcheck_vectors(pk, vec):
    from plonkish_b200.sumcheck import SumCheckProver

    for c in vec["sumcheck"]:
        res = [pk.ResidentScalars(fe(p)) for p in c["polys"]]
        prover = SumCheckProver(res, terms_of(c), c["common"])
        assert prover.degree == c["degree"]
        for r in c["rounds"]:
            assert prover.round_evals().tobytes() == fe(r["evals_1_to_degree"]).tobytes()
            prover.fix_var(fe(r["challenge"]))
        assert prover.final_evals().tobytes() == fe(c["final_evals"]).tobytes()
        prover.free()
        for r in res:
            r.release()
