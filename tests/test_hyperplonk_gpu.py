"""GPU parity tests for the full HyperPlonk vanilla_plonk prover (plonkish_b200/hyperplonk.py over the C ABI):
  * plonkish_cuda_fr_affine_table / plonkish_cuda_fr_evaluate against Python integers;
  * the compiled zero check (permutation constraint with the rotated z included) round by round against the tree;
  * HyperPlonk::prove (backend/hyperplonk.rs:164-291): proof BYTES identical to the all-integer reference prover with the
    oracle's MSM (tests/hyperplonk_ref.py), and the GPU's proof accepted by the integer restatement of the reference's
    verifier (hyperplonk.rs:293-362), a flipped witness rejected."""
import numpy as np
import pytest

import hyperplonk_ref as ref
from oracle import bigint_ref as br
from test_hyperplonk_cpu import affine_table_python

pytestmark = pytest.mark.gpu
R = br.R


@pytest.fixture(scope="module")
def pk():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import plonkish_b200

    return plonkish_b200


def _fe(rng):
    return int.from_bytes(rng.bytes(40), "little") % R


def _ints(arr):
    rinv = pow(1 << 256, -1, R)
    return [int.from_bytes(row.tobytes(), "little") * rinv % R for row in np.asarray(arr).reshape(-1, 4)]


@pytest.mark.parametrize("k", [1, 5, 11])
def test_affine_table_matches_python_integers(pk, k):
    rng = np.random.default_rng(100 + k)
    n = 1 << k
    for count, use_const, use_id, nrows in [(0, False, False, 2), (0, True, True, 0), (1, False, False, 0), (3, True, False, 1), (8, True, True, 3),
                                            (11, True, True, 2)]:
        polys = [[_fe(rng) for _ in range(n)] for _ in range(count)]
        coeffs = [1 if i == 1 else _fe(rng) for i in range(count)]
        rotations = [int(rng.integers(-min(k, 3), min(k, 3) + 1)) for _ in range(count)]
        constant, id_coeff = (_fe(rng) if use_const else 0), (_fe(rng) if use_id else 0)
        rows = [int(r_) for r_ in rng.choice(n, size=min(nrows, n), replace=False)]
        values = [_fe(rng) for _ in rows]
        resident = [pk.ResidentScalars(ref.mont_rows(p)) for p in polys]
        out = pk.fr_affine_table(k, resident, ref.mont_rows(coeffs) if count else None, rotations, constant=ref.to_mont(constant) if use_const else None,
                                 identity_coeff=ref.to_mont(id_coeff) if use_id else None, sparse_rows=rows,
                                 sparse_values=ref.mont_rows(values) if rows else None)
        want = affine_table_python(k, polys, coeffs, rotations, constant, id_coeff, rows, values)
        assert out.to_host().tobytes() == ref.mont_rows(want).tobytes()
        for r_ in resident + [out]:
            r_.release()


def test_affine_table_refuses_bad_arguments(pk):
    p = pk.ResidentScalars(ref.mont_rows([1, 2, 3, 4]))
    with pytest.raises(pk.PlonkishCudaError):
        pk.fr_affine_table(3, [p], ref.mont_rows([1]))                       # 4 evaluations are not 2^3
    with pytest.raises(pk.PlonkishCudaError):
        pk.fr_affine_table(2, [p], ref.mont_rows([1]), rotations=[3])        # rotation distance above num_vars (classic.rs:42)
    with pytest.raises(pk.PlonkishCudaError):
        pk.fr_affine_table(2, sparse_rows=[4], sparse_values=ref.mont_rows([1]))
    p.release()


@pytest.mark.parametrize("k", [0, 1, 7, 13])
def test_evaluate_matches_python_integers(pk, k):
    rng = np.random.default_rng(200 + k)
    table = [_fe(rng) for _ in range(1 << k)]
    pts = [[_fe(rng) for _ in range(k)], [int(rng.integers(2)) for _ in range(k)], [0] * k, [1] * k]
    poly = pk.ResidentScalars(ref.mont_rows(table))
    got = pk.fr_evaluate(poly, np.stack([ref.mont_rows(pt) for pt in pts]) if k else np.zeros((len(pts), 0, 4), dtype=np.uint64))
    assert _ints(got) == [ref.evaluate_multilinear(table, pt) for pt in pts]
    poly.release()


def _setup(pk, oracle, k, rng):
    from plonkish_b200 import kzg

    ss = [_fe(rng) for _ in range(k)]
    pp = kzg.setup(oracle.generator(), ref.mont_rows(ss))
    return ss, pp


class _Circuit:
    """PlonkishCircuit (backend.rs:125-133) over fixed witness columns, like backend.rs:143-181 (MockCircuit)."""

    def __init__(self, instances, witness):
        self._instances, self._witness = [list(instances)], [ref.mont_rows(w) for w in witness]

    def instances(self):
        return self._instances

    def synthesize(self, rnd, challenges):
        assert rnd == 0 and not challenges
        return self._witness


def _prove_on_gpu(pk, pp, k, instances, preprocess, witness, cycles):
    from plonkish_b200 import hyperplonk
    from plonkish_b200.transcript import Keccak256Transcript

    info = hyperplonk.vanilla_plonk_circuit_info(k, len(instances), [ref.mont_rows(p) for p in preprocess], cycles)
    hpp, hvp = hyperplonk.preprocess(pp, info)
    t = Keccak256Transcript()
    hyperplonk.prove(hpp, _Circuit(instances, witness), t)
    hpp.release()
    return t.into_proof(), hvp


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
def test_proof_bytes_match_the_integer_reference_prover(pk, oracle, k):
    from plonkish_b200.transcript import Keccak256Transcript

    rng = np.random.default_rng(300 + k)
    ss, pp = _setup(pk, oracle, k, rng)
    eqs_host = [pp.eq(i).to_host() for i in range(k + 1)]
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_circuit(k, rng)
    proof, hvp = _prove_on_gpu(pk, pp, k, instances, preprocess, witness, cycles)
    # the reference side: Python integers + the oracle's MSM
    commit = lambda f: oracle.variable_base_msm(ref.mont_rows(f), eqs_host[k])  # noqa: E731
    sigmas = ref.permutation_polys(k, [6, 7, 8], cycles)
    t = Keccak256Transcript()
    ref.prove_reference(commit, ref.oracle_batch_open(oracle, eqs_host, k), k, instances, preprocess, witness, sigmas, t)
    assert proof == t.into_proof()
    # preprocess commitments of the GPU path are the oracle's
    assert all((c == commit(p)).all() for c, p in zip(hvp.preprocess_comms, preprocess))
    assert all((c == commit(s)).all() for (_, c), s in zip(hvp.permutation_comms, sigmas))
    pp.release()


@pytest.mark.parametrize("k", [7, 10, 13])
def test_gpu_proof_is_accepted_by_the_reference_verifier(pk, oracle, k):
    rng = np.random.default_rng(400 + k)
    ss, pp = _setup(pk, oracle, k, rng)
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_circuit(k, rng)
    proof, hvp = _prove_on_gpu(pk, pp, k, instances, preprocess, witness, cycles)
    affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
    pre = [affine(c) for c in hvp.preprocess_comms]
    perm = [affine(c) for _, c in hvp.permutation_comms]
    ref.verify_reference(oracle.keccak256, ss, k, instances, pre, perm, proof)
    assert len(proof) == 4 * 64 + k * 6 * 32 + 14 * 32 + k * 3 * 32 + k * 64
    # an unsatisfied gate: the prover still runs (it does not check the witness, like the reference without sanity-check),
    # the verifier refuses the proof
    bad = [list(w) for w in witness]
    bad[2][5] = (bad[2][5] + 1) % R
    bad_proof, _ = _prove_on_gpu(pk, pp, k, instances, preprocess, bad, cycles)
    with pytest.raises(AssertionError):
        ref.verify_reference(oracle.keccak256, ss, k, instances, pre, perm, bad_proof)
    # a broken copy constraint likewise (row 3 of w_l is copied or free; changing w_l and repairing the gate through q_c is
    # not possible without touching the preprocessed side, so flip a cell that sits in a cycle when there is one)
    cell = next(((p, r_) for cyc in cycles if len(cyc) > 1 for p, r_ in cyc if p == 8), None)
    if cell is not None:
        bad = [list(w) for w in witness]
        bad[2][cell[1]] = (bad[2][cell[1]] + 1) % R
        bad_proof, _ = _prove_on_gpu(pk, pp, k, instances, preprocess, bad, cycles)
        with pytest.raises(AssertionError):
            ref.verify_reference(oracle.keccak256, ss, k, instances, pre, perm, bad_proof)
    pp.release()


def _lookup_info(k, preprocess, cycles, num_instances):
    from plonkish_b200 import hyperplonk
    from plonkish_b200.expression import Expression

    pi, q_l, q_r, q_m, q_o, q_c, q_lookup, t_l, t_r, t_o, w_l, w_r, w_o = (Expression.polynomial(p) for p in range(13))
    return hyperplonk.PlonkishCircuitInfo(
        k=k, num_instances=[num_instances], preprocess_polys=[ref.mont_rows(p) for p in preprocess], num_witness_polys=[3], num_challenges=[0],
        constraints=[q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi],
        lookups=[[(q_lookup * w_l, t_l), (q_lookup * w_r, t_r), (q_lookup * w_o, t_o)]], permutations=cycles, max_degree=4)   # util.rs:63-86


def _prove_lookup_on_gpu(pk, pp, k, instances, preprocess, witness, cycles):
    from plonkish_b200 import hyperplonk
    from plonkish_b200.transcript import Keccak256Transcript

    hpp, hvp = hyperplonk.preprocess(pp, _lookup_info(k, preprocess, cycles, len(instances)))
    t = Keccak256Transcript()
    try:
        hyperplonk.prove(hpp, _Circuit(instances, witness), t)
    finally:
        hpp.release()
    return t.into_proof(), hvp


@pytest.mark.parametrize("k", [1, 6, 12])
def test_lookup_producers_match_python_integers(pk, oracle, k):
    """plonkish_cuda_fr_expression_table / _lookup_m_poly / _lookup_h_poly (prover.rs:50-250) against Python integers:
    repeated table values count at their last row, an input outside the table is refused."""
    rng = np.random.default_rng(500 + k)
    n = 1 << k
    polys = [[_fe(rng) for _ in range(n)] for _ in range(4)]
    resident = [pk.ResidentScalars(ref.mont_rows(p)) for p in polys]
    terms = [(ref.to_mont(3), [0, 1]), (ref.to_mont(1), [2]), (ref.to_mont(R - 7), [])]
    for common in (-1, 3):
        out = pk.fr_expression_table(resident, terms, common)
        want = [((3 * polys[0][b] * polys[1][b] + polys[2][b] - 7) * (polys[3][b] if common >= 0 else 1)) % R for b in range(n)]
        assert out.to_host().tobytes() == ref.mont_rows(want).tobytes()
        out.release()
    table = ([0, 0] + [_fe(rng) for _ in range(n - 2)]) if n > 2 else [0] * n
    if n > 4:
        table[n - 1] = table[3]
    inputs = [table[int(i)] for i in rng.integers(0, n, n)]
    index = {v: i for i, v in enumerate(table)}
    want_m = [0] * n
    for v in inputs:
        want_m[index[v]] += 1
    tab, inp = pk.ResidentScalars(ref.mont_rows(table)), pk.ResidentScalars(ref.mont_rows(inputs))
    m = pk.lookup_m_poly(inp, tab)
    assert m.to_host().tobytes() == ref.mont_rows(want_m).tobytes()
    gamma = _fe(rng)
    h = pk.lookup_h_poly(inp, tab, m, ref.to_mont(gamma))
    assert h.to_host().tobytes() == ref.mont_rows(ref.lookup_h_python(inputs, table, want_m, gamma)).tobytes()
    bad = list(inputs)
    bad[n // 2] = next(v for v in (12345, 12346, 12347) if v not in index)
    bad_r = pk.ResidentScalars(ref.mont_rows(bad))
    with pytest.raises(pk.PlonkishCudaError, match="Invalid lookup input"):
        pk.lookup_m_poly(bad_r, tab)
    for r_ in resident + [tab, inp, m, h, bad_r]:
        r_.release()


def _zeromorph_setup(pk, oracle, k, s):
    from plonkish_b200 import kzg, zeromorph

    pp = zeromorph.trim(kzg.univariate_setup(oracle.generator(), ref.to_mont(s), 1 << k), 1 << k)
    return pp, pp.commit_pp.to_host()


def _zeromorph_reference(oracle, srs_h, k):
    import zeromorph_ref as zr
    from batch_open_ref import batch_open_reference

    commit = lambda f: zr.commit_coeffs(oracle, srs_h, f)  # noqa: E731

    def open_ref(g_prime, challenges, transcript):
        return zr.open_reference(oracle, srs_h, srs_h, g_prime, challenges, 0, transcript)

    def batch_open(polys, points, evals, transcript):
        return batch_open_reference(oracle, None, k, polys, points, evals, transcript, open_fn=open_ref)

    return commit, batch_open


@pytest.mark.parametrize("k", [2, 3, 5])
def test_zeromorph_proof_bytes_match_the_integer_reference_prover(pk, oracle, k):
    """HyperPlonk<Zeromorph<UnivariateKzg<Bn256>>> (the reference's zeromorph_kzg tests, backend/hyperplonk.rs:426): the same
    prover over the other BN254 multilinear PCS — proof bytes identical to the all-integer prover with Zeromorph's commit
    and open restated in tests/zeromorph_ref.py."""
    from plonkish_b200.transcript import Keccak256Transcript

    rng = np.random.default_rng(900 + k)
    pp, srs_h = _zeromorph_setup(pk, oracle, k, 0x5EED5EED5EED1234567)
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_circuit(k, rng)
    proof, hvp = _prove_on_gpu(pk, pp, k, instances, preprocess, witness, cycles)
    commit, batch_open = _zeromorph_reference(oracle, srs_h, k)
    sigmas = ref.permutation_polys(k, [6, 7, 8], cycles)
    t = Keccak256Transcript()
    ref.prove_reference(commit, batch_open, k, instances, preprocess, witness, sigmas, t)
    assert proof == t.into_proof()
    assert all((c == commit(p)).all() for c, p in zip(hvp.preprocess_comms, preprocess))
    pp.release()


@pytest.mark.parametrize("k", [8, 12])
def test_zeromorph_gpu_proof_is_accepted_by_the_reference_verifier(pk, oracle, k):
    # HyperPlonk::verify with Zeromorph::verify (zeromorph.rs:216-245) as the last step, in G1 with the setup's trapdoor
    import zeromorph_ref as zr

    rng = np.random.default_rng(980 + k)
    s = 0x1357924680ACE1357924680ACE % br.R
    pp, _ = _zeromorph_setup(pk, oracle, k, s)
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_circuit(k, rng)
    proof, hvp = _prove_on_gpu(pk, pp, k, instances, preprocess, witness, cycles)
    affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
    pre = [affine(c) for c in hvp.preprocess_comms]
    perm = [affine(c) for _, c in hvp.permutation_comms]
    pcs_verify = lambda reader, comm, point, value: zr.verify_reader_in_g1(reader, comm, point, value, s)  # noqa: E731
    ref.verify_reference(oracle.keccak256, None, k, instances, pre, perm, proof, pcs_verify=pcs_verify)
    assert len(proof) == 4 * 64 + k * 6 * 32 + 14 * 32 + k * 3 * 32 + (k + 2) * 64
    bad = [list(w) for w in witness]
    bad[2][5] = (bad[2][5] + 1) % br.R
    bad_proof, _ = _prove_on_gpu(pk, pp, k, instances, preprocess, bad, cycles)
    with pytest.raises(AssertionError):
        ref.verify_reference(oracle.keccak256, None, k, instances, pre, perm, bad_proof, pcs_verify=pcs_verify)
    pp.release()


@pytest.mark.parametrize("k", [2, 4])
def test_gemini_proof_bytes_match_the_integer_reference_prover(pk, oracle, k):
    """HyperPlonk<Gemini<UnivariateKzg<Bn256>>> (the reference's gemini_kzg tests, backend/hyperplonk.rs:425): proof bytes
    identical to the all-integer prover whose commit is the oracle's MSM and whose final opening is Gemini::open's host
    logic with every polynomial operation done by the oracle on the CPU."""
    import zeromorph_ref as zr
    from batch_open_ref import batch_open_reference
    from plonkish_b200 import gemini, kzg
    from plonkish_b200.transcript import Keccak256Transcript
    from test_gemini_cpu import _oracle_ops

    rng = np.random.default_rng(1100 + k)
    pp = gemini.GeminiKzgProverParam(kzg.univariate_setup(oracle.generator(), ref.to_mont(0x600DF00D600DF00D), 1 << k))
    srs_h = pp.powers_of_s_g1.to_host()
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_circuit(k, rng)
    proof, hvp = _prove_on_gpu(pk, pp, k, instances, preprocess, witness, cycles)
    commit = lambda f: zr.commit_coeffs(oracle, srs_h, f)  # noqa: E731
    ops = _oracle_ops(oracle)

    def open_ref(g_prime, challenges, transcript):
        gemini.open(gemini.GeminiKzgProverParam(srs_h), zr.mont_rows(g_prime), challenges, transcript, ops)

    batch_open = lambda polys, points, evals, transcript: batch_open_reference(oracle, None, k, polys, points, evals, transcript, open_fn=open_ref)  # noqa: E731
    sigmas = ref.permutation_polys(k, [6, 7, 8], cycles)
    t = Keccak256Transcript()
    ref.prove_reference(commit, batch_open, k, instances, preprocess, witness, sigmas, t)
    assert proof == t.into_proof()
    pp.release()


@pytest.mark.parametrize("k", [8, 11])
def test_gemini_gpu_proof_is_accepted_by_the_reference_verifier(pk, oracle, k):
    # HyperPlonk::verify with Gemini::verify (gemini.rs:168-197) as the last step, in G1 with the setup's trapdoor
    import gemini_ref as gr
    from plonkish_b200 import gemini, kzg

    rng = np.random.default_rng(1200 + k)
    s = 0xFACEFEED0BADF00D123 % br.R
    pp = gemini.GeminiKzgProverParam(kzg.univariate_setup(oracle.generator(), ref.to_mont(s), 1 << k))
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_circuit(k, rng)
    proof, hvp = _prove_on_gpu(pk, pp, k, instances, preprocess, witness, cycles)
    affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
    pre = [affine(c) for c in hvp.preprocess_comms]
    perm = [affine(c) for _, c in hvp.permutation_comms]
    pcs_verify = lambda reader, comm, point, value: gr.verify_reader_in_g1(reader, comm, point, value, s)  # noqa: E731
    ref.verify_reference(oracle.keccak256, None, k, instances, pre, perm, proof, pcs_verify=pcs_verify)
    bad = [list(w) for w in witness]
    bad[2][5] = (bad[2][5] + 1) % br.R
    bad_proof, _ = _prove_on_gpu(pk, pp, k, instances, preprocess, bad, cycles)
    with pytest.raises(AssertionError):
        ref.verify_reference(oracle.keccak256, None, k, instances, pre, perm, bad_proof, pcs_verify=pcs_verify)
    pp.release()


@pytest.mark.parametrize("k", [3, 4])
def test_zeromorph_lookup_proof_bytes_match_the_integer_reference_prover(pk, oracle, k):
    from plonkish_b200.transcript import Keccak256Transcript

    rng = np.random.default_rng(950 + k)
    pp, srs_h = _zeromorph_setup(pk, oracle, k, 0xABCDEF987654321)
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_with_lookup_circuit(k, rng)
    proof, _ = _prove_lookup_on_gpu(pk, pp, k, instances, preprocess, witness, cycles)
    commit, batch_open = _zeromorph_reference(oracle, srs_h, k)
    sigmas = ref.permutation_polys(k, [10, 11, 12], cycles)
    t = Keccak256Transcript()
    ref.prove_reference_lookup(commit, batch_open, k, instances, preprocess, witness, sigmas, t)
    assert proof == t.into_proof()
    pp.release()


@pytest.mark.parametrize("k", [3, 4, 5])
def test_lookup_proof_bytes_match_the_integer_reference_prover(pk, oracle, k):
    """HyperPlonk::prove for vanilla_plonk_with_lookup (the reference's second test circuit, backend/hyperplonk.rs:371-427):
    proof bytes identical to the all-integer prover, preprocess commitments identical to the oracle's."""
    from plonkish_b200.transcript import Keccak256Transcript

    rng = np.random.default_rng(600 + k)
    ss, pp = _setup(pk, oracle, k, rng)
    eqs_host = [pp.eq(i).to_host() for i in range(k + 1)]
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_with_lookup_circuit(k, rng)
    proof, hvp = _prove_lookup_on_gpu(pk, pp, k, instances, preprocess, witness, cycles)
    commit = lambda f: oracle.variable_base_msm(ref.mont_rows(f), eqs_host[k])  # noqa: E731
    sigmas = ref.permutation_polys(k, [10, 11, 12], cycles)
    t = Keccak256Transcript()
    ref.prove_reference_lookup(commit, ref.oracle_batch_open(oracle, eqs_host, k), k, instances, preprocess, witness, sigmas, t)
    assert proof == t.into_proof()
    pp.release()


@pytest.mark.parametrize("k", [8, 11])
def test_gpu_lookup_proof_is_accepted_by_the_reference_verifier(pk, oracle, k):
    rng = np.random.default_rng(700 + k)
    ss, pp = _setup(pk, oracle, k, rng)
    instances, preprocess, witness, cycles = ref.rand_vanilla_plonk_with_lookup_circuit(k, rng)
    proof, hvp = _prove_lookup_on_gpu(pk, pp, k, instances, preprocess, witness, cycles)
    affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
    pre = [affine(c) for c in hvp.preprocess_comms]
    perm = [affine(c) for _, c in hvp.permutation_comms]
    ref.verify_reference_lookup(oracle.keccak256, ss, k, instances, pre, perm, proof)
    assert len(proof) == 6 * 64 + k * 6 * 32 + 20 * 32 + k * 3 * 32 + k * 64
    # a witness value that is not in the table: lookup_m_polys fails like the reference (Error::InvalidSnark, prover.rs:176-178)
    row = next(b for b in range(1 << k) if preprocess[5][b] == 1)
    bad = [list(w) for w in witness]
    bad[0][row] = (bad[0][row] + 1) % R
    with pytest.raises(pk.PlonkishCudaError, match="Invalid lookup input"):
        _prove_lookup_on_gpu(pk, pp, k, instances, preprocess, bad, cycles)
    pp.release()
