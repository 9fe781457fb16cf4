"""The Keccak256 Fiat-Shamir transcript (plonkish_b200/transcript.py, mirror of util/transcript.rs:100-235): public
Keccak known answers, hashlib for the permutation and the sponge, the oracle's independent C Keccak, and — on the GPU —
proof bytes of a zero check followed by a KZG opening, produced once through the CUDA path and once through the oracle,
then replayed the way the verifier reads them (classic.rs:242-262)."""
import hashlib

import numpy as np
import pytest

from oracle import bigint_ref as br
from test_sumcheck_cpu import _expr_int, _ints, _mont

R = br.R


def test_keccak256_public_known_answers_and_hashlib(oracle):
    from plonkish_b200 import transcript as tr

    # Keccak-256 of "" and "abc": the digests every Ethereum toolchain publishes
    kats = {b"": "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
            b"abc": "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"}
    for msg, digest in kats.items():
        assert tr.keccak256(msg).hex() == digest
        assert oracle.keccak256(msg).hex() == digest
    rng = np.random.default_rng(1)
    # the library's host routine and the Python rounds are the same permutation
    from plonkish_b200 import _lib

    for _ in range(20):
        lanes = [int(v) for v in rng.integers(0, 1 << 64, 25, dtype=np.uint64)]
        arr = np.array(lanes, dtype=np.uint64)
        _lib.load().plonkish_cuda_keccak_f1600(arr.ctypes.data)
        assert [int(v) for v in arr] == tr.keccak_f1600(lanes)
    for n in (0, 1, 31, 32, 64, 135, 136, 137, 271, 272, 273, 1000, 5000):
        data = rng.bytes(n)
        assert tr.sha3_256(data) == hashlib.sha3_256(data).digest(), n   # same permutation and sponge, SHA-3 domain byte
        assert tr.keccak256(data) == oracle.keccak256(data), n           # two independent Keccak256 implementations


class OneShotTranscript:
    """The same transcript rules over the oracle's one-shot C Keccak: everything absorbed since the last squeeze is
    hashed at once (FiatShamirTranscript::squeeze_challenge re-seeds the state with the digest, transcript.rs:127-131)."""

    def __init__(self, oracle):
        self.oracle, self.absorbed, self.stream = oracle, b"", bytearray()

    def write_field_elements(self, values):
        for v in values:
            self.absorbed += (v % R).to_bytes(32, "little")
            self.stream += (v % R).to_bytes(32, "big")

    def write_commitments(self, points):
        rq = pow(br.MONT, -1, br.P)
        for p in points:
            x = int.from_bytes(np.ascontiguousarray(p[:4]).tobytes(), "little") * rq % br.P
            y = int.from_bytes(np.ascontiguousarray(p[4:]).tobytes(), "little") * rq % br.P
            self.absorbed += x.to_bytes(32, "little") + y.to_bytes(32, "little")
            self.stream += x.to_bytes(32, "big") + y.to_bytes(32, "big")

    def squeeze_challenge(self):
        h = self.oracle.keccak256(self.absorbed)
        self.absorbed = h
        return int.from_bytes(h, "little") % R


def test_transcript_streaming_matches_one_shot_hashing(oracle):
    from plonkish_b200.transcript import Keccak256Transcript

    rng = np.random.default_rng(2)
    a, b = Keccak256Transcript(), OneShotTranscript(oracle)
    g = oracle.generator()
    pts = [oracle.scalar_mul(g, int(k)) for k in rng.integers(1, 1 << 60, 5)]
    for step in range(6):
        vals = [int.from_bytes(rng.bytes(32), "little") % R for _ in range(int(rng.integers(0, 9)))]
        a.write_field_elements(vals); b.write_field_elements(vals)
        if step % 2:
            a.write_commitments(pts[: step]); b.write_commitments(pts[: step])
        assert a.squeeze_challenge() == b.squeeze_challenge()
        assert a.squeeze_challenge() == b.squeeze_challenge()  # squeezing twice in a row (squeeze_challenges, transcript.rs:18-20)
    assert a.into_proof() == bytes(b.stream)
    with pytest.raises(ValueError):
        a.write_commitment(np.zeros(8, dtype=np.uint64))  # the identity has no coordinates (transcript.rs:175-181)
    # a commitment is streamed as x || y big-endian: the bytes the oracle's transcript_bytes gives
    t = Keccak256Transcript()
    t.write_commitment(pts[0])
    assert t.into_proof() == oracle.transcript_bytes(pts[0])


@pytest.mark.gpu
def test_proof_bytes_of_zero_check_and_opening_match_the_oracle():
    import torch

    assert torch.cuda.is_available()
    import plonkish_b200 as pk
    from oracle import pyoracle as oracle
    from plonkish_b200 import kzg, sumcheck
    from plonkish_b200.transcript import Keccak256Transcript

    k = 9
    n = 1 << k
    rng = np.random.default_rng(5)
    a, b = oracle.random_scalars(n, 1), oracle.random_scalars(n, 2)
    ai, bi = _ints(a), _ints(b)
    c = _mont([x * y % R for x, y in zip(ai, bi)])
    ss = oracle.random_scalars(k, 3)
    pp = kzg.setup(oracle.generator(), ss)
    eqs_host = [pp.eq(i).to_host() for i in range(k + 1)]
    one, minus_one = _mont([1])[0], _mont([R - 1])[0]
    terms = [(one, [1, 2]), (minus_one, [3])]
    terms_int = [(1, [1, 2]), (R - 1, [3])]

    def eq_table(y):
        eq = np.array([1], dtype=object)
        for y_i in y:
            eq = np.concatenate([eq * (1 - y_i) % R, eq * y_i % R])
        return _mont(list(eq))

    # ---- prover on the GPU: commit, zero check eq(x, y) * (a*b - c), then open a at the sum-check point
    t_gpu = Keccak256Transcript()
    comms, resident = kzg.batch_commit(pp, [a, b, c], keep=True)
    t_gpu.write_commitments(comms)                                   # witness commitments (backend/hyperplonk.rs:201-202)
    y = t_gpu.squeeze_challenges(k)                                  # zero-check point (hyperplonk.rs:262)
    eq = pk.eq_table(_mont(y))                                       # MultilinearPolynomial::eq_xy(y) on the device (classic.rs:57-61)
    assert eq.to_host().tobytes() == eq_table(y).tobytes()
    challenges, evals = sumcheck.prove_to_transcript([eq] + resident, terms, 0, t_gpu, common=0)
    t_gpu.write_field_elements(evals[1:])                            # evaluations of the witness polynomials (hyperplonk.rs:279-285)
    value = kzg.open_to_transcript(pp, resident[0], _mont(challenges), t_gpu)
    assert _ints(value)[0] == evals[1]
    proof_gpu = t_gpu.into_proof()

    # ---- the same prover through the oracle
    t_cpu = OneShotTranscript(oracle)
    t_cpu.write_commitments([oracle.variable_base_msm(p, eqs_host[k]) for p in (a, b, c)])
    y2 = [t_cpu.squeeze_challenge() for _ in range(k)]
    assert y2 == y
    cur = [eq_table(y2), a, b, c]
    claim, chal2 = 0, []
    for _ in range(k):
        tail = _ints(oracle.sumcheck_round(cur, terms, 0))
        msg = [(claim - tail[0]) % R] + tail
        t_cpu.write_field_elements(msg)
        ch = t_cpu.squeeze_challenge()
        chal2.append(ch)
        claim = sumcheck.interpolate_at(msg, ch)
        cur = [oracle.fix_var(p, _mont([ch])[0]) for p in cur]
    finals = [_ints(p)[0] for p in cur]
    t_cpu.write_field_elements(finals[1:])
    qs, _ = oracle.quotients(a, _mont(chal2))
    t_cpu.write_commitments([oracle.variable_base_msm(q, eqs_host[i]) for i, q in enumerate(qs)])
    assert proof_gpu == bytes(t_cpu.stream)
    assert len(proof_gpu) == 3 * 64 + k * 4 * 32 + 3 * 32 + k * 64

    # ---- the verifier's replay from the bytes alone (classic.rs:242-262; kzg.rs:304-362 would follow with pairings)
    pos = 3 * 64
    v = OneShotTranscript(oracle)
    v.absorbed = b"".join(proof_gpu[i: i + 32][::-1] for i in range(0, pos, 32))  # read_commitment re-absorbs x, y little-endian
    yv = [v.squeeze_challenge() for _ in range(k)]
    claim = 0
    rv = []
    for _ in range(k):
        msg = [int.from_bytes(proof_gpu[pos + 32 * j: pos + 32 * j + 32], "big") for j in range(4)]
        pos += 4 * 32
        assert (msg[0] + msg[1]) % R == claim
        v.write_field_elements(msg)
        ch = v.squeeze_challenge()
        rv.append(ch)
        claim = sumcheck.interpolate_at(msg, ch)
    assert yv == y and rv == challenges
    wit = [int.from_bytes(proof_gpu[pos + 32 * j: pos + 32 * j + 32], "big") for j in range(3)]
    eq_at_r = 1
    for y_i, r_i in zip(yv, rv):
        eq_at_r = eq_at_r * ((y_i * r_i + (1 - y_i) * (1 - r_i)) % R) % R
    assert claim == _expr_int(terms_int, 0, [eq_at_r] + wit)
    for r in resident + [eq]:
        r.release()
    pp.release()
