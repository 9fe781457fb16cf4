"""Randomised stress of the GPU arithmetic and the MSM against Python integers / the oracle (run on a GPU box):
python tools/fuzz_gpu.py [seconds] [seed] [pcs].  The four phases take 0.4 / 0.6 / 0.3 / 0.2 of `seconds`.
Field level: products, fused two-product sums and squarings on structured limb patterns (all-ones / all-zero limbs,
values around the modulus) and random values.  MSM level: random sizes, both bucket layouts, duplicates, identities,
negated pairs, skewed scalars."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import plonkish_b200 as pk
from plonkish_b200 import _lib
from oracle import pyoracle as po, bigint_ref as br

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
only = sys.argv[3] if len(sys.argv) > 3 else ""   # "pcs": the Zeromorph / Gemini phase alone (the other phases get no time)
share = (lambda f: 0.0) if only == "pcs" else (lambda f: f)
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)

def limbs32(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint64).reshape(-1, 4).copy()
def ints(arr):
    return [int.from_bytes(row.tobytes(), "little") for row in arr]
def field_op(op, a, b=None):
    out = np.zeros_like(a)
    _lib.check(_lib.lib().plonkish_cuda_debug_field_op(0, op, a.ctypes.data, None if b is None else b.ctypes.data, out.ctypes.data, a.shape[0]), "debug_field_op")
    return out

def structured(mod, count):
    vals = []
    pats = [0, 0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, 1, 0xFFFFFFFE]
    while len(vals) < count:
        v = 0
        for i in range(8):
            choice = int(rng.integers(0, 8))
            limb = pats[choice] if choice < len(pats) else int(rng.integers(0, 1 << 32))
            v |= limb << (32 * i)
        vals.append(v % mod)
    return vals

t_end = time.time() + seconds * share(0.4)
rounds = 0
while time.time() < t_end:
    for mod, op_mul, op_sum, op_sqr in ((br.P, 0, 8, 10), (br.R, 5, 9, 11)):
        n = 20000
        x = structured(mod, n // 2) + [mod - 1 - int(v) for v in rng.integers(0, 1000, n // 4)] + [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(n // 4)]
        y = structured(mod, n // 2) + [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(n // 2)]
        rinv = pow(br.MONT, -1, mod)
        X, Y = limbs32(x), limbs32(y)
        assert ints(field_op(op_mul, X, Y)) == [a * b * rinv % mod for a, b in zip(x, y)], "mul"
        assert ints(field_op(op_sqr, X)) == [a * a * rinv % mod for a in x], "sqr"
        got = ints(field_op(op_sum, X, Y))
        if op_sum == 8:
            want = [(x[i] * y[i] + y[i] * x[(i + 1) % n]) * rinv % mod for i in range(n)]
        else:
            want = [(x[i] * x[i] + y[i] * y[(i + 1) % n]) * rinv % mod for i in range(n)]
        assert got == want, "mul_sum"
    rounds += 1
print(f"field fuzz: {rounds} rounds x 2 fields x 20000 elements x 3 ops ok", flush=True)

t_end = time.time() + seconds * share(0.6)
cases = 0
while time.time() < t_end:
    n = int(2 ** rng.uniform(0, 17))
    bases = po.known_dlog_bases(int(rng.integers(1, 1000)), int(rng.integers(1, 1000)), n)
    kind = int(rng.integers(0, 6))
    sc = pk.random_scalars(n, seed=int(rng.integers(0, 1 << 31)))
    if kind == 1:   # small integers / selectors
        vals = rng.integers(0, 3, n)
        sc = np.stack([po.from_canonical(1, po.int_to_limbs([0, 1, br.R - 1][int(v)]))[0] for v in vals]) if n <= 4096 else sc
    elif kind == 2:  # duplicates
        bases[:] = bases[rng.integers(0, max(1, n // 8), n)]
    elif kind == 3:  # identities mixed in
        bases[rng.random(n) < 0.3] = 0
    elif kind == 4 and n >= 2:  # P and -P with equal scalars
        half = n // 2
        neg = bases[:half].copy()
        for i in range(min(half, 64)):
            y = int.from_bytes(neg[i, 4:].tobytes(), "little")
            neg[i, 4:] = np.frombuffer(((br.P - y) % br.P).to_bytes(32, "little"), dtype=np.uint64) if y else neg[i, 4:]
        m = min(half, 64)
        bases[half:half + m] = neg[:m]
        sc[half:half + m] = sc[:m]
    want = po.variable_base_msm(sc, bases)
    assert pk.variable_base_msm(sc, bases).tobytes() == want.tobytes(), ("plain", n, kind)
    for mode in (pk.G1Bases.TABLE, pk.G1Bases.PLAIN):
        reg = pk.G1Bases(bases, mode=mode)
        assert pk.variable_base_msm(sc, reg).tobytes() == want.tobytes(), (mode, n, kind)
        use = int(rng.integers(1, n + 1))
        assert pk.variable_base_msm(sc[:use], reg).tobytes() == po.variable_base_msm(sc[:use], bases[:use]).tobytes(), ("prefix", mode, n, use, kind)
        reg.release()
    cases += 1
print(f"msm fuzz: {cases} random cases x (unregistered, table, plain, prefix) ok", flush=True)

# ---- the callers either side of the MSM: random open / merge / sum-check cases against the oracle
from plonkish_b200 import kzg
from plonkish_b200.sumcheck import SumCheckProver

t_end = time.time() + seconds * share(0.3)
cases = 0
one = po.from_canonical(1, po.int_to_limbs(1))[0]
while time.time() < t_end:
    k = int(rng.integers(1, 13))
    seed = int(rng.integers(0, 1 << 30))
    ss = po.random_scalars(k, seed)
    pp = kzg.setup(po.generator(), ss)
    eqs = [pp.eq(i).to_host() for i in range(k + 1)]
    count = int(rng.integers(1, 6))
    polys = [po.random_scalars(1 << k, seed + 1 + j) for j in range(count)]
    comms, res = kzg.batch_commit(pp, polys, keep=True)
    for p, c in zip(polys, comms):
        assert c.tobytes() == po.variable_base_msm(p, eqs[k]).tobytes(), ("commit", k)
    coeffs = po.random_scalars(count, seed + 50)
    merged = kzg.linear_combination(res, coeffs)
    merged_h = po.fr_linear_combination(polys, coeffs)
    assert merged.to_host().tobytes() == merged_h.tobytes(), ("merge", k, count)
    point = po.random_scalars(k, seed + 60)
    q_comms, value = kzg.open_resident(pp, merged, point)
    qs, want = po.quotients(merged_h, point)
    assert value.tobytes() == want.tobytes() and all(a.tobytes() == po.variable_base_msm(q, eqs[i]).tobytes() for i, (a, q) in enumerate(zip(q_comms, qs))), ("open", k)
    # a random expression over the same tables
    nterms = int(rng.integers(1, 7))
    terms = []
    for t in range(nterms):
        nf = int(rng.integers(0 if t else 1, 5))
        terms.append((one if rng.random() < 0.4 else po.random_scalars(1, seed + 70 + t)[0], [int(i) for i in rng.integers(0, count, nf)]))
    common = int(rng.integers(-1, count))
    prover = SumCheckProver(res, terms, common)
    cur = polys
    for rnd in range(k):
        assert prover.round_evals().tobytes() == po.sumcheck_round(cur, terms, common).tobytes(), ("sumcheck", k, rnd)
        ch = po.random_scalars(1, seed + 90 + rnd)[0]
        prover.fix_var(ch)
        cur = [po.fix_var(p, ch) for p in cur]
    assert prover.final_evals().tobytes() == np.stack([p[0] for p in cur]).tobytes()
    prover.free()
    for r in res + [merged]:
        r.release()
    pp.release()
    cases += 1
print(f"caller fuzz: {cases} random setup / commit / merge / open / sum-check cases ok", flush=True)

# ---- Zeromorph / Gemini over a univariate SRS: random sizes, setup lengths and prefix slices, each proof against the scheme's
#      verifier equation in G1 (tests/zeromorph_ref.py, tests/gemini_ref.py) for the oracle's evaluation of the polynomial
sys.path.insert(0, "tests")
import gemini_ref as gr
import zeromorph_ref as zr
from hyperplonk_ref import ProofReader
from plonkish_b200 import gemini, zeromorph
from plonkish_b200.sumcheck import _to_int, _to_mont
from plonkish_b200.transcript import Keccak256Transcript

t_end = time.time() + seconds * (1.0 if only == "pcs" else 0.2)
cases = 0
while time.time() < t_end:
    k = int(rng.integers(2, 13))
    n = 1 << k
    extra = int(rng.integers(0, 40)) if rng.random() < 0.3 else 0
    s = int(rng.integers(2, 1 << 62)) * int(rng.integers(2, 1 << 62)) % br.R
    full = kzg.univariate_setup(po.generator(), _to_mont(s), n + extra)
    zpp = zeromorph.trim(full, n, prefix_tables=bool(rng.random() < 0.5))
    poly_h = po.random_scalars(n, int(rng.integers(0, 1 << 30)))
    poly = pk.ResidentScalars(poly_h)
    for scheme in ("zeromorph", "gemini"):
        t = Keccak256Transcript()
        gpp = gemini.GeminiKzgProverParam(zpp.commit_pp)
        comm = zeromorph.commit(zpp, poly) if scheme == "zeromorph" else gemini.commit(gpp, poly)
        t.write_commitment(comm)
        point = t.squeeze_challenges(k)
        value = _to_int(po.evaluate_multilinear(poly_h, zr.mont_rows(point)))
        t.write_field_element(value)
        if scheme == "zeromorph":
            assert zeromorph.open(zpp, poly, point, value, t) == value, ("zeromorph remainder", k)
        else:
            gemini.open(gpp, poly, point, t)
        proof = t.into_proof()
        reader = ProofReader(po.keccak256, proof)
        c = reader.read_commitment()
        assert reader.squeeze_challenges(k) == point and reader.read_field_element() == value
        if scheme == "zeromorph":
            zr.verify_reader_in_g1(reader, c, point, value, s, extra)
        else:
            gr.verify_reader_in_g1(reader, c, point, value, s)
        assert reader.pos == len(proof), (scheme, k)
    poly.release()
    if extra:
        zpp.release()
    else:
        zeromorph.release_prefix_tables(zpp.commit_pp)
    full.release()
    cases += 1
print(f"pcs fuzz: {cases} random Zeromorph + Gemini openings (setup longer than the polynomial in a third, prefix slices in half) verified", flush=True)
