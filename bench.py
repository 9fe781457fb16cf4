#!/usr/bin/env python
"""bench.py — BN254 G1 variable-base MSM throughput on B200 (BASELINE.json metric).

A "step" is one MSM over synthetic inputs: uniform random Fr scalars (Montgomery
limbs) and known-discrete-log bases B_i = (a + i*d)G generated on the GPU.
  N = 1   one MSM of 2^LOG_N points (default 2^24, the size the metric is quoted at).
  N > 1   one MSM of N * 2^LOG_N points, point-sharded 2^LOG_N per rank (weak scaling):
          every rank runs the whole pipeline on its slice, the 128-byte projective
          partials are all-gathered with NCCL and folded on every rank
          (reference: msm.rs:101-114 chunk-per-thread + fold).
`value`  device-resident inputs (scalars and bases already in HBM), CUDA-event timed.
`e2e`    N = 1: the C-ABI host call plonkish_cuda_msm_bn254_g1 with the step's scalars in
         pinned host memory and the bases resident (registered once, like the SRS of a
         ProverParam); N > 1: pinned host scalars -> H2D -> sharded MSM -> result to host.
`--impl reference`  the CPU restatement of the reference algorithm (oracle/, C port of
         msm.rs:84-181, one pthread per chunk like rayon) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BN254 G1 variable_base_msm throughput"
UNIT = "Mpoints/s"
IMAD_PER_MODMUL = 136          # 8x8 product + 8x8 reduction + 8 quotient words
MODMUL_PER_MIXED_ADD = 10      # XYZZ madd-2008-s: 8M + 2S


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24, help="log2 of the points per GPU")
    ap.add_argument("--cpu-log-n", type=int, default=21, help="log2 of the cpu_baseline sample")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--plain-bases", action="store_true", help="do not expand the resident bases into the table of window multiples")
    ap.add_argument("--prove-k", type=int, default=24, help="k of the HyperPlonk prove MSM-sequence surrogate (0 = skip)")
    return ap.parse_args()


# ------------------------------------------------------------------ clock sampling
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append((time.time(), parts))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [p for (t, p) in self.rows if t0 <= t <= t1] or [p for (_, p) in self.rows]
        clocks, reasons, smax, power = [], set(), None, []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for p in rows:
            try:
                clocks.append(float(p[1]))
                smax = float(p[2])
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, val in zip(names, p[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(clocks) if clocks else None,
            "sm_max_mhz": smax,
            "power_w_max": max(power) if power else None,
            "samples": len(clocks),
            "reasons": sorted(reasons),
        }


# -------------------------------------------------------------------- reference arm
def run_reference(args) -> None:
    """CPU restatement of the reference's variable_base_msm on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po

    po.build()
    cores = po.host_threads()
    log_n = min(args.log_n, args.cpu_log_n)
    n = 1 << log_n
    scalars = po.random_scalars(n, seed=1234)
    bases = po.known_dlog_bases(3, 5, n)
    want = po.known_dlog_answer(3, 5, scalars)
    for _ in range(min(args.warmup, 1)):
        po.variable_base_msm(scalars, bases, cores)
    times = []
    for _ in range(args.steps):
        t = time.perf_counter()
        got = po.variable_base_msm(scalars, bases, cores)
        times.append(time.perf_counter() - t)
        assert (got == want).all(), "oracle result differs from the known-dlog answer"
    sec = sum(times) / len(times)
    value = n / sec / 1e6
    sample = f"2^{log_n} points per step (bounded sample of the 2^{args.log_n}-point workload), {cores} pthreads, C port of msm.rs:84-181"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
        "config": {"workload": f"BN254 G1 variable_base_msm, uniform random scalars, known-dlog bases, 2^{log_n} points/step on host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------- HyperPlonk::prove MSM-sequence surrogate
FQ_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


def g1_generator(np):
    """(1, 2) as Montgomery limbs, the layout of bn256::G1Affine::generator()."""
    b = b"".join((v * (1 << 256) % FQ_MODULUS).to_bytes(32, "little") for v in (1, 2))
    return np.frombuffer(b, dtype=np.uint64).copy()


def prove_msm_sequence(pk, torch, np, k: int, dev, cpu: bool):
    """The MSM calls HyperPlonk::prove makes for vanilla_plonk at 2^k rows:
    4 commits of 2^k points (3 witness polys + 1 permutation z-poly, backend/hyperplonk.rs:201,
    251-252) and the k quotient commitments of MultilinearKzg::open with 2^(k-1), ..., 2, 1 points
    (kzg.rs:291-293) against the resident eqs[i] slices of an SRS built on the device
    (MultilinearKzg::setup, kzg.rs:167-212).  Two forms:
      gpu_ms           every call takes host scalars (the drop-in for msm.rs alone; the opened polynomial's
                       quotients are stand-ins of the right sizes);
      gpu_resident_ms  the committed polynomials stay in HBM (batch_commit keep), g_prime is merged there
                       (multilinear.rs:203-213) and open() computes its quotients there (multilinear.rs:72-107):
                       the same 4 + k commitments plus the real quotient arithmetic, no scalar uploaded twice."""
    from plonkish_b200 import kzg

    n = 1 << k
    ss = pk.random_scalars(k, seed=77)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pp = kzg.setup(g1_generator(np), ss)
    setup_s = time.perf_counter() - t0
    regs = pp.eqs
    polys = []
    for j in range(4):
        t = torch.empty((n, 4), dtype=torch.int64).pin_memory()
        h = t.numpy().view(np.uint64)
        h[:] = pk.random_scalars(n, seed=4242 + j)
        polys.append(h)
    host = polys[0]
    coeffs = pk.random_scalars(4, seed=78)
    point = pk.random_scalars(k, seed=79)

    def run():
        # batch_commit of the 3 witness polynomials (hyperplonk.rs:201), then the z-poly commit (:251)
        outs = list(pk.variable_base_msm_batch(polys[:3], regs[k])) + [pk.variable_base_msm(polys[3], regs[k])]
        # open: the k quotient commitments in one call (kzg.rs:291-293), small ones concurrently
        outs += list(pk.variable_base_msm_many([host[: 1 << i] for i in reversed(range(k))], [regs[i] for i in reversed(range(k))]))
        return outs

    def run_resident():
        c3, r3 = kzg.batch_commit(pp, polys[:3], keep=True)
        c1, r1 = kzg.batch_commit(pp, polys[3:], keep=True)
        g_prime = kzg.linear_combination(r3 + r1, coeffs)
        q_comms, value = kzg.open_resident(pp, g_prime, point)
        for r in r3 + r1 + [g_prime]:
            r.release()
        return list(c3) + list(c1), q_comms, value

    run()
    t0 = time.perf_counter()
    outs = run()
    gpu_ms = (time.perf_counter() - t0) * 1e3
    run_resident()
    t0 = time.perf_counter()
    comms, q_comms, value = run_resident()
    resident_ms = (time.perf_counter() - t0) * 1e3
    res = {"k": k, "msm_calls": 4 + k, "points": 4 * n + n - 1, "gpu_ms": gpu_ms, "gpu_resident_ms": resident_ms,
           "srs_setup_on_device_s": setup_s, "srs_points": 2 * n - 1}
    if cpu:
        from oracle import pyoracle as po

        cores = po.host_threads()
        eqs_h = [r.to_host() for r in regs]
        t0 = time.perf_counter()
        ref = [po.variable_base_msm(p, eqs_h[k], cores) for p in polys]
        t_commit = time.perf_counter() - t0
        ref_host_open = [po.variable_base_msm(host[: 1 << i], eqs_h[i], cores) for i in reversed(range(k))]
        res["cpu_ms"] = (time.perf_counter() - t0) * 1e3
        res["cpu_cores"] = cores
        res["bit_exact_vs_cpu"] = bool(all((a == b).all() for a, b in zip(outs, ref + ref_host_open)))
        # the resident form on the CPU: merge, quotients, their commitments (single-threaded field part)
        t0 = time.perf_counter()
        g_prime_h = po.fr_linear_combination(polys, coeffs)
        qs, want_value = po.quotients(g_prime_h, point)
        ref_q = [po.variable_base_msm(qs[i], eqs_h[i], cores) for i in range(k)]
        res["cpu_resident_ms"] = (t_commit + time.perf_counter() - t0) * 1e3
        res["resident_bit_exact_vs_cpu"] = bool(all((a == b).all() for a, b in zip(comms, ref)) and all((a == b).all() for a, b in zip(q_comms, ref_q))
                                                and (value == want_value).all())
    pp.release()
    return res


def srs_setup_bench(pk, torch, np, k: int, cpu: bool):
    """fixed_base_msm + batch_normalize (msm.rs:16-31, 50-81; kzg.rs:195-208) on the GPU: 2^22 scalars host to host, and the
    CPU port on a bounded sample with the window the reference would pick for a 2^k setup."""
    n = 1 << 22
    sc = pk.random_scalars(n, seed=91)
    g = g1_generator(np)
    pk.fixed_base_msm(g, sc[: 1 << 16])
    t0 = time.perf_counter()
    got = pk.fixed_base_msm(g, sc)
    sec = time.perf_counter() - t0
    res = {"what": "fixed_base_msm + batch_normalize of 2^22 scalars, host scalars in, affine points out (table build included)",
           "gpu_mpoints_per_s": n / sec / 1e6, "gpu_ms": sec * 1e3}
    if cpu:
        from oracle import pyoracle as po

        m = 1 << 17
        window = po.window_size((2 << k) - 2)
        cores = po.host_threads()
        t0 = time.perf_counter()
        want = po.fixed_base_msm(g, sc[:m], window=window, num_threads=cores)
        sec_c = time.perf_counter() - t0
        res.update({"cpu_mpoints_per_s": m / sec_c / 1e6, "cpu_cores": cores, "cpu_sample": f"2^17 scalars, window {window} (table build included)",
                    "bit_exact_vs_cpu": bool((got[:m] == want).all())})
    return res


def sum_check_bench(pk, torch, np, k: int, cpu: bool):
    """The zero check of HyperPlonk::prove for vanilla_plonk (backend/hyperplonk.rs:262-277) as ClassicSumCheck runs it
    (piop/sum_check/classic.rs:208-240): 9 tables (eq, 5 selectors, 3 witness columns) of 2^k evaluations, degree 4,
    k rounds of round-polynomial evaluations + table folds on the GPU; the challenges are stand-ins for the transcript's.
    CPU: the oracle's restatement of the same rounds, one thread, on tables of 2^18 evaluations."""
    from plonkish_b200 import sumcheck

    n = 1 << k
    tables = [pk.ResidentScalars(pk.random_scalars(n, seed=300 + i)) for i in range(9)]
    one = sumcheck._to_mont(1)
    terms = [(one, [1, 6]), (one, [2, 7]), (one, [3, 6, 7]), (one, [4, 8]), (one, [5])]
    chal = pk.random_scalars(k, seed=399)

    def run():
        prover = sumcheck.SumCheckProver(tables, terms, common=0)
        msgs = []
        for rnd in range(k):
            msgs.append(prover.round_evals())
            prover.fix_var(chal[rnd])
        finals = prover.final_evals()
        prover.free()
        return msgs, finals

    run()
    t0 = time.perf_counter()
    msgs, finals = run()
    gpu_ms = (time.perf_counter() - t0) * 1e3
    res = {"what": "zero check of vanilla_plonk as ClassicSumCheck<EvaluationsProver> runs it: 9 resident tables of 2^k evaluations, degree 4, "
                   "k rounds (round-polynomial evaluations at X = 1..4 + fold of every table), stand-in challenges",
           "k": k, "gpu_ms": gpu_ms, "gpu_mpairs_per_s_first_round_equiv": (n - 1) / gpu_ms / 1e3}
    if cpu:
        from oracle import pyoracle as po

        kc = min(18, k)
        host = [t.to_host(0, 1 << kc) for t in tables]
        sub = [pk.ResidentScalars(h) for h in host]
        prover = sumcheck.SumCheckProver(sub, terms, common=0)
        ok = True
        cur = host
        t0 = time.perf_counter()
        ref = []
        for rnd in range(kc):
            ref.append(po.sumcheck_round(cur, terms, 0))
            cur = [po.fix_var(p, chal[rnd]) for p in cur]
        cpu_ms = (time.perf_counter() - t0) * 1e3
        for rnd in range(kc):
            ok = ok and prover.round_evals().tobytes() == ref[rnd].tobytes()
            prover.fix_var(chal[rnd])
        ok = ok and prover.final_evals().tobytes() == np.stack([p[0] for p in cur]).tobytes()
        prover.free()
        for r in sub:
            r.release()
        res.update({"cpu_ms": cpu_ms, "cpu_k": kc, "cpu_cores": 1, "cpu_mpairs_per_s": ((1 << kc) - 1) / cpu_ms / 1e3, "bit_exact_vs_cpu": bool(ok)})
    for t in tables:
        t.release()
    return res


def prove_pipeline_bench(pk, torch, np, k: int, cpu: bool):
    """The GPU-side compute of a HyperPlonk proof for vanilla_plonk end to end, driven by the reference's Keccak256
    transcript (util/transcript.rs:100-235): commit the three witness polynomials (backend/hyperplonk.rs:201, kept
    resident), zero check of the gate over eq(x, y) and five resident selectors (hyperplonk.rs:262-277 through
    piop/sum_check/classic.rs:208-240), g_prime merge of the witness polynomials (pcs/multilinear.rs:203-213) and its
    KZG opening at the sum-check point (kzg.rs:276-302).  The permutation / lookup arguments and witness generation are
    not part of it.  With cpu=True the same proof is rebuilt through the oracle: the bytes must be identical."""
    from plonkish_b200 import kzg, sumcheck
    from plonkish_b200.transcript import FR_MODULUS, Keccak256Transcript

    n = 1 << k
    pp = kzg.setup(g1_generator(np), pk.random_scalars(k, seed=501))
    witness = []
    for j in range(3):
        t = torch.empty((n, 4), dtype=torch.int64).pin_memory()
        h = t.numpy().view(np.uint64)
        h[:] = pk.random_scalars(n, seed=510 + j)
        witness.append(h)
    selectors = [pk.ResidentScalars(pk.random_scalars(n, seed=520 + j)) for j in range(5)]  # q_l, q_r, q_m, q_o, q_c: preprocessed
    one = sumcheck._to_mont(1)
    # tables: 0 eq, 1..5 selectors, 6..8 witness
    terms = [(one, [1, 6]), (one, [2, 7]), (one, [3, 6, 7]), (one, [4, 8]), (one, [5])]

    def run():
        t = Keccak256Transcript()
        comms, resident = kzg.batch_commit(pp, witness, keep=True)
        t.write_commitments(comms)
        y = t.squeeze_challenges(k)
        eq = pk.eq_table(np.stack([sumcheck._to_mont(v) for v in y]))
        # the claimed sum of a random (unsatisfied) instance is whatever the first message implies: run with 0, as the
        # reference's prover would with a satisfying witness; the arithmetic per round is the same
        challenges, evals = sumcheck.prove_to_transcript([eq] + selectors + resident, terms, 0, t, common=0)
        t.write_field_elements(evals[6:])
        coeffs = t.squeeze_challenges(3)
        g_prime = kzg.linear_combination(resident, np.stack([sumcheck._to_mont(c) for c in coeffs]))
        point = np.stack([sumcheck._to_mont(c) for c in challenges])
        kzg.open_to_transcript(pp, g_prime, point, t)
        for r in resident + [eq, g_prime]:
            r.release()
        return t.into_proof()

    run()
    t0 = time.perf_counter()
    proof = run()
    gpu_ms = (time.perf_counter() - t0) * 1e3
    res = {"what": "commit 3 witness polynomials -> zero check (9 tables, degree 4) -> g_prime merge -> KZG open, Keccak256 transcript on the host, "
                   "all polynomial data resident in HBM after one upload; permutation / lookup arguments and witness generation not included",
           "k": k, "gpu_ms": gpu_ms, "proof_bytes": len(proof)}
    if cpu:
        from oracle import pyoracle as po
        from plonkish_b200.sumcheck import interpolate_at

        cores = po.host_threads()
        eqs_h = [e.to_host() for e in pp.eqs]
        sel_h = [s_.to_host() for s_ in selectors]
        t0 = time.perf_counter()
        t = Keccak256Transcript()
        t.write_commitments([po.variable_base_msm(w, eqs_h[k], cores) for w in witness])
        y = t.squeeze_challenges(k)
        cur = [po.kzg_eq_scalars(np.stack([sumcheck._to_mont(v) for v in y]))[k]] + sel_h + [np.array(w) for w in witness]
        claim, challenges = 0, []
        for _ in range(k):
            tail = [sumcheck._to_int(r) for r in po.sumcheck_round(cur, terms, 0)]
            msg = [(claim - tail[0]) % FR_MODULUS] + tail
            t.write_field_elements(msg)
            ch = t.squeeze_challenge()
            challenges.append(ch)
            claim = interpolate_at(msg, ch)
            cur = [po.fix_var(p, sumcheck._to_mont(ch)) for p in cur]
        t.write_field_elements([sumcheck._to_int(p[0]) for p in cur[6:]])
        coeffs = t.squeeze_challenges(3)
        g_prime_h = po.fr_linear_combination(witness, np.stack([sumcheck._to_mont(c) for c in coeffs]))
        qs, _ = po.quotients(g_prime_h, np.stack([sumcheck._to_mont(c) for c in challenges]))
        t.write_commitments([po.variable_base_msm(q, eqs_h[i], cores) for i, q in enumerate(qs)])
        res.update({"cpu_ms": (time.perf_counter() - t0) * 1e3, "cpu_cores": cores, "proof_bytes_identical_to_cpu": bool(t.into_proof() == proof)})
    for s_ in selectors:
        s_.release()
    pp.release()
    return res


def univariate_sequence(pk, torch, np, k: int, dev):
    """BASELINE.json config 4 restated synthetically: UnivariateKzg commit = one MSM over
    the SRS prefix (pcs/univariate/kzg.rs:24-30, witness-like 68-bit limb values with zero padding) and
    batch_open = two MSMs of ~2^k uniform coefficients (univariate/kzg.rs:330,353), host scalars,
    resident powers_of_s_g1."""
    n = 1 << k
    d_bases = pk.synth_bases_device(n, 11, 13, device=dev)
    torch.cuda.synchronize()
    reg = pk.G1Bases(d_bases)
    rng = np.random.default_rng(22)
    limbs = np.zeros((n, 4), dtype=np.uint64)
    live = n - n // 8                       # last eighth zero padding
    limbs[:live, 0] = rng.integers(0, 1 << 63, size=live, dtype=np.uint64)
    limbs[:live, 1] = rng.integers(0, 1 << 4, size=live, dtype=np.uint64)   # 68-bit canonical values ...
    # ... as Montgomery representations they are full-width; the MSM sees uniform digits, so the 68-bit
    # shape is emulated directly in Montgomery form: values whose Montgomery limbs are small.
    uniform = pk.random_scalars(n, seed=2222)
    host = [torch.from_numpy(a.view(np.int64)).pin_memory().numpy().view(np.uint64) for a in (limbs, uniform, uniform)]

    def run():
        return [pk.variable_base_msm(h, reg) for h in host]

    run()
    t0 = time.perf_counter()
    run()
    ms = (time.perf_counter() - t0) * 1e3
    reg.release()
    return {"k": k, "msm_calls": 3, "points": 3 * n, "gpu_ms": ms,
            "what": "commit (small-limb scalars, 1/8 zero padding) + batch_open (2 uniform MSMs) of 2^k points each, host scalars"}


# ------------------------------------------------------------------------- our arm
def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import plonkish_b200 as pk
    from plonkish_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    n = 1 << args.log_n
    total_n = n * world
    a, d = 3, 5
    first = rank * n
    # synthetic inputs (seeded): this rank's slice of the (world * n)-point MSM
    scalars_host_t = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    scalars_np = scalars_host_t.numpy().view(np.uint64)
    scalars_np[:] = pk.random_scalars(n, seed=1000 + rank)
    d_scalars = scalars_host_t.to(dev)
    d_bases = pk.synth_bases_device(n, a, d, device=dev, first=first)
    torch.cuda.synchronize()
    # The bases are the static SRS of a ProverParam: made resident once, outside the timed
    # region (plonkish_cuda_bases_register_device).  --plain-bases keeps the plain affine
    # array; the default expands it into the table of window multiples.
    t_reg = time.perf_counter()
    reg = pk.G1Bases(d_bases, mode=pk.G1Bases.PLAIN if args.plain_bases else 0)
    torch.cuda.synchronize()
    register_s = time.perf_counter() - t_reg
    step_bases = d_bases if (args.plain_bases and args.window_bits) else reg

    def step_device():
        if distributed:
            return pk.variable_base_msm_sharded(d_scalars, step_bases, window_bits=args.window_bits)
        return pk.variable_base_msm_device(d_scalars, step_bases, window_bits=args.window_bits)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step_device()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = pk.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]  # per-step boundaries for min / median
    e0.record()
    for i in range(args.steps):
        out = step_device()
        marks[i].record()
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = pk.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    step_ms = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(args.steps)]
    if distributed:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_per_step = ms_total / args.steps
    value = total_n / (ms_per_step * 1e-3) / 1e6
    result_dev = out.cpu().numpy().view(np.uint64)

    # ---- e2e: host buffers through the public API, copies inside the timed region
    if not distributed:
        def step_e2e():
            return pk.variable_base_msm(scalars_np, reg)
    else:
        def step_e2e():
            return pk.variable_base_msm_sharded_host(scalars_np, reg).cpu().numpy().view(np.uint64)

    for _ in range(2):
        e2e_out = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_out = step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if distributed:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = total_n / (e2e_ms * 1e-3) / 1e6
    assert (np.asarray(e2e_out).view(np.uint64) == result_dev).all(), "e2e and device-resident results differ"

    # ---- roofline of the dominant kernel (K3 accumulate), timed live with CUDA events
    plan = pk.msm_plan(n, args.window_bits, local_rank, bases=None if step_bases is d_bases else reg)
    stage_runs = [pk.profile_stages_device(d_scalars, step_bases, window_bits=args.window_bits) for _ in range(3)]
    stages = {k: statistics.mean(r[k] for r in stage_runs) for k in stage_runs[0]}
    pipe = pk.bench_integer_pipe(local_rank)
    madd_streams = pk.bench_madd(local_rank)
    # the algorithmic figure is fixed at 16 windows x 10 modmul x 136 IMAD
    # = 21 760 IMAD per point, independent of the window width the plan actually uses.
    imad_per_launch = float(n) * 16 * MODMUL_PER_MIXED_ADD * IMAD_PER_MODMUL
    # executed per mixed addition: 6 products (136 each), 2 squarings with the symmetric partial products taken once
    # (36 + 64 + 8 = 108) and the y-coordinate's fused two-product reduction (200)
    executed_imad = float(n) * plan["windows"] * (6 * IMAD_PER_MODMUL + 2 * 108 + 200)
    achieved = imad_per_launch / (stages["accumulate"] * 1e-3) / 1e12
    peak = max(pipe["imad_wide_per_s"], pipe["imad_wide_chain_per_s"], pipe["fq_mul_per_s"] * IMAD_PER_MODMUL) / 1e12
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_file):
        try:
            hbm_peak, hbm_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    entries = float(n) * plan["windows"]
    if plan["idx_bits"]:  # plain bases: u16 digit in, u32 entry out; then u32 in, u32 out
        sort_bytes = entries * (2 + 4) + entries * (4 + 4)
    else:                 # table layout: staged two-level partition
        sort_bytes = entries * (4 + 6) + entries * (2 + 6 + 4)  # level 1: digit in, value + key out; level 2: key (hist), key + value in, entry out
    sort_ms = stages["bin_scatter"] + stages["bin_sort"]
    roofline = {
        "kernel": "k_accumulate (XYZZ mixed additions, 254-bit Montgomery, IMAD.WIDE carry chains)",
        "bound": "imad", "achieved": achieved, "peak": peak, "unit": "TIMAD/s (32x32->64 multiply-adds)",
        "frac": achieved / peak,
        "frac_note": "above 1 is expected with the table layout: the algorithmic figure is fixed at 16 windows per point "
                     "(SURVEY.md 8d: 16 windows x 10 modmul x 136) while the kernel executes plan['windows'] windows of 6 x 136 + 2 x 108 + 200 "
                     "multiply instructions (symmetric squarings; the y-coordinate's two products share one reduction); executed_frac is the executed-work fraction",
        "executed_frac": (executed_imad / (stages["accumulate"] * 1e-3) / 1e12) / peak,
        "peak_source": "measured in this run by plonkish_cuda_bench_integer_pipe: max(independent mad.wide.u32 stream, "
                       "IMAD.WIDE.U32.X carry-chain stream, library fq_mul stream x 136); MEASURED_PEAKS.json has no integer-pipe figure",
        "algorithmic_imad_per_launch": imad_per_launch,
        "executed_imad_per_launch": executed_imad,
        "executed_timad_per_s": executed_imad / (stages["accumulate"] * 1e-3) / 1e12,
        "kernel_ms": stages["accumulate"],
        # what the same mixed addition reaches with everything in registers (no gathers, no bucket logic):
        # the fraction of THAT ceiling the kernel runs at, in executed products
        "madd_stream_ceiling_products_per_s": madd_streams["madd_1acc_128regs"],
        "frac_of_madd_stream": (float(n) * plan["windows"] * MODMUL_PER_MIXED_ADD / (stages["accumulate"] * 1e-3)) / madd_streams["madd_1acc_128regs"],
        # dram__bytes_read.sum + dram__bytes_write.sum of one 2^24-point launch in the committed ncu --set full capture
        # (profiles/r01_final_kernels_ncu_raw.csv: 27.25 GB + 0.39 GB at 12 windows), scaled to this launch's entry count
        "traffic": 27.64e9 * (float(n) * plan["windows"]) / (16777216.0 * 12),
        "traffic_source": "ncu capture under profiles/ (not measured in this run); algorithmic gather = entries x 68 B",
        "algorithmic_bytes": float(n) * plan["windows"] * 68,
    }
    roofline_sort = {
        "kernel": "the two sort levels (k_scatter_staged_b, k_bucket_hist_b, k_bucket_scatter_staged_b; plain bases: k_scatter_bins, k_sort_bins)", "bound": "hbm", "achieved": sort_bytes / (sort_ms * 1e-3) / 1e9,
        "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
        "frac": sort_bytes / (sort_ms * 1e-3) / 1e9 / hbm_peak, "kernel_ms": sort_ms,
        # table layout, 2^24, c = 22: k_scatter_staged_b 0.81 + 1.16 GB, k_bucket_hist_b 0.41 GB, k_bucket_scatter_staged_b 1.36 + 0.90 GB (same capture)
        "traffic": (5.51e9 * entries / (16777216.0 * 13)) if not plan["idx_bits"] else None,
        "algorithmic_bytes": sort_bytes,
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "ms_step_min": min(step_ms), "ms_step_median": statistics.median(step_ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
        "config": {
            "workload": f"one BN254 G1 variable_base_msm of {world} x 2^{args.log_n} points (2^{args.log_n} per GPU), "
                        "uniform random Fr scalars, known-dlog bases (a+i*d)G, bases resident",
            "points_per_gpu": n, "window_bits": plan["window_bits"], "windows": plan["windows"],
            "bases": "plain affine array (one bucket set per window)" if step_bases is d_bases or args.plain_bases else
                     f"resident table of window multiples, {plan['windows']} x 64 B per point, built once in {register_s:.2f} s (untimed)",
            "parallelism": f"point-sharded x{world}" + (", NCCL all_gather of 128-byte partials" if distributed else ""),
            "l2": "inputs (scalars 32 B + bases 64 B per point) exceed the 126 MB L2 at this size; no explicit flush",
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": 64 * world,
                "path": "plonkish_cuda_msm_bn254_g1 (C ABI, pinned host scalars, registered bases)" if not distributed
                        else "per rank: plonkish_cuda_msm_bn254_g1_host_partial (C ABI, pinned host scalars, registered bases) -> NCCL all_gather of partials -> fold -> host"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_sort": roofline_sort,
        "stages_ms": stages,
        "integer_pipe": pipe,
        "madd_streams": madd_streams,
    }

    if rank == 0 and not args.no_cpu_baseline and not distributed:
        from oracle import pyoracle as po

        po.build()
        cores = po.host_threads()
        log_c = min(args.log_n, args.cpu_log_n)
        m = 1 << log_c
        sc = scalars_np[:m].copy()
        bs = d_bases[:m].cpu().numpy().view(np.uint64)
        t0 = time.perf_counter()
        want = po.variable_base_msm(sc, bs, cores)
        sec = time.perf_counter() - t0
        got = pk.variable_base_msm(sc, bs)
        assert (got == want).all(), "GPU result differs from the CPU oracle on the cpu_baseline sample"
        line["cpu_baseline"] = {
            "value": m / sec / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first 2^{log_c} points of the step's inputs, one pass, C port of msm.rs:84-181 with {cores} pthreads; "
                      "GPU result on the same sample checked bit-exact",
        }
    if rank == 0 and not distributed and args.prove_k:
        del d_scalars, d_bases
        reg.release()
        torch.cuda.empty_cache()
        seq = {"what": "MSM calls of HyperPlonk::prove for vanilla_plonk (4 x 2^k + 2^(k-1) + ... + 1 points) against an SRS built on the device; "
                       "gpu_ms: host scalars per call; gpu_resident_ms: polynomials kept in HBM, g_prime merge and quotients on the GPU; "
                       "sum-check and the other field-only prover work stay in the Rust caller and are not included"}
        seq["k%d" % args.prove_k] = prove_msm_sequence(pk, torch, np, args.prove_k, dev, cpu=False)
        if not args.no_cpu_baseline:
            seq["k20"] = prove_msm_sequence(pk, torch, np, min(20, args.prove_k), dev, cpu=True)
        line["hyperplonk_prove_msm"] = seq
        line["univariate_kzg_k22"] = univariate_sequence(pk, torch, np, 22, dev)
        line["srs_fixed_base_msm"] = srs_setup_bench(pk, torch, np, args.prove_k, cpu=not args.no_cpu_baseline)
        line["sum_check_zero_check"] = sum_check_bench(pk, torch, np, args.prove_k, cpu=not args.no_cpu_baseline)
        pipe = {"k%d" % args.prove_k: prove_pipeline_bench(pk, torch, np, args.prove_k, cpu=False)}
        if not args.no_cpu_baseline:
            pipe["k18"] = prove_pipeline_bench(pk, torch, np, min(18, args.prove_k), cpu=True)
        line["hyperplonk_prove_pipeline"] = pipe
    if rank == 0:
        emit(line)
    if distributed:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version banner) must not pollute the one-JSON-line contract:
    everything written to fd 1 during the run goes to stderr; the result line goes to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    args = parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
