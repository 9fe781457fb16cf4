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
         PAGEABLE host memory (what a Rust Vec<Fr> is) and the bases resident (registered
         once, like the SRS of a ProverParam); `e2e_pinned` is the same from pinned memory.
         N > 1: host scalars -> H2D -> sharded MSM -> NCCL gather -> result to host.
`--impl reference`  the CPU restatement of the reference algorithm (oracle/, C port of
         msm.rs:84-181, one pthread per chunk like rayon) on the same 2^LOG_N points per step.
Every timed leg is checked against an independent answer (known discrete log of the synthetic
bases, the KZG trapdoor of the synthetic SRS, or the CPU port) and carries "parity_checked".
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
    ap.add_argument("--cpu-log-n", type=int, default=24, help="log2 of the cpu_baseline sample (capped at --log-n)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="the reference arm stops adding steps beyond this many seconds")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--plain-bases", action="store_true", help="do not expand the resident bases into the table of window multiples")
    ap.add_argument("--prove-k", type=int, default=24, help="k of the HyperPlonk prove legs (0 = skip)")
    ap.add_argument("--reps", type=int, default=3, help="repetitions of every auxiliary leg (min / median reported)")
    ap.add_argument("--no-skew", action="store_true")
    ap.add_argument("--no-single-process", action="store_true", help="N > 1: skip rank 0's single-process multi-GPU leg")
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
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
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
    """CPU restatement of the reference's variable_base_msm on this box's host cores, on the same 2^log_n points per
    step as the GPU arm (rank 0 only under torchrun)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po

    po.build()
    cores = po.host_threads()
    log_n = args.log_n
    n = 1 << log_n
    scalars = po.random_scalars(n, seed=1000)
    bases = po.known_dlog_bases(3, 5, n)
    want = po.known_dlog_answer(3, 5, scalars)
    t_all = time.perf_counter()
    warm = 0
    for _ in range(min(args.warmup, 1)):
        t = time.perf_counter()
        po.variable_base_msm(scalars, bases, cores)
        warm += 1
        if time.perf_counter() - t > 20.0:
            break
    times = []
    for i in range(args.steps):
        t = time.perf_counter()
        got = po.variable_base_msm(scalars, bases, cores)
        times.append(time.perf_counter() - t)
        assert (got == want).all(), "oracle result differs from the known-dlog answer"
        # the whole run has to end within minutes: stop early when the next step would not fit
        if i + 1 < args.steps and (time.perf_counter() - t_all) + times[-1] > args.cpu_budget_s:
            break
    sec = sum(times) / len(times)
    value = n / sec / 1e6
    sample = (f"2^{log_n} points per step (the GPU arm's workload), {len(times)} timed steps of {args.steps} requested "
              f"(budget {args.cpu_budget_s:.0f} s), {cores} pthreads, C port of msm.rs:84-181, window = floor(ln(n / {cores}))")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
        "config": {"workload": f"BN254 G1 variable_base_msm, uniform random scalars, known-dlog bases, 2^{log_n} points/step on host cores",
                   "points_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "parity_checked": True,
        "gpu_launches": 0,
    })


# ----------------------------------------------------------------------- small helpers
FQ_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def g1_generator(np):
    """(1, 2) as Montgomery limbs, the layout of bn256::G1Affine::generator()."""
    b = b"".join((v * (1 << 256) % FQ_MODULUS).to_bytes(32, "little") for v in (1, 2))
    return np.frombuffer(b, dtype=np.uint64).copy()


def pinned_copy(torch, np, arr):
    t = torch.empty(arr.shape, dtype=torch.int64).pin_memory()
    h = t.numpy().view(np.uint64)
    h[:] = arr
    return h, t  # keep t alive


def timed_reps(fn, reps: int):
    """fn() `reps` times after one warm-up -> (last result, {"ms_min", "ms_median", "ms_all"})."""
    fn()
    out, ms = None, []
    for _ in range(max(reps, 1)):
        t0 = time.perf_counter()
        out = fn()
        ms.append((time.perf_counter() - t0) * 1e3)
    return out, {"ms_min": min(ms), "ms_median": statistics.median(ms), "ms_all": [round(v, 3) for v in ms]}


def fr_int(limbs) -> int:
    import numpy as np

    return int.from_bytes(np.ascontiguousarray(limbs, dtype=np.uint64).tobytes(), "little") * pow(1 << 256, -1, FR_MODULUS) % FR_MODULUS


class Trapdoor:
    """The synthetic SRS has a known trapdoor: eqs[k][j] = eq_j(ss) * G (kzg.rs:174-212), so the commitment of a
    multilinear polynomial f of i variables against eqs[i] is f(ss[:i]) * G — one multilinear evaluation (threaded
    field arithmetic in the oracle) and one scalar multiplication, independent of the GPU path and of the MSM size."""

    def __init__(self, po, np, ss):
        self.po, self.np, self.ss = po, np, ss
        self.cores = po.host_threads()
        self.g = po.generator()

    def commit(self, evals):
        evals = self.np.ascontiguousarray(evals, dtype=self.np.uint64).reshape(-1, 4)
        k = evals.shape[0].bit_length() - 1
        v = self.po.evaluate_multilinear(evals, self.ss[:k], self.cores)
        return self.po.scalar_mul(self.g, fr_int(v))


# ------------------------------------------------- HyperPlonk::prove MSM-sequence surrogate
def prove_msm_sequence(pk, torch, np, k: int, dev, cpu: bool, reps: int):
    """The MSM calls HyperPlonk::prove makes for vanilla_plonk at 2^k rows:
    4 commits of 2^k points (3 witness polys + 1 permutation z-poly, backend/hyperplonk.rs:201,
    251-252) and the k quotient commitments of MultilinearKzg::open with 2^(k-1), ..., 2, 1 points
    (kzg.rs:291-293) against the resident eqs[i] slices of an SRS built on the device
    (MultilinearKzg::setup, kzg.rs:167-212).  Two forms:
      gpu_ms           every call takes host scalars (the drop-in for msm.rs alone; the opened polynomial's
                       quotients are stand-ins of the right sizes);
      gpu_resident_ms  the committed polynomials stay in HBM (batch_commit keep), g_prime is merged there
                       (multilinear.rs:203-213) and open() computes its quotients there (multilinear.rs:72-107):
                       the same 4 + k commitments plus the real quotient arithmetic, no scalar uploaded twice.
    Parity: every commitment against the SRS trapdoor (Trapdoor); with cpu=True also against the CPU port's MSMs."""
    from oracle import pyoracle as po
    from plonkish_b200 import kzg

    n = 1 << k
    ss = pk.random_scalars(k, seed=77)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pp = kzg.setup(g1_generator(np), ss)
    setup_s = time.perf_counter() - t0
    regs = pp.eqs
    polys = [pk.random_scalars(n, seed=4242 + j) for j in range(4)]   # pageable, like poly.evals()
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

    outs, t_host = timed_reps(run, reps)
    (comms, q_comms, value), t_res = timed_reps(run_resident, reps)
    res = {"k": k, "msm_calls": 4 + k, "points": 4 * n + n - 1, "gpu_ms": t_host["ms_min"], "gpu_ms_median": t_host["ms_median"],
           "gpu_resident_ms": t_res["ms_min"], "gpu_resident_ms_median": t_res["ms_median"], "reps": reps,
           "srs_setup_on_device_s": setup_s, "srs_points": 2 * n - 1, "host_scalars": "pageable"}
    # ---- parity through the trapdoor: commit(f) = f(ss) * G
    td = Trapdoor(po, np, ss)
    want_commits = [td.commit(p) for p in polys]
    ok = all((a == b).all() for a, b in zip(outs[:4], want_commits)) and all((a == b).all() for a, b in zip(comms, want_commits))
    ok = ok and all((outs[4 + j] == td.commit(host[: 1 << i])).all() for j, i in enumerate(reversed(range(k))))
    g_prime_h = po.fr_linear_combination(polys, coeffs)
    qs, want_value = po.quotients(g_prime_h, point)
    ok = ok and (value == want_value).all() and all((q_comms[i] == td.commit(qs[i])).all() for i in range(k))
    res["parity_checked"] = bool(ok)
    res["parity_how"] = "every commitment == f(ss) * G for the SRS trapdoor ss (oracle field arithmetic + one scalar multiplication); f(point) vs the oracle's quotients"
    assert ok, f"prove MSM sequence k={k}: a commitment differs from the trapdoor answer"
    if cpu:
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
        assert res["bit_exact_vs_cpu"] and res["resident_bit_exact_vs_cpu"]
    pp.release()
    return res


def srs_setup_bench(pk, torch, np, k: int, cpu: bool, reps: int):
    """fixed_base_msm + batch_normalize (msm.rs:16-31, 50-81; kzg.rs:195-208) on the GPU: 2^22 scalars host to host, and the
    CPU port on a bounded sample with the window the reference would pick for a 2^k setup."""
    from oracle import pyoracle as po

    n = 1 << 22
    sc = pk.random_scalars(n, seed=91)
    g = g1_generator(np)
    got, t = timed_reps(lambda: pk.fixed_base_msm(g, sc), reps)
    res = {"what": "fixed_base_msm + batch_normalize of 2^22 scalars, pageable host scalars in, affine points out (table build included)",
           "gpu_mpoints_per_s": n / t["ms_min"] / 1e3, "gpu_ms": t["ms_min"], "gpu_ms_median": t["ms_median"], "reps": reps}
    # parity: sum_i got[i] = (sum_i sc[i]) * G on a strided sample through the oracle, and (cpu) the CPU port element-wise
    idx = np.arange(0, n, n // 4096)
    spot = all((got[i] == po.scalar_mul(g, fr_int(sc[i]))).all() for i in idx[:64])
    res["parity_checked"] = bool(spot)
    res["parity_how"] = "64 strided outputs == scalar * G by the oracle's double-and-add" + ("; first 2^17 outputs vs the CPU port" if cpu else "")
    if cpu:
        m = 1 << 17
        window = po.window_size((2 << k) - 2)
        cores = po.host_threads()
        t0 = time.perf_counter()
        want = po.fixed_base_msm(g, sc[:m], window=window, num_threads=cores)
        sec_c = time.perf_counter() - t0
        res.update({"cpu_mpoints_per_s": m / sec_c / 1e6, "cpu_cores": cores, "cpu_sample": f"2^17 scalars, window {window} (table build included)",
                    "bit_exact_vs_cpu": bool((got[:m] == want).all())})
        res["parity_checked"] = bool(spot and res["bit_exact_vs_cpu"])
    assert res["parity_checked"], "fixed_base_msm differs from the oracle"
    return res


def sum_check_bench(pk, torch, np, k: int, cpu: bool, reps: int):
    """The zero check of HyperPlonk::prove for vanilla_plonk (backend/hyperplonk.rs:262-277) as ClassicSumCheck runs it
    (piop/sum_check/classic.rs:208-240): 9 tables (eq, 5 selectors, 3 witness columns) of 2^k evaluations, degree 4,
    k rounds of round-polynomial evaluations + table folds on the GPU; the challenges are stand-ins for the transcript's.
    Parity: every round message and the final evaluations against the oracle's restatement on all host cores (the same
    2^k tables); cpu_ms is the oracle's single-threaded time at k = 18."""
    from oracle import pyoracle as po
    from plonkish_b200 import sumcheck

    n = 1 << k
    host_tables = [pk.random_scalars(n, seed=300 + i) for i in range(9)]
    tables = [pk.ResidentScalars(t) for t in host_tables]
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

    (msgs, finals), t = timed_reps(run, reps)
    res = {"what": "zero check of vanilla_plonk as ClassicSumCheck<EvaluationsProver> runs it: 9 resident tables of 2^k evaluations, degree 4, "
                   "k rounds (round-polynomial evaluations at X = 1..4 + fold of every table), stand-in challenges",
           "k": k, "gpu_ms": t["ms_min"], "gpu_ms_median": t["ms_median"], "reps": reps,
           "gpu_mpairs_per_s_first_round_equiv": (n - 1) / t["ms_min"] / 1e3}
    cores = po.host_threads()
    cur, ok = host_tables, True
    t0 = time.perf_counter()
    for rnd in range(k):
        ok = ok and msgs[rnd].tobytes() == po.sumcheck_round(cur, terms, 0, num_threads=cores if len(cur[0]) >= 1 << 12 else 1).tobytes()
        cur = [po.fix_var(p, chal[rnd], cores if len(p) >= 1 << 14 else 1) for p in cur]
    ok = ok and finals.tobytes() == np.stack([p[0] for p in cur]).tobytes()
    res.update({"parity_checked": bool(ok), "parity_how": f"all {k} round messages and the 9 final evaluations == the oracle's restatement on {cores} threads",
                "oracle_check_s": time.perf_counter() - t0})
    assert ok, "sum-check rounds differ from the oracle"
    if cpu:
        kc = min(18, k)
        host = [t_[: 1 << kc] for t_ in host_tables]
        cur = host
        t0 = time.perf_counter()
        for rnd in range(kc):
            po.sumcheck_round(cur, terms, 0)
            cur = [po.fix_var(p, chal[rnd]) for p in cur]
        cpu_ms = (time.perf_counter() - t0) * 1e3
        res.update({"cpu_ms": cpu_ms, "cpu_k": kc, "cpu_cores": 1, "cpu_mpairs_per_s": ((1 << kc) - 1) / cpu_ms / 1e3})
    for t_ in tables:
        t_.release()
    return res


def oracle_batch_open(po, np, commit, num_vars, polys, points, evals, transcript, cores):
    """additive::batch_open (pcs/multilinear.rs:134-235) restated over the oracle's field kernels (tests/batch_open_ref.py
    is the all-integer restatement the unit tests use; this one scales to 2^24): merged polynomial per point, the degree-2
    sum check in coefficient form (classic/coeff.rs:132-146), g_prime, open.  polys: host [2^k, 4] arrays; points: lists of
    canonical integers; evals: (poly, point, value)."""
    from plonkish_b200.sumcheck import _to_int, _to_mont

    r = FR_MODULUS
    ell = max(len(evals) - 1, 0).bit_length()
    t = transcript.squeeze_challenges(ell)
    eq_xt = [1]
    for v in t:
        eq_xt = [e * (1 - v) % r for e in eq_xt] + [e * v % r for e in eq_xt]
    by_point = [[] for _ in points]
    for (poly, point, _), w in zip(evals, eq_xt):
        by_point[point].append((poly, w))
    merged = []
    for entries in by_point:                                   # multilinear.rs:150-167
        if len(entries) == 1:
            merged.append((entries[0][1], polys[entries[0][0]]))
        else:
            merged.append((1, po.fr_affine(1 << num_vars, [polys[i] for i, _ in entries], np.stack([_to_mont(w) for _, w in entries]), num_threads=cores)))
    claim = sum(v * w for (_, _, v), w in zip(evals, eq_xt)) % r
    P = len(points)
    cur = [po.kzg_eq_scalars(np.stack([_to_mont(v) for v in pt]))[num_vars] for pt in points] + [m for _, m in merged]
    terms = [(_to_mont(scalar), [j, P + j]) for j, (scalar, _) in enumerate(merged)]
    inv2 = pow(2, -1, r)
    challenges = []
    for _ in range(num_vars):
        thr = cores if len(cur[0]) >= 1 << 14 else 1
        h1, h2 = (_to_int(x) for x in po.sumcheck_round(cur, terms, -1, num_threads=thr))
        c0 = (claim - h1) % r
        c2 = (h2 - 2 * h1 + c0) * inv2 % r
        c1 = (claim - 2 * c0 - c2) % r
        transcript.write_field_elements([c0, c1, c2])
        ch = transcript.squeeze_challenge()
        challenges.append(ch)
        claim = (c0 + ch * (c1 + ch * c2)) % r
        cur = [po.fix_var(p, _to_mont(ch), thr) for p in cur]
    coeffs = []
    for (scalar, _), pt in zip(merged, points):               # multilinear.rs:203-213
        e = 1
        for a_, b_ in zip(challenges, pt):
            e = e * ((a_ * b_ + (1 - a_) * (1 - b_)) % r) % r
        coeffs.append(scalar * e % r)
    g_prime = po.fr_affine(1 << num_vars, [m for _, m in merged], np.stack([_to_mont(c) for c in coeffs]), num_threads=cores)
    qs, _ = po.quotients(g_prime, np.stack([_to_mont(c) for c in challenges]))
    transcript.write_commitments([commit(q, i) for i, q in enumerate(qs)])


def synth_vanilla_plonk_circuit(pk, po, np, k: int, seed: int):
    """A satisfied vanilla_plonk circuit of 2^k rows in the shape of rand_vanilla_plonk_circuit (backend/hyperplonk/util.rs:
    100-169), generated vectorised: k random public inputs on the rows bh[1..k]; every row is an addition gate (q_l = q_r = 1,
    q_o = -1) or a multiplication gate (q_m = 1, q_o = -1) with a random q_c, w_o set to close the gate, the last row zero;
    about half of the rows of the upper half copy w_l and w_r from random cells (w_l / w_r / w_o, row >= 1) of the lower
    half.  Returns (instances as integers, [q_l, q_r, q_m, q_o, q_c], [w_l, w_r, w_o] as Montgomery arrays, the three
    permutation polynomials of preprocessor.rs:172-203 as canonical uint64 columns)."""
    from plonkish_b200.sumcheck import _to_mont

    n, half = 1 << k, 1 << (k - 1)
    rng = np.random.default_rng(seed)
    cores = po.host_threads()
    one, minus_one, zero = _to_mont(1), _to_mont(FR_MODULUS - 1), np.zeros(4, dtype=np.uint64)
    instances = [fr_int(row) for row in pk.random_scalars(k, seed=seed + 1)]
    order = po.bh_iter(k)
    pi = np.zeros((n, 4), dtype=np.uint64)
    for i, v in enumerate(instances):
        pi[int(order[i + 1])] = _to_mont(v)
    is_add = rng.integers(0, 2, n).astype(bool)
    q_c = pk.random_scalars(n, seed=seed + 2)
    w_l, w_r = pk.random_scalars(n, seed=seed + 3), pk.random_scalars(n, seed=seed + 4)

    def close(lo, hi):  # w_o of the rows [lo, hi): w_l + w_r + q_c + pi or w_l * w_r + q_c + pi
        s_ = po.fr_vec_op("add", w_l[lo:hi], w_r[lo:hi], cores)
        m_ = po.fr_vec_op("mul", w_l[lo:hi], w_r[lo:hi], cores)
        body = np.where(is_add[lo:hi, None], s_, m_)
        return po.fr_vec_op("add", po.fr_vec_op("add", body, q_c[lo:hi], cores), pi[lo:hi], cores)

    w_o = np.zeros((n, 4), dtype=np.uint64)
    w_o[:half] = close(0, half)
    cols = [w_l, w_r, w_o]
    # copies (util.rs:116-128): target rows in [half, n - 1), sources in rows [1, half)
    copy_rows = np.nonzero(rng.integers(0, 2, n - 1 - half).astype(bool))[0] + half
    src_col = rng.integers(0, 3, (2, copy_rows.size))
    src_row = rng.integers(1, half, (2, copy_rows.size))
    for side, target in enumerate((w_l, w_r)):
        for c in range(3):
            sel = src_col[side] == c
            target[copy_rows[sel]] = cols[c][src_row[side][sel]]
    w_o[half:] = close(half, n)
    q_l = np.where(is_add[:, None], one, zero)
    q_r = q_l.copy()
    q_m = np.where(is_add[:, None], zero, one)
    q_o = np.tile(minus_one, (n, 1))
    for arr in (q_l, q_r, q_m, q_o, q_c, w_l, w_r, w_o):   # the last row is never assigned (util.rs:115)
        arr[n - 1] = 0
    # permutation polynomials: every cycle is a source cell and its copies, sorted by (poly, row) (util.rs:398-404); the
    # cell after c in the cycle takes c's id (preprocessor.rs:191-198)
    targets = np.concatenate([0 * n + copy_rows, 1 * n + copy_rows]).astype(np.int64)
    sources = np.concatenate([src_col[0] * n + src_row[0], src_col[1] * n + src_row[1]]).astype(np.int64)
    uniq = np.unique(sources)
    cells = np.concatenate([uniq, targets])
    group = np.concatenate([uniq, sources])
    o = np.lexsort((cells, group))
    cells, group = cells[o], group[o]
    first = np.r_[True, group[1:] != group[:-1]]
    last = np.r_[first[1:], True]
    prev = np.r_[cells[-1:], cells[:-1]]
    # at the first cell of a cycle the predecessor is the cycle's last cell
    last_of_group = np.repeat(cells[last], np.diff(np.r_[np.nonzero(first)[0], cells.size]))
    prev = np.where(first, last_of_group, prev)
    sigma = np.arange(3 * n, dtype=np.uint64)
    sigma[cells] = prev.astype(np.uint64)
    return instances, [q_l, q_r, q_m, q_o, q_c], [w_l, w_r, w_o], [sigma[i * n:(i + 1) * n].copy() for i in range(3)]


def oracle_hyperplonk_prove(po, np, commit, k, expression, instances, host_polys, sigma_mont, transcript, cores):
    """HyperPlonk::prove (backend/hyperplonk.rs:164-291) through the oracle's field kernels on the host cores: the same
    compiled expression (plonkish_b200/expression.py, pinned against the hand-written integer prover in tests/), every
    table built, summed and folded by the C port.  host_polys: pi, q_l, q_r, q_m, q_o, q_c, w_l, w_r, w_o as [2^k, 4]
    arrays.  Returns nothing; the proof goes to `transcript`."""
    from plonkish_b200 import hyperplonk as hp
    from plonkish_b200.expression import compile_expression
    from plonkish_b200.sumcheck import _to_int, _to_mont, interpolate_at

    n = 1 << k
    for v in instances:
        transcript.common_field_element(v)
    witness = host_polys[6:9]
    transcript.write_commitments([commit(w, k) for w in witness])
    beta = transcript.squeeze_challenge()
    gamma = transcript.squeeze_challenge()
    (z,) = po.permutation_z_polys(1, witness, sigma_mont, _to_mont(beta), _to_mont(gamma), num_threads=cores)
    transcript.write_commitments([commit(z, k)])
    alpha = transcript.squeeze_challenge()
    y = transcript.squeeze_challenges(k)
    polys = list(host_polys) + list(sigma_mont) + [z]
    compiled = compile_expression(expression, [beta, gamma, alpha])
    bh = hp.BooleanHypercube(k)
    b_ = np.arange(n, dtype=np.uint64)
    maps = {}

    def rotation_map(rot):                                    # BooleanHypercube::rotation_map (bh.rs:135-137)
        if rot not in maps:
            m = b_.copy()
            for _ in range(rot):
                m = (m << np.uint64(1)) ^ ((m >> np.uint64(k - 1)) * np.uint64(bh.primitive))
            for _ in range(-rot):
                m = (m >> np.uint64(1)) ^ ((m & np.uint64(1)) * np.uint64(bh.x_inv))
            maps[rot] = m.astype(np.uint32)
        return maps[rot]

    eq = po.kzg_eq_scalars(np.stack([_to_mont(v) for v in y]))[k]
    tables, query_table = [], {}
    for atom in compiled.atoms:
        if atom.is_leaf() and atom.leaf()[0] == "poly" and atom.leaf()[2] == 0:
            query_table[atom.leaf()[1]] = len(tables)
            tables.append(polys[atom.leaf()[1]])
            continue
        if atom.is_leaf() and atom.leaf()[0] == "eq_xy":
            tables.append(eq)
            continue
        srcs, coeffs, rows, sparse, id_coeff = [], [], [], [], None
        for leaf, c in atom.terms.items():
            if leaf[0] == "poly":
                srcs.append(polys[leaf[1]]); coeffs.append(c); rows.append(rotation_map(leaf[2]) if leaf[2] else None)
            elif leaf[0] == "eq_xy":
                srcs.append(eq); coeffs.append(c); rows.append(None)
            elif leaf[0] == "identity":
                id_coeff = c
            else:
                sparse.append((bh.nth(leaf[1] % n), c))
        tab = po.fr_affine(n, srcs, np.stack([_to_mont(c) for c in coeffs]) if srcs else None, rows, _to_mont(atom.const) if atom.const else None,
                           None if id_coeff is None else _to_mont(id_coeff), cores)
        for row, c in sparse:
            tab[row] = po.fe_op("add", 1, tab[row], _to_mont(c))
        tables.append(tab)
    queries = hp.pcs_query(expression, 1)
    for q in queries:
        if q.poly not in query_table:
            query_table[q.poly] = len(tables)
            tables.append(polys[q.poly])
    terms = [(_to_mont(c), idx) for c, idx in compiled.terms]
    claim, x = 0, []
    cur = tables
    for _ in range(k):
        thr = cores if len(cur[0]) >= 1 << 12 else 1
        tail = [_to_int(r_) for r_ in po.sumcheck_round(cur, terms, compiled.common, num_threads=thr)]
        msg = [(claim - tail[0]) % FR_MODULUS] + tail
        transcript.write_field_elements(msg)
        ch = transcript.squeeze_challenge()
        x.append(ch)
        claim = interpolate_at(msg, ch)
        cur = [po.fix_var(p, _to_mont(ch), thr) for p in cur]
    offsets = hp.point_offset(queries)
    evals = []
    for q in queries:
        if q.rotation == 0:
            values = [_to_int(cur[query_table[q.poly]][0])]
        else:
            values = [_to_int(po.evaluate_multilinear(polys[q.poly], np.stack([_to_mont(v) for v in pt]), cores)) for pt in hp.rotation_eval_points(x, q.rotation)]
        evals.extend((q.poly, offsets[q.rotation] + j, v) for j, v in enumerate(values))
    transcript.write_field_elements([v for _, _, v in evals])
    oracle_batch_open(po, np, commit, k, polys, hp.points(queries, x), evals, transcript, cores)


def hyperplonk_prove_bench(pk, torch, np, k: int, cpu: bool, reps: int):
    """HyperPlonk::prove for vanilla_plonk (backend/hyperplonk.rs:164-291; `cargo bench --bench proof_system -- --system
    hyperplonk --circuit vanilla_plonk --k K`, BASELINE.json configs 1 and 3) on a satisfied synthetic circuit of 2^k rows:
    the complete proof — witness commitments, the permutation grand product and its commitment, the zero check over the
    composed expression (gate + l_1 (z - 1) + the permutation constraint with z at the rotated row, degree 5), the
    evaluations at x and at the two rotated points, additive::batch_open over three points — with every polynomial
    operation on the GPU and the Keccak256 transcript on the host.  Witness polynomials start in pageable host memory;
    preprocess (selector and permutation commitments) is outside the timed region as in the reference's bench
    (benchmark/benches/proof_system.rs).  Parity: the proof is accepted by the integer restatement of the reference's
    verifier (tests/hyperplonk_ref.py: hyperplonk.rs:293-362 with the pairing equation checked through the SRS
    trapdoor); with cpu=True the proof BYTES are also compared with the oracle's prover (C port, all host cores),
    whose time is the CPU number beside."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    import hyperplonk_ref as ref
    from oracle import bigint_ref as br
    from oracle import pyoracle as po
    from plonkish_b200 import hyperplonk, kzg
    from plonkish_b200.sumcheck import _to_mont
    from plonkish_b200.transcript import Keccak256Transcript

    n = 1 << k
    cores = po.host_threads()
    ss = pk.random_scalars(k, seed=601)
    pp = kzg.setup(g1_generator(np), ss)
    t0 = time.perf_counter()
    instances, preprocess, witness, sigma = synth_vanilla_plonk_circuit(pk, po, np, k, seed=610)
    gen_s = time.perf_counter() - t0
    info = hyperplonk.vanilla_plonk_circuit_info(k, k, preprocess, [[(6, 1)], [(7, 1)], [(8, 1)]])
    t0 = time.perf_counter()
    hpp, hvp = hyperplonk.preprocess(pp, info, permutation_columns=sigma)
    preprocess_ms = (time.perf_counter() - t0) * 1e3

    class Circuit:
        def instances(self):
            return [instances]

        def synthesize(self, rnd, challenges):
            return witness

    circuit = Circuit()
    phases = []

    def run():
        t, marks = Keccak256Transcript(), []
        hyperplonk.prove(hpp, circuit, t, marks)
        phases.append({b_[0]: round((b_[1] - a_[1]) * 1e3, 2) for a_, b_ in zip(marks, marks[1:])})
        return t.into_proof()

    proof, tm = timed_reps(run, reps)
    res = {"what": "HyperPlonk::prove for vanilla_plonk, the complete proof: batch_commit of 3 witness polynomials (pageable host memory) -> permutation "
                   "grand product z + commit -> zero check over the composed expression incl. the permutation constraint with z rotated (23 tables, "
                   "degree 5) -> 14 evaluations (2 at the rotated points) -> additive::batch_open over 3 points; Keccak256 transcript on the host; "
                   "witness generation and preprocess outside the timed region",
           "k": k, "gpu_ms": tm["ms_min"], "gpu_ms_median": tm["ms_median"], "gpu_ms_all": tm["ms_all"], "reps": reps, "proof_bytes": len(proof),
           "phases_ms": phases[1:], "preprocess_ms": preprocess_ms, "circuit_generation_s": gen_s}
    # ---- the reference's verifier, restated with integers, on the GPU's proof
    t0 = time.perf_counter()
    affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
    ref.verify_reference(po.keccak256, [fr_int(s_) for s_ in ss], k, instances, [affine(c) for c in hvp.preprocess_comms],
                         [affine(c) for _, c in hvp.permutation_comms], proof)
    res.update({"verifier_accepts": True, "verifier_s": time.perf_counter() - t0})
    how = "the proof is accepted by the integer restatement of HyperPlonk::verify (pairing equation through the SRS trapdoor)"
    if cpu:
        eqs_h = [e.to_host() for e in pp.eqs]
        commit = lambda f, i: po.variable_base_msm(f, eqs_h[i], cores)  # noqa: E731
        order = po.bh_iter(k)
        pi = np.zeros((n, 4), dtype=np.uint64)
        for i, v in enumerate(instances):
            pi[int(order[i + 1])] = _to_mont(v)
        canon = np.zeros((3 * n, 4), dtype=np.uint64)
        canon[:, 0] = np.concatenate(sigma)
        sigma_mont = [a_.copy() for a_ in np.split(po.from_canonical(1, canon), 3)]
        t0 = time.perf_counter()
        t = Keccak256Transcript()
        oracle_hyperplonk_prove(po, np, commit, k, hvp.expression, instances, [pi] + preprocess + witness, sigma_mont, t, cores)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        same = bool(t.into_proof() == proof)
        res.update({"cpu_ms": cpu_ms, "cpu_cores": cores, "cpu_how": "C port of the prover on the host cores: MSMs (msm.rs:84-181), tables, sum-check rounds "
                    "and folds threaded; quotients and the transcript single-threaded", "proof_bytes_identical_to_cpu": same})
        assert same, f"hyperplonk prove k={k}: proof bytes differ from the oracle prover's"
        how += "; proof bytes identical to the oracle prover's (CPU port)"
    res["parity_checked"] = True
    res["parity_how"] = how
    hpp.release()
    pp.release()
    return res


# ------------------------------------------------- the other two BN254 multilinear PCSs (Zeromorph, Gemini)
def _timed_ops(base, names):
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


def pcs_schemes_bench(pk, torch, np, k: int, reps: int, with_prove: bool = True, which: str = "both", prefix_tables: bool = False):
    """Zeromorph<UnivariateKzg> and Gemini<UnivariateKzg> (pcs/multilinear/zeromorph.rs, gemini.rs; the reference tests
    HyperPlonk over them at backend/hyperplonk.rs:425-426) on one GPU: commit + open of one 2^k-evaluation polynomial
    with per-operation times, and HyperPlonk::prove for vanilla_plonk over the scheme on the synthetic circuit of
    hyperplonk_prove_bench.  Parity: every opening satisfies the scheme's verifier equation (tests/zeromorph_ref.py,
    tests/gemini_ref.py: the pairing check evaluated in G1 through the SRS trapdoor) for the oracle's evaluation of the
    polynomial, and every HyperPlonk proof is accepted by the verifier restatement with that check as its last step."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    import gemini_ref as gr
    import hyperplonk_ref as ref
    import zeromorph_ref as zr
    from oracle import bigint_ref as br
    from oracle import pyoracle as po
    from plonkish_b200 import gemini, hyperplonk, kzg, zeromorph
    from plonkish_b200.sumcheck import _to_int, _to_mont
    from plonkish_b200.transcript import Keccak256Transcript

    n, s = 1 << k, 0x2468ACE13579BDF2468ACE13579BDF % br.R
    t0 = time.perf_counter()
    powers = kzg.univariate_setup(g1_generator(np), _to_mont(s), n)
    out = {"what": "commit + open of one 2^k-evaluation polynomial and HyperPlonk::prove for vanilla_plonk over Zeromorph / Gemini on the univariate "
                   "KZG SRS; open_ops_ms: wall time per polynomial operation of the best open",
           "k": k, "reps": reps, "srs_setup_on_device_s": time.perf_counter() - t0, "prefix_tables": prefix_tables}
    if prefix_tables:  # own resident slices for the prefixes the quotient / fold commitments run against
        t0 = time.perf_counter()
        powers.prefix_tables = zeromorph.build_prefix_tables(powers, n)
        out["prefix_tables_s"] = time.perf_counter() - t0
    poly_h = pk.random_scalars(n, seed=1234)
    poly = pk.ResidentScalars(poly_h)
    circuit_parts = synth_vanilla_plonk_circuit(pk, po, np, k, seed=610) if with_prove else None
    schemes = {
        "zeromorph": (zeromorph, zeromorph.trim(powers, n), ("quotients", "commit_quotients", "q_hat", "f", "div_linear", "commit"),
                      lambda reader, c, pt, val: zr.verify_reader_in_g1(reader, c, pt, val, s)),
        "gemini": (gemini, gemini.GeminiKzgProverParam(powers), ("folds", "commit_folds", "evaluate", "linear_combination", "div_linear", "commit"),
                   lambda reader, c, pt, val: gr.verify_reader_in_g1(reader, c, pt, val, s)),
    }
    for name, (mod, pp, op_names, pcs_verify) in schemes.items():
        if which not in ("both", name):
            continue
        ops = _timed_ops(mod.GpuOps, op_names)

        def run():
            ops.spans = {}
            t = Keccak256Transcript()
            comm = mod.commit(pp, poly)
            t.write_commitment(comm)
            point = t.squeeze_challenges(k)
            t.write_field_element(0)           # stands for the evaluation (written before open; the proof does not depend on it)
            t1 = time.perf_counter()
            value = mod.open(pp, poly, point, 0, t, ops) if mod is zeromorph else mod.open(pp, poly, point, t, ops)
            return point, value, t.into_proof(), (time.perf_counter() - t1) * 1e3, dict(ops.spans)

        runs = []
        _, tm = timed_reps(lambda: runs.append(run()), reps)
        best = min(runs[1:], key=lambda r_: r_[3])
        point, value, proof = best[0], best[1], best[2]
        res = {"commit_plus_open_ms": tm["ms_min"], "commit_plus_open_ms_median": tm["ms_median"], "open_ms": best[3],
               "open_ops_ms": {a_: round(b_, 3) for a_, b_ in best[4].items()}, "open_proof_bytes": len(proof) - 96}
        want = _to_int(po.evaluate_multilinear(poly_h, zr.mont_rows(point), po.host_threads()))
        assert value is None or value == want, f"{name}: the remainder of quotients differs from the oracle's evaluation"
        reader = ref.ProofReader(po.keccak256, proof)
        c = reader.read_commitment()
        assert reader.squeeze_challenges(k) == point
        reader.read_field_element()
        pcs_verify(reader, c, point, want)
        assert reader.pos == len(proof)
        res["parity_checked"] = True
        res["parity_how"] = "the opening satisfies the scheme's verify equation (G1, SRS trapdoor) for the oracle's multilinear evaluation"
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

            hp_proof, tp = timed_reps(prove, reps)
            affine = lambda limbs: br.point_from_bytes(np.ascontiguousarray(limbs).tobytes())  # noqa: E731
            ref.verify_reference(po.keccak256, None, k, instances, [affine(c_) for c_ in hvp.preprocess_comms],
                                 [affine(c_) for _, c_ in hvp.permutation_comms], hp_proof, pcs_verify=pcs_verify)
            res["hyperplonk_prove"] = {"gpu_ms": tp["ms_min"], "gpu_ms_median": tp["ms_median"], "proof_bytes": len(hp_proof), "phases_ms": phases[-1],
                                       "verifier_accepts": True, "parity_checked": True}
            hpp.release()
        out[name] = res
    poly.release()
    zeromorph.release_prefix_tables(powers)
    powers.release()
    return out


def univariate_sequence(pk, torch, np, k: int, dev, reps: int):
    """BASELINE.json config 4 restated synthetically: UnivariateKzg commit = one MSM over
    the SRS prefix (pcs/univariate/kzg.rs:24-30, witness-like canonical 68-bit values with zero padding) and
    batch_open = two MSMs of ~2^k uniform coefficients (univariate/kzg.rs:330,353), pageable host scalars,
    resident powers_of_s_g1.  Parity: known discrete log of the synthetic SRS."""
    from oracle import pyoracle as po

    n = 1 << k
    d_bases = pk.synth_bases_device(n, 11, 13, device=dev)
    torch.cuda.synchronize()
    reg = pk.G1Bases(d_bases)
    rng = np.random.default_rng(22)
    canon = np.zeros((n, 4), dtype=np.uint64)
    live = n - n // 8                       # last eighth zero padding
    canon[:live, 0] = rng.integers(0, 1 << 63, size=live, dtype=np.uint64)
    canon[:live, 1] = rng.integers(0, 1 << 4, size=live, dtype=np.uint64)   # canonical 68-bit limb values (aggregation circuit limbs)
    limbs = po.from_canonical(1, canon)     # their Montgomery representations are full width: that is what the MSM is handed
    uniform = pk.random_scalars(n, seed=2222)
    host = [limbs, uniform, uniform]

    outs, t = timed_reps(lambda: [pk.variable_base_msm(h, reg) for h in host], reps)
    ok = all((o == po.known_dlog_answer(11, 13, h)).all() for o, h in zip(outs, host))
    reg.release()
    assert ok, "univariate sequence: a commitment differs from the known-dlog answer"
    return {"k": k, "msm_calls": 3, "points": 3 * n, "gpu_ms": t["ms_min"], "gpu_ms_median": t["ms_median"], "reps": reps, "parity_checked": bool(ok),
            "parity_how": "known discrete log of the synthetic SRS",
            "what": "commit (68-bit canonical values in Montgomery form, 1/8 zero padding) + batch_open (2 uniform MSMs) of 2^k points each, pageable host scalars"}


# ----------------------------------------------------------------------------- skew set
def skew_scalars(pk, po, np, kind: str, n: int, rng):
    def mont_const(v):
        return np.frombuffer((v % FR_MODULUS * (1 << 256) % FR_MODULUS).to_bytes(32, "little"), dtype=np.uint64)

    out = np.zeros((n, 4), dtype=np.uint64)
    if kind == "selector":      # 0 / 1 / -1 with half zeros (backend/hyperplonk/util.rs:133-152)
        pick = rng.integers(0, 4, n)
        out[pick == 2] = mont_const(1)
        out[pick == 3] = mont_const(FR_MODULUS - 1)
    elif kind == "small-ints":  # permutation polynomials: values < 3 * 2^k (backend/hyperplonk/preprocessor.rs:184-190)
        canon = np.zeros((n, 4), dtype=np.uint64)
        canon[:, 0] = rng.integers(0, 3 * n, n, dtype=np.uint64)
        out = po.from_canonical(1, canon)
    elif kind == "all-ones":
        out[:] = mont_const(1)
    elif kind == "same-wide":
        out[:] = mont_const(0x2AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA)
    return out


def skew_bench(pk, torch, np, d_uniform, reg, uniform_ms: float, n: int, dev):
    """SURVEY.md §8(d) skew set on the timed configuration (table layout, device-resident scalars): preprocess-shaped
    scalars.  Each leg is checked against the known-dlog answer; `vs_uniform` = ms / the uniform step."""
    from oracle import pyoracle as po

    rng = np.random.default_rng(1)
    res = {"uniform_ms": uniform_ms}
    for kind in ("selector", "small-ints", "all-ones", "same-wide"):
        sc = skew_scalars(pk, po, np, kind, n, rng)
        d_sc = torch.from_numpy(sc.view(np.int64)).to(dev)
        got = pk.variable_base_msm_device(d_sc, reg).cpu().numpy().view(np.uint64)
        ok = bool((got == po.known_dlog_answer(3, 5, sc)).all())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(3):
            e0.record()
            pk.variable_base_msm_device(d_sc, reg)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[kind] = {"ms": min(ts), "ms_median": statistics.median(ts), "vs_uniform": min(ts) / uniform_ms, "parity_checked": ok}
        assert ok, f"skew leg {kind}: result differs from the known-dlog answer"
        del d_sc
    res["slowest_vs_uniform"] = max(v["vs_uniform"] for v in res.values() if isinstance(v, dict))
    return res


# ------------------------------------------------------------------------- our arm
def sum_points(points_u64):
    """Affine sum of [k, 8] Montgomery points with the big-integer reference (independent of the C oracle and the GPU)."""
    from oracle import bigint_ref as br

    acc = None
    for p in points_u64:
        acc = br.add(acc, br.point_from_bytes(p.tobytes()))
    return br.point_to_bytes(acc)


def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import plonkish_b200 as pk
    from plonkish_b200 import _lib
    from oracle import pyoracle as po  # the checker: every timed result is compared with an independent answer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")  # host-side barrier while rank 0 drives all GPUs by itself
    _lib.lib()
    po.build()

    n = 1 << args.log_n
    total_n = n * world
    a, d = 3, 5
    first = rank * n
    # synthetic inputs (seeded): this rank's slice of the (world * n)-point MSM.  The host copy is plain numpy
    # memory — pageable, like the Vec<Fr> behind poly.evals() — plus a pinned copy for the pinned-source number.
    scalars_np = pk.random_scalars(n, seed=1000 + rank)
    scalars_pin, _keep_pin = pinned_copy(torch, np, scalars_np)
    d_scalars = torch.from_numpy(scalars_np.view(np.int64)).to(dev)
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

    def expected_total(local_scalars, local_first):
        """The known-dlog answer of the whole (all ranks) MSM: every rank evaluates its slice's answer with the oracle
        (two field sums + one scalar multiplication), rank 0 adds the points with the big-integer reference."""
        mine = po.known_dlog_answer(a + local_first * d, d, local_scalars)
        if not distributed:
            return mine.tobytes()
        t = torch.from_numpy(mine.view(np.int64)).to(dev)
        allp = torch.empty(world * 8, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allp, t)
        return sum_points(allp.cpu().numpy().view(np.uint64).reshape(world, 8))

    want = expected_total(scalars_np, first)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_device(step, steps, warmup):
        for _ in range(warmup):
            out = step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]  # per-step boundaries for min / median
        barrier()
        e0.record()
        for i in range(steps):
            out = step()
            marks[i].record()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        step_ms = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(steps)]
        if distributed:
            t = torch.tensor([ms_total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        return out, ms_total / steps, step_ms

    def timed_host(step, steps, warmup=2):
        for _ in range(warmup):
            out = step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = step()
        barrier()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        if distributed:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return out, ms

    def step_device():
        if distributed:
            return pk.variable_base_msm_sharded(d_scalars, step_bases, window_bits=args.window_bits)
        return pk.variable_base_msm_device(d_scalars, step_bases, window_bits=args.window_bits)

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    barrier()
    launches0 = pk.launch_count()
    t_wall0 = time.time()
    out, ms_per_step, step_ms = timed_device(step_device, args.steps, 0)
    t_wall1 = time.time()
    launches = pk.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    value = total_n / (ms_per_step * 1e-3) / 1e6
    result_dev = out.cpu().numpy().view(np.uint64)
    assert result_dev.tobytes() == want, "the timed device-resident result differs from the known-dlog answer"

    # ---- e2e: host buffers through the public API, copies inside the timed region
    def make_e2e(src):
        if not distributed:
            return lambda: pk.variable_base_msm(src, reg)
        return lambda: pk.variable_base_msm_sharded_host(src, reg).cpu().numpy().view(np.uint64)

    staged0 = pk.staged_bytes()
    e2e_out, e2e_ms = timed_host(make_e2e(scalars_np), args.steps)
    staged_per_step = (pk.staged_bytes() - staged0) // (args.steps + 2)
    staging_rate = pk.staging_rate_gbps()  # this rank's; below ~35 GB/s the library cuts the upload into five chunks instead of three
    assert np.asarray(e2e_out).view(np.uint64).tobytes() == want, "the e2e (pageable) result differs from the known-dlog answer"
    e2e_pin_out, e2e_pin_ms = timed_host(make_e2e(scalars_pin), args.steps)
    assert np.asarray(e2e_pin_out).view(np.uint64).tobytes() == want, "the e2e (pinned) result differs from the known-dlog answer"
    e2e_value = total_n / (e2e_ms * 1e-3) / 1e6

    # ---- strong scaling: ONE MSM of 2^log_n points in total, point-sharded over the ranks (BASELINE: "at 2^24 on 1/2/4/8")
    strong = None
    if distributed and not args.plain_bases:
        n_s = n // world
        d_bases_s = pk.synth_bases_device(n_s, a, d, device=dev, first=rank * n_s)
        torch.cuda.synchronize()
        reg_s = pk.G1Bases(d_bases_s)
        d_scalars_s = d_scalars[:n_s].contiguous()
        want_s = expected_total(scalars_np[:n_s], rank * n_s)
        out_s, ms_s, step_ms_s = timed_device(lambda: pk.variable_base_msm_sharded(d_scalars_s, reg_s), args.steps, 3)
        assert out_s.cpu().numpy().view(np.uint64).tobytes() == want_s, "strong-scaling result differs from the known-dlog answer"
        src_s = scalars_np[:n_s]
        e2e_s_out, e2e_s_ms = timed_host(lambda: pk.variable_base_msm_sharded_host(src_s, reg_s).cpu().numpy().view(np.uint64), args.steps)
        assert np.asarray(e2e_s_out).view(np.uint64).tobytes() == want_s
        plan_s = pk.msm_plan(n_s, 0, local_rank, bases=reg_s)
        strong = {
            "what": f"one MSM of 2^{args.log_n} points in total, {n_s} per GPU, device-resident slices, NCCL all_gather of the partials; max over ranks",
            "total_points": n, "points_per_gpu": n_s, "ms_per_step": ms_s, "ms_step_min": min(step_ms_s), "value": n / (ms_s * 1e-3) / 1e6, "unit": UNIT,
            "e2e_ms_per_step": e2e_s_ms, "e2e_value": n / (e2e_s_ms * 1e-3) / 1e6, "e2e_source": "pageable host scalars",
            "window_bits": plan_s["window_bits"], "windows": plan_s["windows"],
            # the weak step of this same run is one GPU doing 2^log_n points: the 1-GPU time of the strong-scaling problem
            "one_gpu_ms_same_run": ms_per_step, "efficiency_vs_one_gpu_same_run": ms_per_step / (world * ms_s),
            "parity_checked": True,
        }
        reg_s.release()
        del d_bases_s, d_scalars_s

    # ---- roofline of the dominant kernel (K3 accumulate), timed live with CUDA events
    plan = pk.msm_plan(n, args.window_bits, local_rank, bases=None if step_bases is d_bases else reg)
    stage_runs = [pk.profile_stages_device(d_scalars, step_bases, window_bits=args.window_bits) for _ in range(3)]
    stages = {k: statistics.mean(r[k] for r in stage_runs) for k in stage_runs[0]}
    pipe = pk.bench_integer_pipe(local_rank)
    madd_streams = pk.bench_madd(local_rank)
    fp64 = pk.bench_fp64_pipe(local_rank)
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
        decompose_bytes = float(n) * 32 + entries * 2
    else:                 # table layout: staged two-level partition
        sort_bytes = entries * (4 + 6) + entries * (2 + 6 + 4)  # level 1: digit in, value + key out; level 2: key (hist), key + value in, entry out
        decompose_bytes = float(n) * 32 + entries * 4
    sort_ms = stages["bin_scatter"] + stages["bin_sort"]
    traffic = load_ncu_traffic()
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
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture, scaled to this launch's entry count
        "traffic": (traffic["k_accumulate"]["bytes"] * entries / traffic["k_accumulate"]["entries"]) if traffic and "k_accumulate" in traffic else None,
        "traffic_source": traffic.get("source") if traffic else "no ncu capture found under profiles/",
        "algorithmic_bytes": entries * 68,
    }
    roofline_sort = {
        "kernel": "decompose + the two sort levels (table layout: k_decompose_b, k_scatter_staged_b, k_bucket_hist_b, k_bucket_scatter_staged_b; plain bases: k_decompose, k_scatter_bins, k_sort_bins)",
        "bound": "hbm", "achieved": (sort_bytes + decompose_bytes) / ((sort_ms + stages["decompose"]) * 1e-3) / 1e9,
        "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
        "frac": (sort_bytes + decompose_bytes) / ((sort_ms + stages["decompose"]) * 1e-3) / 1e9 / hbm_peak, "kernel_ms": sort_ms + stages["decompose"],
        "sort_only": {"achieved": sort_bytes / (sort_ms * 1e-3) / 1e9, "frac": sort_bytes / (sort_ms * 1e-3) / 1e9 / hbm_peak, "kernel_ms": sort_ms},
        "traffic": (traffic["sort"]["bytes"] * entries / traffic["sort"]["entries"]) if traffic and "sort" in traffic and not plan["idx_bits"] else None,
        "algorithmic_bytes": sort_bytes + decompose_bytes,
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "ms_step_min": min(step_ms), "ms_step_median": statistics.median(step_ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
        "parity_checked": True,
        "parity_how": "timed device-resident result, e2e (pageable) and e2e (pinned) results == the known-discrete-log answer of the synthetic bases "
                      "(oracle: two field sums + one scalar multiplication per rank, points added with the big-integer reference)",
        "config": {
            "workload": f"one BN254 G1 variable_base_msm of {world} x 2^{args.log_n} points (2^{args.log_n} per GPU), "
                        "uniform random Fr scalars, known-dlog bases (a+i*d)G, bases resident",
            "points_per_gpu": n, "window_bits": plan["window_bits"], "windows": plan["windows"],
            "bases": "plain affine array (one bucket set per window)" if step_bases is d_bases or args.plain_bases else
                     f"resident table of window multiples, {plan['windows']} x 64 B per point ({plan['windows'] * 64 * n / 1e9:.1f} GB), built once in {register_s:.2f} s (untimed)",
            "parallelism": f"point-sharded x{world}" + (", NCCL all_gather of 128-byte partials" if distributed else ""),
            "l2": "inputs (scalars 32 B + bases 64 B per point) exceed the 126 MB L2 at this size; no explicit flush",
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": 64 * world,
                "host_memory": "pageable (plain numpy, like a Rust Vec<Fr>): staged through the library's pinned ring",
                "staged_bytes_per_step": int(staged_per_step), "staging_rate_gbps_rank0": round(staging_rate, 1),
                "path": "plonkish_cuda_msm_bn254_g1 (C ABI, pageable host scalars, registered bases)" if not distributed
                        else "per rank: plonkish_cuda_msm_bn254_g1_host_partial (C ABI, pageable host scalars, registered bases) -> NCCL all_gather of partials -> fold -> host"},
        "e2e_pinned": {"value": total_n / (e2e_pin_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_pin_ms, "host_memory": "pinned (cudaHostAlloc)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_sort": roofline_sort,
        "stages_ms": stages,
        "integer_pipe": pipe,
        "fp64_pipe": fp64,
        "madd_streams": madd_streams,
    }
    if strong:
        line["strong_2p%d" % args.log_n] = strong

    if rank == 0 and not args.no_cpu_baseline and not distributed:
        cores = po.host_threads()
        log_c = min(args.log_n, args.cpu_log_n)
        m = 1 << log_c
        sc = scalars_np[:m]
        bs = d_bases[:m].cpu().numpy().view(np.uint64)
        t0 = time.perf_counter()
        want_cpu = po.variable_base_msm(sc, bs, cores)
        sec = time.perf_counter() - t0
        got = result_dev if m == n else pk.variable_base_msm(sc, reg)
        assert (got == want_cpu).all(), "GPU result differs from the CPU oracle on the cpu_baseline sample"
        line["cpu_baseline"] = {
            "value": m / sec / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{'all' if m == n else 'first'} 2^{log_c} points of the step's inputs, one pass, C port of msm.rs:84-181 with {cores} pthreads; "
                      "the GPU's timed result on the same points checked bit-exact against it",
            "seconds": sec,
        }
    if rank == 0 and not distributed and step_bases is reg:
        # the same MSM on the plain affine bases (no table of window multiples: what an unregistered caller gets,
        # one bucket set per window, c = 16), beside the headline that uses 12 x 64 B of table per point
        reg_plain = pk.G1Bases(d_bases, mode=pk.G1Bases.PLAIN)
        out_p, ms_p, step_ms_p = timed_device(lambda: pk.variable_base_msm_device(d_scalars, reg_plain), args.steps, 3)
        assert out_p.cpu().numpy().view(np.uint64).tobytes() == want, "the plain-bases result differs from the known-dlog answer"
        plan_p = pk.msm_plan(n, 0, local_rank, bases=reg_plain)
        line["plain_bases"] = {"value": n / (ms_p * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_p, "ms_step_min": min(step_ms_p),
                               "window_bits": plan_p["window_bits"], "windows": plan_p["windows"],
                               "bases": "resident plain affine array, 64 B per point", "parity_checked": True}
        reg_plain.release()
    if rank == 0 and not distributed and not args.no_skew and step_bases is reg:
        line["skew"] = skew_bench(pk, torch, np, d_scalars, reg, ms_per_step, n, dev)
    if not distributed and args.prove_k:
        del d_scalars, d_bases
        reg.release()
        torch.cuda.empty_cache()
        cpu = not args.no_cpu_baseline
        seq = {"what": "MSM calls of HyperPlonk::prove for vanilla_plonk (4 x 2^k + 2^(k-1) + ... + 1 points) against an SRS built on the device; "
                       "gpu_ms: pageable host scalars per call; gpu_resident_ms: polynomials kept in HBM, g_prime merge and quotients on the GPU; "
                       "sum-check and the other field-only prover work stay in the Rust caller and are not included; min of `reps` runs"}
        seq["k%d" % args.prove_k] = prove_msm_sequence(pk, torch, np, args.prove_k, dev, cpu=False, reps=args.reps)
        if cpu:
            seq["k20"] = prove_msm_sequence(pk, torch, np, min(20, args.prove_k), dev, cpu=True, reps=args.reps)
        line["hyperplonk_prove_msm"] = seq
        line["univariate_kzg_k22"] = univariate_sequence(pk, torch, np, min(22, args.prove_k), dev, args.reps)
        line["srs_fixed_base_msm"] = srs_setup_bench(pk, torch, np, args.prove_k, cpu=cpu, reps=args.reps)
        line["sum_check_zero_check"] = sum_check_bench(pk, torch, np, args.prove_k, cpu=cpu, reps=args.reps)
        prove_leg = {"k%d" % args.prove_k: hyperplonk_prove_bench(pk, torch, np, args.prove_k, cpu=False, reps=args.reps)}
        if cpu:
            prove_leg["k20"] = hyperplonk_prove_bench(pk, torch, np, min(20, args.prove_k), cpu=True, reps=args.reps)
        line["hyperplonk_prove"] = prove_leg
        try:
            # own resident slices for the SRS prefixes the quotient / fold commitments run against (disclosed in the leg)
            line["pcs_schemes"] = pcs_schemes_bench(pk, torch, np, min(20, args.prove_k), reps=args.reps, prefix_tables=True)
        except Exception as e:  # noqa: BLE001 - keep the primary numbers if this leg cannot run
            line["pcs_schemes"] = {"error": f"{type(e).__name__}: {e}"}
    if distributed and not args.no_single_process and not args.plain_bases:
        # rank 0 alone drives all `world` GPUs through the C ABI's multi-GPU entry (one process, NCCL gather inside the
        # library); the other ranks free their memory and wait on a host-side (gloo) barrier so their GPUs stay idle
        if rank != 0:
            del d_scalars, d_bases
            reg.release()
            torch.cuda.empty_cache()
        dist.barrier(group=host_group)
        if rank == 0:
            try:
                line["single_process"] = single_process_leg(pk, torch, np, args, world, a, d, want, ms_per_step, e2e_ms, e2e_pin_ms)
            except Exception as e:  # noqa: BLE001 - keep the primary numbers if this leg cannot run
                line["single_process"] = {"error": f"{type(e).__name__}: {e}"}
        dist.barrier(group=host_group)
    if rank == 0:
        emit(line)
    if distributed:
        dist.destroy_process_group()


def single_process_leg(pk, torch, np, args, world: int, a: int, d: int, want: bytes, weak_ms: float, e2e_ms: float, e2e_pin_ms: float):
    """plonkish_cuda_msm_bn254_g1_multi over all GPUs from ONE process: the same world x 2^log_n-point MSM the ranks just
    did together (same seeds, so the same expected point), host scalars in, affine point out."""
    n = 1 << args.log_n
    total = n * world
    host = np.empty((total, 4), dtype=np.uint64)          # pageable
    for r in range(world):
        host[r * n:(r + 1) * n] = pk.random_scalars(n, seed=1000 + r)
    shards = [pk.synth_bases_device(n, a, d, device=torch.device("cuda", g), first=g * n) for g in range(world)]
    for g in range(world):
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    reg = pk.ShardedG1Bases.from_device(shards, total)
    register_s = time.perf_counter() - t0
    del shards

    def timed(src, steps):
        for _ in range(2):
            out = pk.variable_base_msm(src, reg)
        ts = []
        for _ in range(steps):
            t = time.perf_counter()
            out = pk.variable_base_msm(src, reg)
            ts.append((time.perf_counter() - t) * 1e3)
        return out, ts

    out, ts = timed(host, args.steps)
    assert out.tobytes() == want, "single-process multi-GPU result differs from the known-dlog answer"
    res = {"what": f"one process, {world} GPUs, plonkish_cuda_msm_bn254_g1_multi: one host thread per device, chunk-pipelined uploads, ncclAllGather of the partials",
           "points": total, "e2e_pageable_ms": statistics.mean(ts), "e2e_pageable_ms_min": min(ts), "e2e_pageable_value": total / (statistics.mean(ts) * 1e-3) / 1e6,
           "torchrun_e2e_pageable_ms": e2e_ms, "torchrun_e2e_pinned_ms": e2e_pin_ms, "torchrun_device_ms": weak_ms,
           "register_s": register_s, "parity_checked": True}
    try:
        pin, keep = pinned_copy(torch, np, host)
        out, ts = timed(pin, args.steps)
        assert out.tobytes() == want
        res.update({"e2e_pinned_ms": statistics.mean(ts), "e2e_pinned_ms_min": min(ts), "e2e_pinned_value": total / (statistics.mean(ts) * 1e-3) / 1e6,
                    "vs_torchrun_pinned": statistics.mean(ts) / e2e_pin_ms})
        del pin, keep
    except RuntimeError as e:  # pinning world x 512 MB may be refused
        res["e2e_pinned_error"] = str(e)
    reg.release()
    return res


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest committed `ncu --set full` summary
    (profiles/*_traffic.json, written by tools/ncu_summary.py from the raw CSV page)."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None
    try:
        data = json.load(open(files[-1]))
        data["source"] = "profiles/" + os.path.basename(files[-1]) + " (ncu --set full capture of this round, not measured in this run)"
        return data
    except Exception:  # noqa: BLE001
        return None


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
