"""GPU parity suite, second part (-m gpu): the headline configuration's own code path, the pageable
upload path, the reference's timer lines, the borrowed-slice cache and handle lifetimes.  Bit-exact."""
import os
import re
import tempfile
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk():
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import plonkish_b200

    plonkish_b200._lib.lib()
    return plonkish_b200


def _pinned(torch, arr):
    t = torch.empty(arr.shape, dtype=torch.int64).pin_memory()
    h = t.numpy().view(np.uint64)
    h[:] = arr
    return t, h


def test_table_at_2pow24_known_discrete_log(pk, oracle):
    """BASELINE's metric size on the code path bench.py times: registered bases expanded into the table of window
    multiples (c = 22, 12 windows), which is the only size with 1 024-thread sort stages (tile >= 16 384 points,
    msm_kernels.cuh pk_make_plan_b), device-resident scalars AND host scalars (three upload chunks) from pinned and
    from pageable memory, all against the known-discrete-log answer."""
    import torch

    n = 1 << 24
    sc = pk.random_scalars(n, seed=2424)
    want = oracle.known_dlog_answer(3, 5, sc)
    d_bs = pk.synth_bases_device(n, 3, 5)
    reg = pk.G1Bases(d_bs, mode=pk.G1Bases.TABLE)
    plan = pk.msm_plan(n, 0, 0, bases=reg)
    assert plan["window_bits"] == 22 and plan["windows"] == 12 and plan["idx_bits"] == 0 and plan["tile"] >= 16384
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    got = pk.variable_base_msm_device(d_sc, reg).cpu().numpy().view(np.uint64)
    assert got.tobytes() == want.tobytes(), "device-resident scalars, table layout"
    del d_sc
    keep, pinned = _pinned(torch, sc)
    assert pk.variable_base_msm(pinned, reg).tobytes() == want.tobytes(), "pinned host scalars, 3 chunks"
    staged0 = pk.staged_bytes()
    assert pk.variable_base_msm(sc, reg).tobytes() == want.tobytes(), "pageable host scalars through the staging ring"
    assert pk.staged_bytes() - staged0 == n * 32, "the pageable upload did not go through the staging ring"
    reg.release()


@pytest.mark.parametrize("n", [(1 << 15) + 1, (1 << 20) + 3, (1 << 22) + 5])
def test_pageable_pinned_and_unstaged_uploads_agree(pk, oracle, n):
    # a Rust Vec<Fr> is pageable (kzg.rs:255 passes poly.evals()); the staged path must give the same point as the
    # direct copies, for a registered slice and for per-call bases, single / batch / many entries
    import torch

    sc = pk.random_scalars(n, seed=n % 997)
    bs_dev = pk.synth_bases_device(n, 7, 11)
    want = oracle.known_dlog_answer(7, 11, sc)
    reg = pk.G1Bases(bs_dev)
    keep, pinned = _pinned(torch, sc)
    before = pk.staged_bytes()
    assert pk.variable_base_msm(pinned, reg).tobytes() == want.tobytes()
    assert pk.staged_bytes() == before, "a pinned source must not be staged"
    assert pk.variable_base_msm(sc, reg).tobytes() == want.tobytes()
    # upload chunks of 4 MiB and more go through the ring, smaller ones down the driver's pageable path
    assert (pk.staged_bytes() - before > 0) == (n >= 1 << 20)
    os.environ["PLONKISH_CUDA_COPY_THREADS"] = "3"  # (read once, at the first staged upload: harmless later)
    outs = pk.variable_base_msm_batch([sc, pinned, sc], reg)
    assert all(o.tobytes() == want.tobytes() for o in outs)
    half = n // 2
    outs = pk.variable_base_msm_many([sc, sc[:half]], [reg, reg])
    assert outs[0].tobytes() == want.tobytes()
    assert outs[1].tobytes() == oracle.known_dlog_answer(7, 11, sc[:half]).tobytes()
    if n <= 1 << 20:
        bs = bs_dev.cpu().numpy().view(np.uint64)
        assert pk.variable_base_msm(sc, bs).tobytes() == want.tobytes(), "pageable scalars and pageable bases"
    reg.release()


# ---- the reference's timer lines -------------------------------------------------------------------------------
from plotter_parse import capture_fd2 as _capture_fd2, plotter_parse as _plotter_parse  # noqa: E402


def test_timer_lines_parse_like_the_plotter(pk, oracle):
    """msm.rs:92 wraps every MSM in start_timer("variable_base_msm-{n}"); plotter.rs:142-153 sums those lines into its
    "multiexp" / "pcs multiexp" buckets and subtracts them from the enclosing timer.  One line per MSM, each MSM's
    own n, and the lines of one call must not add up to more than the call took."""
    from plonkish_b200 import kzg

    k = 12
    n = 1 << k
    g = oracle.generator()
    pp = kzg.setup(g, pk.random_scalars(k, seed=5))
    polys = [pk.random_scalars(n, seed=60 + j) for j in range(3)]
    point = pk.random_scalars(k, seed=6)
    pk.timer_config(1, 0)
    try:
        def run():
            t0 = time.perf_counter()
            single = pk.variable_base_msm(polys[0], pp.eqs[k])
            t1 = time.perf_counter()
            batch = pk.variable_base_msm_batch(polys, pp.eqs[k])
            t2 = time.perf_counter()
            many = pk.variable_base_msm_many([polys[0][: 1 << i] for i in range(k)], [pp.eqs[i] for i in range(k)])
            t3 = time.perf_counter()
            res = pk.ResidentScalars(polys[1])
            t4 = time.perf_counter()
            opened = kzg.open_resident(pp, res, point)
            t5 = time.perf_counter()
            res.release()
            return (single, batch, many, opened), [(t1 - t0), (t2 - t1), (t3 - t2), (t5 - t4)]

        _capture_fd2(run)  # warm-up: scratch buffers grow on first use
        ((single, batch, many, opened), walls), text = _capture_fd2(run)
    finally:
        pk.timer_config(0, 0)
    logs = _plotter_parse(text)
    names = [l["name"] for l in logs]
    sizes = [f"variable_base_msm-{1 << i}" for i in range(k)]
    assert names == [f"variable_base_msm-{n}"] + [f"variable_base_msm-{n}"] * 3 + sizes + sizes, names
    assert all(l["depth"] == 0 and not l["children"] and l["ns"] > 0 for l in logs)
    groups = [logs[0:1], logs[1:4], logs[4:4 + k], logs[4 + k:]]
    for grp, wall in zip(groups, walls):
        total = sum(l["ns"] for l in grp) * 1e-9
        assert total <= wall * 1.02 + 2e-4, (total, wall)      # never more than the call (the plotter subtracts)
        assert total >= wall * 0.5, (total, wall)               # and it accounts for the call, not for a sliver of it
    # nested depth: two levels of the caller's timers are open
    pk.timer_config(1, 2)
    try:
        _, text2 = _capture_fd2(lambda: pk.variable_base_msm(polys[0], pp.eqs[k]))
    finally:
        pk.timer_config(0, 0)
    wrapped = "Start:   outer\n··Start:   inner\n" + text2 + "··End:     inner ....1.000ms\nEnd:     outer ....2.000ms\n"
    top = _plotter_parse(wrapped)
    assert len(top) == 1 and top[0]["children"][0]["children"][0]["name"] == f"variable_base_msm-{n}"
    assert single.tobytes() == batch[0].tobytes() == oracle.variable_base_msm(polys[0], pp.eqs[k].to_host()).tobytes()
    pp.release()


# ---- the borrowed-slice cache ------------------------------------------------------------------------------------
def test_cached_bases_follow_the_content_not_the_address(pk, oracle):
    """The shim inside variable_base_msm only has a borrowed slice.  A freed Vec's address is commonly reused by the next
    Vec of the same length (a second setup with another `s`, IPA's folded generators): the cache must notice."""
    n = 1 << 11
    pk.cache_limit(0)
    buf = np.ascontiguousarray(oracle.known_dlog_bases(3, 5, n))
    sc = oracle.random_scalars(n, 1)
    h1 = pk.cached_bases(buf)
    assert pk.variable_base_msm(sc, h1).tobytes() == oracle.known_dlog_answer(3, 5, sc).tobytes()
    h1b = pk.cached_bases(buf)
    assert h1b.handle == h1.handle, "an unchanged slice must hit"
    # prefix of the cached slice (&powers_of_s_g1[..len], univariate/kzg.rs:28): same entry
    hp = pk.cached_bases(buf[: n // 2])
    assert hp.handle == h1.handle
    assert pk.variable_base_msm(sc[: n // 2], hp).tobytes() == oracle.known_dlog_answer(3, 5, sc[: n // 2]).tobytes()
    # "free + malloc returned the same address": other bases in the same memory
    buf[:] = oracle.known_dlog_bases(11, 13, n)
    h2 = pk.cached_bases(buf)
    assert h2.handle != h1.handle, "stale handle returned for reused memory"
    assert pk.variable_base_msm(sc, h2).tobytes() == oracle.known_dlog_answer(11, 13, sc).tobytes()
    with pytest.raises(pk.PlonkishCudaError):
        pk.variable_base_msm(sc, h1)  # the stale entry was released
    # only the tail differs beyond a prefix request: the prefix's last point is compared with the device copy
    buf[n // 2 - 1] = oracle.known_dlog_bases(2, 1, 1)[0]
    hq = pk.cached_bases(buf[: n // 2])
    assert hq.handle != h2.handle
    # a longer slice at the same address replaces the entry
    big = np.ascontiguousarray(oracle.known_dlog_bases(3, 5, 2 * n))
    ha = pk.cached_bases(big[:n])
    hb = pk.cached_bases(big)
    assert hb.handle != ha.handle and pk.cache_stats()["entries"] == 2
    sc2 = oracle.random_scalars(2 * n, 2)
    assert pk.variable_base_msm(sc2, hb).tobytes() == oracle.known_dlog_answer(3, 5, sc2).tobytes()
    # LRU bound: a limit below one table evicts everything but the newest use
    stats = pk.cache_stats()
    pk.cache_limit(max(1, stats["bytes"] // 3))
    assert pk.cache_stats()["bytes"] <= max(1, stats["bytes"] // 3) or pk.cache_stats()["entries"] <= 1
    pk.cache_limit(0)
    pk.cache_evict(big)
    pk.cache_evict(buf)
    assert pk.cache_stats()["entries"] == 0


def test_short_prefix_of_a_long_table_uses_the_plain_layout(pk, oracle):
    n = 1 << 16
    d_bs = pk.synth_bases_device(n, 3, 5)
    reg = pk.G1Bases(d_bs, mode=pk.G1Bases.TABLE)
    assert pk.msm_plan(n, 0, 0, bases=reg)["idx_bits"] == 0          # table layout
    assert pk.msm_plan(1 << 8, 0, 0, bases=reg)["idx_bits"] != 0     # 1/256 of the slice: plain layout on row 0
    for m in (1 << 8, 1000, (1 << 12) + 1, n // 2):
        sc = oracle.random_scalars(m, m)
        assert pk.variable_base_msm(sc, reg).tobytes() == oracle.known_dlog_answer(3, 5, sc).tobytes(), m
    reg.release()


def test_release_while_another_thread_is_using_the_handle(pk, oracle):
    """Rust's Drop may release a slice or a polynomial on one thread while a rayon worker's call on it is still in
    flight (hyrax.rs:176-180): handles are reference counted, the call must finish with the right point."""
    import torch

    n = 1 << 20
    sc = pk.random_scalars(n, seed=99)
    want = oracle.known_dlog_answer(3, 5, sc)
    keep, pinned = _pinned(torch, sc)
    for _ in range(3):
        d_bs = pk.synth_bases_device(n, 3, 5)
        torch.cuda.synchronize()
        reg = pk.G1Bases(d_bs, mode=pk.G1Bases.TABLE)
        del d_bs
        out = {}

        def work():
            out["points"] = []
            try:
                for _ in range(4):
                    out["points"].append(pk.variable_base_msm(pinned, reg))
            except pk.PlonkishCudaError as e:  # a later call finds the handle gone
                out["error"] = str(e)

        t = threading.Thread(target=work)
        t.start()
        time.sleep(0.004)
        handle = reg.handle
        # (through the C ABI directly: the worker keeps presenting the same handle, as a Rust caller holding a copy would)
        from plonkish_b200 import _lib

        _lib.check(_lib.lib().plonkish_cuda_bases_release(handle), "plonkish_cuda_bases_release")
        # memory pressure on the freed table: a new registration may land on the same addresses
        other = pk.G1Bases(pk.synth_bases_device(n, 9, 2), mode=pk.G1Bases.TABLE)
        t.join()
        other.release()
        reg.handle = 0
        for p in out.get("points", []):
            assert p.tobytes() == want.tobytes()
        if "error" in out:
            assert "unknown bases handle" in out["error"]
        assert handle


def test_resident_polynomial_released_under_a_running_sum_check(pk, oracle):
    from plonkish_b200 import sumcheck

    k = 10
    tabs_h = [pk.random_scalars(1 << k, seed=700 + i) for i in range(3)]
    tabs = [pk.ResidentScalars(t) for t in tabs_h]
    one = sumcheck._to_mont(1)
    terms = [(one, [0, 1]), (one, [2])]
    prover = sumcheck.SumCheckProver(tabs, terms)
    for t in tabs:
        t.release()                       # Drop before the first round
    junk = [pk.ResidentScalars(pk.random_scalars(1 << k, seed=800 + i)) for i in range(6)]  # reuses pool memory if it was freed
    got = prover.round_evals()
    assert got.tobytes() == oracle.sumcheck_round(tabs_h, terms).tobytes()
    prover.free()
    for j in junk:
        j.release()


def test_multi_entry_on_the_visible_gpus(pk, oracle):
    # one process, G devices, NCCL gather (G = 1 exercises the same entry without a communicator)
    from plonkish_b200 import _lib

    gpus = _lib.lib().plonkish_cuda_device_count()
    n = (1 << 20) + 77
    sc = pk.random_scalars(n, seed=5)
    bs = pk.synth_bases_device(n, 3, 5).cpu().numpy().view(np.uint64)
    want = oracle.known_dlog_answer(3, 5, sc)
    for g in sorted({1, min(2, gpus), gpus}):
        reg = pk.ShardedG1Bases(bs, g)
        assert pk.variable_base_msm(sc, reg).tobytes() == want.tobytes(), g
        reg.release()
        assert pk.variable_base_msm(sc[:5000], bs[:5000], n_gpus=g).tobytes() == oracle.known_dlog_answer(3, 5, sc[:5000]).tobytes(), g


def _skewed(pk, n, kind, seed):
    """Scalar sets of SURVEY.md 8d: a 0 / 1 / -1 selector column with half zeros, small integers, one value everywhere."""
    from oracle import bigint_ref as br

    rng = np.random.default_rng(seed)
    r = br.R if hasattr(br, "R") else 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    mont = lambda v: np.frombuffer((v % r * (1 << 256) % r).to_bytes(32, "little"), dtype=np.uint64)  # noqa: E731
    if kind == "selector":
        pool = np.stack([mont(0), mont(1), mont(r - 1)])
        return pool[rng.choice(3, size=n, p=[0.5, 0.25, 0.25])].copy()
    if kind == "small":
        pool = np.stack([mont(v) for v in range(64)])
        return pool[rng.integers(0, 64, size=n)].copy()
    return np.repeat(pk.random_scalars(1, seed)[:1], n, axis=0).copy()


@pytest.mark.parametrize("n", [70001, (1 << 19) + 3])
def test_accumulate_tiers_and_equal_runs_give_the_same_point(pk, oracle, n):
    """k_accumulate's runs (tiers of shrinking length derived from the entry count on the device, msm_kernels.cuh
    pk_acc_run) against equal runs and other tier counts, on uniform scalars and on the skew set — whose entry counts are
    a fraction of what the launch is sized for — all against the known-discrete-log answer."""
    import torch

    d_bs = pk.synth_bases_device(n, 3, 5)
    reg = pk.G1Bases(d_bs, mode=pk.G1Bases.TABLE)
    try:
        for kind in ("uniform", "selector", "small", "same"):
            sc = pk.random_scalars(n, seed=n % 991) if kind == "uniform" else _skewed(pk, n, kind, n % 991)
            want = oracle.known_dlog_answer(3, 5, sc).tobytes()
            d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
            for knob, value in ((None, None), ("PLONKISH_CUDA_ACC_TIERS", "1"), ("PLONKISH_CUDA_ACC_TIERS", "3"), ("PLONKISH_CUDA_ACC_L", "64"),
                                ("PLONKISH_CUDA_ACC_WAVES", "2.5")):
                if knob:
                    os.environ[knob] = value
                try:
                    got = pk.variable_base_msm_device(d_sc, reg).cpu().numpy().view(np.uint64).tobytes()
                finally:
                    if knob:
                        os.environ.pop(knob)
                assert got == want, (kind, knob, value)
    finally:
        reg.release()


def test_five_chunk_upload_and_the_staging_rate(pk, oracle):
    """The chunk geometry the host-scalar MSM switches to when the staging ring is slow (ranks sharing a host), forced
    here through PLONKISH_CUDA_HOST_CUTS, from pageable and from pinned memory; and the ring reports the rate of its
    last large uploads."""
    import torch

    n = (1 << 23) + 7
    sc = pk.random_scalars(n, seed=523)
    want = oracle.known_dlog_answer(3, 5, sc).tobytes()
    reg = pk.G1Bases(pk.synth_bases_device(n, 3, 5), mode=pk.G1Bases.TABLE)
    keep, pinned = _pinned(torch, sc)
    try:
        assert pk.variable_base_msm(sc, reg).tobytes() == want
        assert pk.staging_rate_gbps() > 0.5, "a 256 MiB pageable upload went through the ring: it must have measured its rate"
        os.environ["PLONKISH_CUDA_HOST_CUTS"] = "0.0625,0.16,0.32,0.58,1"
        try:
            assert pk.variable_base_msm(sc, reg).tobytes() == want, "five chunks, pageable"
            assert pk.variable_base_msm(pinned, reg).tobytes() == want, "five chunks, pinned"
        finally:
            os.environ.pop("PLONKISH_CUDA_HOST_CUTS")
    finally:
        reg.release()
