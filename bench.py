#!/usr/bin/env python3
"""bench.py — headline benchmark of the SDE path-simulation hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU algorithm (oracle port)

Workload (BASELINE.json configs[1], "C2"): 1-D GBM, Euler–Maruyama, RQMC scrambled Sobol,
2^24 paths x 252 steps per GPU, full-path f64 output [N, 253, 1] (33.96 GB) left resident in HBM.
A "step" is one pass of the hot path over that batch.  Multi-GPU: one process per GPU (torchrun),
rank r simulates scenarios [r*2^24, (r+1)*2^24) — disjoint Sobol index ranges, no data-path collective —
so scaling is weak and `value` is the aggregate path-steps/s.

The JSON line carries `roofline` (HBM write bound; algorithmic bytes = 8*P per path-step incl. the t0 row),
`cpu_baseline` (the CPU oracle on a bounded sample, on this box's host cores), `e2e` (same workload through
the C-ABI host-buffer call: chunked simulate + D2H into pinned host memory inside the timed region, with its own
roofline against the measured pinned D2H rate), `clocks`, `gpu_launches`, and beside the headline:
`timed_output_parity` (rows of the timed output against the CPU oracle), `plan_create_ms` (cached / cold NVRTC),
`configs` (every other BASELINE.json config on this GPU, a few launches each: C1, C2 with the reference's cp_shift
scramble, C2 in [T][P][N] layout, C3 full paths and terminal, C4 moments, C5 per GPU) and `c5_strong` (BASELINE C5:
2^33 paths x 365 steps split over the N GPUs, the NCCL all-gather + device Chan merge of the moments inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sde-sim-rs_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GBM_EQ = ["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"]
D = 252
TIMES = [k / D for k in range(D + 1)]
INIT = {"X1": 1.0}
N_PER_GPU = 1 << 24
SEED = 42
METRIC = "path_steps_per_sec"
UNIT = "path-steps/s"
WORKLOAD = "C2: 1-D GBM Euler-Maruyama, RQMC scrambled Sobol (XOR digital shift), 2^24 paths x 252 steps per GPU, full-path f64 output [N,253,1]"


HESTON_EQ = ["dS = ( 0.05 * S ) * dt + ( max(v, 0.0)^0.5 * S ) * dW1",
             "dv = ( 2.0 * (0.04 - v) ) * dt + ( -0.21 * max(v, 0.0)^0.5 ) * dW1 + ( 0.2142428528562855 * max(v, 0.0)^0.5 ) * dW2"]


def basket_equations(n_assets=64, rho=0.5):
    """C4: correlated GBM basket, correlation through shared dW names with Cholesky loadings."""
    import numpy as np

    corr = np.full((n_assets, n_assets), rho)
    np.fill_diagonal(corr, 1.0)
    L = np.linalg.cholesky(corr)
    eqs = []
    for i in range(n_assets):
        sig = 0.1 + 0.2 * i / max(n_assets - 1, 1)
        eqs.append(f"dS{i} = ( 0.05 * S{i} ) * dt + " + " + ".join(f"( {sig * L[i, j]:.17g} * S{i} ) * dW{j + 1}" for j in range(i + 1)))
    return eqs, {f"S{i}": 100.0 for i in range(n_assets)}


def _measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic_bytes():
    """dram bytes per launch of the fused kernel from the committed ncu --set full capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw.instant,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_ready(self, timeout_s: float = 2.0):
        """nvidia-smi needs 0.1-0.5 s before its first sample: wait for it so the timed region is covered from its start."""
        t_end = time.perf_counter() + timeout_s
        while not self.rows and self.proc is not None and time.perf_counter() < t_end:
            time.sleep(0.01)

    def mark_begin(self):
        """Samples that arrived before this call (warm-up, nvidia-smi start-up) are not part of the timed region."""
        self.t_begin = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        t_begin = getattr(self, "t_begin", 0.0)
        rows = [r for (t, r) in list(self.rows) if t_begin <= t <= t_end + 0.05]   # the reader thread may still append
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons, pw = [], 0, set(), []
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "sm_mhz_min": sm[0] if sm else None, "power_w_max": max(pw) if pw else None}


def _host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm must not inherit that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_reference_leg(sample_paths: int, repeats: int = 1):
    """The reference's CPU algorithm (oracle port, OpenMP over scenarios like rayon in src/sim/mod.rs:41-43)
    on a bounded sample of the C2 workload.  Returns (path_steps_per_s, cores, seconds_per_pass)."""
    from oracle import oracle as orc

    orc.build()
    U = orc.Universe(GBM_EQ, TIMES)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.simulate(U, INIT, sample_paths, "euler", "sobol", seed=SEED, scramble="xor", nthreads=_host_threads())
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample_paths * D / best, _host_threads(), best



def timed_output_parity(out, n, offset):
    """Rows of the buffer the timed region just wrote against the CPU oracle (the checker, not the thing measured):
    the first 256, 256 in the middle and the last 256 scenarios; fast tier => <= 1e-11 relative."""
    import numpy as np

    from oracle import oracle as orc

    orc.build()
    U = orc.Universe(GBM_EQ, TIMES)
    worst, rows = 0.0, 0
    for lo in (0, n // 2 + 77, n - 256):
        lo = max(0, min(lo, n - 256))
        cnt = min(256, n - lo)
        ref = orc.simulate(U, INIT, cnt, "euler", "sobol", seed=SEED, scramble="xor", scenario_offset=offset + lo, nthreads=2)
        got = out[lo:lo + cnt].cpu().numpy()
        worst = max(worst, float(np.max(np.abs(got - ref) / np.abs(ref))))
        rows += cnt
    return {"max_rel_err": worst, "rows_checked": rows, "tolerance": 1e-11, "ok": bool(worst <= 1e-11),
            "against": "CPU oracle (oracle/sde_oracle.cpp) on the same scenario indices"}


def _time_plan(torch, plan, init, n, out, reps, **kw):
    plan.run(init, n, seed=SEED, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run(init, n, seed=SEED, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def plan_create_times(S, dev):
    """Plan creation: from the ahead-of-time cubin cache (what a deployment sees) and cold (NVRTC sm_100a compile of a model the
    cache has never seen: the headline model with a perturbed coefficient)."""
    t0 = time.perf_counter()
    p = S.Plan(S.Universe(GBM_EQ, TIMES), "euler", "sobol", output="paths", scramble="xor", icdf="fast", arithmetic="fast", device=dev)
    cached_ms, was_cached = (time.perf_counter() - t0) * 1e3, p.prelowered
    eq = [f"dX1 = ( 0.05 * X1 ) * dt + ( 0.1{int(time.time() * 1e3) % 100000:05d} * X1) * dW1"]
    t0 = time.perf_counter()
    S.Plan(S.Universe(eq, TIMES), "euler", "sobol", output="paths", scramble="xor", icdf="fast", arithmetic="fast", device=dev)
    return {"from_cache_ms": cached_ms, "cache_hit": bool(was_cached), "cold_nvrtc_ms": (time.perf_counter() - t0) * 1e3,
            "what": "sde_plan_create of the C2 plan: lowering + cubin (disk cache / NVRTC --gpu-architecture=sm_100a) + table upload"}


def config_legs(S, torch, dev, holder, peak_gbs, dfma_tflops):
    """Every other BASELINE.json config on this GPU: one warm launch + a few timed ones each (CUDA events, device-resident
    output).  Full-path configs against the HBM write roofline (8 P bytes per path-step, t0 row included); terminal /
    moment configs against the measured DFMA peak with the flop model of SURVEY.md 8(d).  `holder[0]` is the 34 GB
    buffer of the headline leg: reused by the 2^24 x 253 legs, released before the 67 GB of C3's full paths."""
    legs = {}
    fast = dict(icdf="fast", arithmetic="fast")

    def hbm(n, t_len, p_, ms):
        gbs = n * t_len * p_ * 8 / (ms * 1e-3) * 1e-9
        return {"bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs}

    def fp64(rate, flop):
        tf = rate * flop * 1e-12
        return {"bound": "fp64", "achieved": tf, "peak": dfma_tflops, "unit": "TFLOP/s", "frac": (tf / dfma_tflops) if dfma_tflops else None,
                "flop_model_per_path_step": flop}

    def leg(name, eqs, times, init, n, scheme, rng, reps, roof, reuse=False, **kw):
        plan = S.Plan(S.Universe(eqs, times), scheme, rng, device=dev, **kw)
        shape = plan.output_shape(n)
        numel = 1
        for d_ in shape:
            numel *= d_
        o = holder[0].view(-1)[:numel].view(shape) if reuse else torch.empty(shape, dtype=torch.float64, device=f"cuda:{dev}")
        ms = _time_plan(torch, plan, init, n, o, reps)
        steps = len(times) - 1
        rate = n * steps / (ms * 1e-3)
        legs[name] = {"value": rate, "unit": UNIT, "ms": ms, "paths": n, "steps": steps, "launches_timed": reps, "roofline": roof(n, steps, ms, rate)}
        del o, plan
        torch.cuda.empty_cache()

    g252, g365, g1000 = TIMES, [k / 365 for k in range(366)], [k / 1000 for k in range(1001)]
    n2 = 1 << 24
    heston_init = {"S": 100.0, "v": 0.04}
    leg("C1_gbm_euler_pseudo_10k_full_paths", GBM_EQ, g252, INIT, 10_000, "euler", "pseudo", 5, lambda n, s_, ms, r: hbm(n, s_ + 1, 1, ms), output="paths", **fast)
    leg("C2_cp_shift_per_path_full_paths", GBM_EQ, g252, INIT, n2, "euler", "sobol", 3, lambda n, s_, ms, r: hbm(n, s_ + 1, 1, ms), reuse=True,
        output="paths", scramble="cp_shift_per_path", **fast)
    leg("C2_layout_TPN_full_paths", GBM_EQ, g252, INIT, n2, "euler", "sobol", 3, lambda n, s_, ms, r: hbm(n, s_ + 1, 1, ms), reuse=True,
        output="paths", layout="TPN", scramble="xor", **fast)
    holder[0] = None
    torch.cuda.empty_cache()
    leg("C3_heston_rk_sobol_4M_x_1000_full_paths", HESTON_EQ, g1000, heston_init, 1 << 22, "runge-kutta", "sobol", 2,
        lambda n, s_, ms, r: hbm(n, s_ + 1, 2, ms), output="paths", scramble="xor", **fast)
    leg("C3_heston_rk_sobol_4M_x_1000_terminal", HESTON_EQ, g1000, heston_init, 1 << 22, "runge-kutta", "sobol", 2,
        lambda n, s_, ms, r: fp64(r, 100), output="terminal", scramble="xor", **fast)
    beq, binit = basket_equations(64)
    leg("C4_basket64_euler_sobol_1M_x_252_moments", beq, g252, binit, 1 << 20, "euler", "sobol", 2,
        lambda n, s_, ms, r: fp64(r, 5632), output="moments", scramble="xor", **fast)
    leg("C5_gbm_euler_pseudo_2p30_x_365_moments_per_gpu", GBM_EQ, g365, INIT, 1 << 30, "euler", "pseudo", 1,
        lambda n, s_, ms, r: fp64(r, 26), output="moments", generator="philox", **fast)
    leg("C5_same_with_the_reference_chacha8_stream", GBM_EQ, g365, INIT, 1 << 30, "euler", "pseudo", 1,
        lambda n, s_, ms, r: fp64(r, 26), output="moments", **fast)
    return legs


def c5_strong_leg(S, torch, dist, world, rank, dev, total_paths):
    """BASELINE C5: `total_paths` x 365 steps of GBM terminal-only MC split over the `world` GPUs (strong scaling); every rank
    reduces its shard to (count, mean, M2), the triples are all-gathered over NCCL and Chan-merged by the library's
    device kernel — all inside the timed region (CUDA events, max over ranks).  collective_ms: the all-gather + merge
    alone, timed on its own afterwards."""
    g365 = [k / 365 for k in range(366)]
    kw = dict(seed=2024, output="moments", icdf="fast", arithmetic="fast", device=dev, generator="philox")
    S.simulate_sharded(GBM_EQ, g365, min(total_paths, world << 20), INIT, "pseudo", "euler", **kw)      # plan, communicator
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = S.simulate_sharded(GBM_EQ, g365, total_paths, INIT, "pseudo", "euler", **kw)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{dev}")
    coll_ms = 0.0
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        g = torch.empty((world,) + tuple(res.values.shape), dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_gather_into_tensor(g, res.values)
        torch.cuda.synchronize()
        dist.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            dist.all_gather_into_tensor(g, res.values)
            S.merge_moments_device(g)
        c1.record()
        torch.cuda.synchronize()
        coll_ms = c0.elapsed_time(c1) / 5
    ms = float(t.item())
    m = res.to_numpy()[0]
    mu, sig, dt = 0.05, 0.1, 1.0 / 365
    mean = (1 + mu * dt) ** 365
    var = ((1 + mu * dt) ** 2 + sig * sig * dt) ** 365 - mean**2
    return {"value": total_paths * 365 / (ms * 1e-3), "unit": UNIT, "ms": ms, "total_paths": total_paths, "steps": 365, "n_gpus": world,
            "scaling": "strong", "generator": "philox (Philox4x32-10: statistical parity with the reference's ChaCha8 stream, tests/test_gpu_philox.py)",
            "collective": "ncclAllGather of 3 doubles per rank + device Chan merge (sde_moments_merge_device), inside the timed region",
            "collective_ms": coll_ms, "count": float(m[0]), "mean": float(m[1]), "mean_closed_form": mean,
            "mean_err_in_standard_errors": abs(float(m[1]) - mean) / (var / total_paths) ** 0.5}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 1 << 20                                     # ~1 s of CPU work per step on 16 cores
    W, K = max(args.warmup, 0), max(args.steps, 1)
    from oracle import oracle as orc

    orc.build()
    U = orc.Universe(GBM_EQ, TIMES)
    cores = _host_threads()
    for _ in range(min(W, 1)):
        orc.simulate(U, INIT, sample, "euler", "sobol", seed=SEED, scramble="xor", nthreads=cores)
    K = min(K, 5)
    t0 = time.perf_counter()
    for _ in range(K):
        orc.simulate(U, INIT, sample, "euler", "sobol", seed=SEED, scramble="xor", nthreads=cores)
    dt = (time.perf_counter() - t0) / K
    value = sample * D / dt
    desc = f"{sample} paths x {D} steps of C2 per step (oracle C++ port of the reference algorithm, OpenMP, {cores} threads)"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": min(W, 1),
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import sde_sim_rs as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION: keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    N, S_, P, T = args.paths, D, 1, D + 1
    offset = rank * N

    plan_ms = plan_create_times(S, dev) if rank == 0 else None
    plan = S.Plan(S.Universe(GBM_EQ, TIMES), "euler", "sobol", output="paths", layout="NTP", scramble="xor",
                  icdf=args.icdf, arithmetic=args.arithmetic, device=dev, tile_steps=args.tile_steps, block_threads=args.block)
    out = torch.empty((N, T, P), dtype=torch.float64, device=f"cuda:{dev}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: `value`
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for _ in range(args.warmup):
        plan.run(INIT, N, seed=SEED, scenario_offset=offset, out=out)
    barrier()
    launches0 = plan.launches
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.run(INIT, N, seed=SEED, scenario_offset=offset, out=out)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    gpu_launches = plan.launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * N * S_ / (ms_per_step * 1e-3)

    # the same launch on its own (informational, not the headline): 5 launches 0.3 s apart, each timed separately, so the
    # board power controller is idle when it starts — the gap to `ms_per_step` is the sw_power_cap clock drop under
    # back-to-back launches
    isolated_ms = None
    if rank == 0:
        singles = []
        for _ in range(5):
            time.sleep(0.3)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            plan.run(INIT, N, seed=SEED, scenario_offset=offset, out=out)
            b.record()
            torch.cuda.synchronize()
            singles.append(a.elapsed_time(b))
        isolated_ms = sorted(singles)[len(singles) // 2]
    if world > 1:
        dist.barrier()

    # parity of what was just timed: the buffer the last timed launch wrote, against the CPU oracle on the same scenarios
    parity = None
    if rank == 0:
        plan.run(INIT, N, seed=SEED, scenario_offset=offset, out=out)       # (the isolated launches above wrote the same values)
        torch.cuda.synchronize()
        parity = timed_output_parity(out, N, offset)
        assert parity["ok"], parity
    if world > 1:
        dist.barrier()

    # ---- the other BASELINE configs on this GPU (rank 0; the other ranks wait) and C5 strong-scaled over all ranks
    configs, peaks = None, None
    holder = [out]
    del out
    if not args.no_configs:
        if rank == 0:
            import ctypes as C

            fill, dfma, ffma = C.c_double(0), C.c_double(0), C.c_double(0)
            S._ffi.check(S._ffi.lib().sde_measure_peaks(dev, C.byref(fill), C.byref(dfma), C.byref(ffma)))
            peaks = {"fill_gbs": fill.value, "dfma_tflops": dfma.value, "ffma_tflops": ffma.value,
                     "how": "sde_measure_peaks: 8 GiB pure-write fill; 8 independent FMA chains per thread (own kernels, this box, this run)"}
            configs = config_legs(S, torch, dev, holder, _measured_peak_gbs()[0], dfma.value)
        if world > 1:
            dist.barrier()
    holder[0] = None
    torch.cuda.empty_cache()
    c5 = None
    if not args.no_configs:
        c5 = c5_strong_leg(S, torch, dist, world, rank, dev, args.c5_paths)

    # ---- end-to-end leg: host buffers through the C-ABI (sde_plan_run_host), D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        nbytes = N * T * P * 8
        try:
            host = torch.empty((N, T, P), dtype=torch.float64, pin_memory=True)
        except Exception as ex:  # noqa: BLE001
            host = None
            e2e = {"value": None, "unit": UNIT, "error": f"pinned host allocation of {nbytes} B failed: {ex}"}
        if host is not None:
            plan.run_host(INIT, N, seed=SEED, scenario_offset=offset, out=host)          # warm-up (also faults pages in)
            barrier()
            k2 = max(1, min(args.e2e_steps, args.steps))
            t0 = time.perf_counter()
            for _ in range(k2):
                plan.run_host(INIT, N, seed=SEED, scenario_offset=offset, out=host)      # synchronous
            barrier()
            dt = (time.perf_counter() - t0) / k2
            tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{dev}")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            assert float(host[0, 0, 0]) == 1.0 and bool(torch.isfinite(host[-1]).all())
            # the ceiling of this leg: pinned device -> host copies on this box (4 GiB blocks, all ranks at once like the leg itself)
            probe = torch.empty(1 << 29, dtype=torch.float64, device=f"cuda:{dev}")
            hview = host.view(-1)[: 1 << 29]
            hview.copy_(probe)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                hview.copy_(probe, non_blocking=True)
            barrier()
            d2h = torch.tensor([3 * (1 << 32) / (time.perf_counter() - t0) * 1e-9], dtype=torch.float64, device=f"cuda:{dev}")
            if world > 1:
                dist.all_reduce(d2h, op=dist.ReduceOp.MIN)
            d2h_gbs = float(d2h.item())
            del probe
            e2e = {"value": world * N * S_ / dt, "unit": UNIT, "h2d_bytes_per_step": 8 * P + 8 * S_ * 1,
                   "d2h_bytes_per_step": nbytes, "ms_per_step": dt * 1e3, "steps": k2,
                   "api": "sde_plan_run_host (C-ABI, pinned host output, 512 MiB chunks, copy overlapped with compute)",
                   "roofline": {"bound": "pinned D2H over PCIe", "achieved": nbytes / dt * 1e-9, "peak": d2h_gbs, "unit": "GB/s per GPU",
                                "frac": nbytes / dt * 1e-9 / d2h_gbs,
                                "how": f"peak = slowest rank's 4 GiB pinned cudaMemcpyAsync D2H rate with all {world} rank(s) copying at once"}}
            del host

    if rank == 0:
        peak, peak_src = _measured_peak_gbs()
        alg_bytes = N * T * P * 8
        achieved = alg_bytes / (ms_per_step * 1e-3) * 1e-9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": _ncu_traffic_bytes(), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel": "sde_sim_kernel",
                    "frac_of_8TBps_nominal": achieved / 8000.0,
                    "isolated_launch": {"ms": isolated_ms, "achieved": alg_bytes / (isolated_ms * 1e-3) * 1e-9,
                                        "frac": alg_bytes / (isolated_ms * 1e-3) * 1e-9 / peak,
                                        "how": "median of 5 launches 0.3 s apart, CUDA events around each (not part of the timed region)"}}
        cpu = None
        if not args.no_cpu:
            v, cores, secs = cpu_reference_leg(args.cpu_sample)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_sample} paths x {D} steps of C2, one pass ({secs:.1f} s), C++ oracle port, OpenMP"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "paths_per_gpu": N, "time_steps": S_, "processes": P, "seed": SEED,
                       "rng": "sobol/xor", "icdf": args.icdf, "arithmetic": args.arithmetic, "layout": "NTP",
                       "l2_policy": "each step writes 33.96 GB >> 126 MB L2 (no flush needed)",
                       "parallelism": f"paths sharded over {world} GPU(s), disjoint Sobol index ranges"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
            "timed_output_parity": parity, "plan_create_ms": plan_ms, "device_peaks": peaks, "configs": configs, "c5_strong": c5,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def claim_stdout():
    """stdout carries ONE JSON line.  Libraries write there too (NCCL prints its version banner on fd 1 at NCCL_DEBUG=WARN
    and VERSION), so fd 1 is pointed at stderr for the whole run and the line goes to a private copy of the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--paths", type=int, default=N_PER_GPU)
    ap.add_argument("--icdf", default="fast", choices=["fast", "reference"])
    ap.add_argument("--arithmetic", default="fast", choices=["fast", "strict"])
    ap.add_argument("--tile-steps", type=int, default=0)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs beside the headline (other BASELINE configs, C5 strong scaling)")
    ap.add_argument("--c5-paths", type=int, default=1 << 33, help="total paths of the C5 strong-scaling leg (BASELINE: 2^33)")
    ap.add_argument("--cpu-sample", type=int, default=1 << 21)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
