#!/usr/bin/env python
"""bench.py -- SAR Gibbs chain-iterations/s at N = 100k, 64 chains per B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA arm
    python bench.py --impl reference [...]                         CPU arm (oracle port of the R sampler)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A "step" is one block of ITERS Gibbs iterations for all chains resident on a GPU
(workload cfg4: synthetic SAR, N = 1e5, kNN(k=10) W, K = 20, 64 chains per GPU,
Chebyshev log-determinant).  Chains are independent, so GPUs shard chains and
never communicate inside the loop ("scaling": "weak", 64 chains per GPU).

Prints ONE JSON line.  `value` is device-resident throughput (inputs already in
HBM, draws left on the device); `e2e` is a complete fit through the C ABI from
pinned HOST buffers: create (H2D of y, X, W + setup incl. the Chebyshev trace
moments) + burn-in + sampling + D2H of the draws.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_OBS, K_REG, KNN, CHAINS_PER_GPU = 100_000, 20, 10, 64
CHEB_ORDER, CHEB_PROBES = 50, 64
ITERS_PER_STEP = 500
E2E_BURN, E2E_SAVE = 500, 1000           # defaults of bm(), R/60_interface.R:39
METRIC = "SAR Gibbs chain-iter/s at N=100k, 64 chains/GPU"
CFG5_N, CFG5_CHAINS = 1_000_000, 512      # BASELINE.json configs[4]: SEM, N = 1e6, 512 chains over 2/4/8 GPUs
CFG5_ITERS = 200


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the persistent kernel, from the committed ncu capture
    (profiles/ncu_traffic.json, stamped with the commit it was taken at); None when no capture is committed."""
    try:
        d = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        return int(d["dram_bytes_per_launch"]), d
    except Exception:
        return None, None


def load_peaks():
    try:
        d = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  `nvidia-smi -lms 20` runs as a process of its own (NVML inside
    this process crashed ncu when the bench was profiled); it needs ~0.2 s to start and the timed region of the default run is
    ~30 ms, so it is started before the warm-up steps and the rows are cut to the timed region by their timestamps (the last
    warm-up steps run the same kernel, so if fewer than two rows fall inside, the rows of the warm-up are kept as well and
    `window` says so)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        ok = [(t, r) for t, r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        inside = [r for t, r in ok if self.t0 is not None and self.t0 - 0.02 <= t <= (self.t1 or t) + 0.02]
        window = "timed region"
        if len(inside) < 2:
            inside, window = [r for _, r in ok], "warm-up + timed region (same kernel)"
        sm = [float(r[1]) for r in inside]
        mx = [float(r[2]) for r in inside if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i in range(4) if r[4 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window, "source": "nvidia-smi -lms 20"}


def make_cfg4(n=N_OBS):
    from bsreg_b200.synthetic import make_problem
    return make_problem("sar", n=n, k_reg=K_REG, knn=KNN)


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference sampler, one chain per worker process
# ----------------------------------------------------------------------------------------
_W = {}


def _cpu_worker(args):
    chain_id, iters = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle.sampler import NumpyRNG, RefChain
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    p, nodes = _W["p"], _W["nodes"]
    ch = RefChain("sar", p.y, p.X, W=p.W, ldet=dict(kind="nodes", re=nodes[0], im=nodes[1], w=nodes[2]),
                  rng=NumpyRNG(1000 + chain_id))
    t0 = time.perf_counter()
    for _ in range(iters):
        ch.sample()
    return time.perf_counter() - t0


def cpu_nodes(p):
    from oracle import core
    mom = core.trace_moments(p.W, 12, 8, 1)      # timing only: the node count (cost per log-det) is what matters
    mom = np.concatenate([mom, np.zeros(CHEB_ORDER - 12)])
    return core.chebyshev_nodes(mom, len(p.y))


def cpu_throughput(p, nodes, iters, procs):
    """chain-iter/s of `procs` independent oracle chains run in parallel processes."""
    import multiprocessing as mp
    _W["p"], _W["nodes"] = p, nodes
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        t0 = time.perf_counter()
        inner = pool.map(_cpu_worker, [(c, iters) for c in range(procs)])
        wall = time.perf_counter() - t0
    # steady-state rate: the slowest chain's own loop time (process start and chain setup excluded, as the GPU value
    # excludes bsreg_create); `wall` is reported beside it
    return procs * iters / max(inner), wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = make_cfg4()
    nodes = cpu_nodes(p)
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, CHAINS_PER_GPU))
    iters = args.ref_iters                # 300: about 2-3 s per step on every core used
    for _ in range(min(args.warmup, 1)):
        cpu_throughput(p, nodes, 20, procs)
    t0 = time.perf_counter()
    vals = [cpu_throughput(p, nodes, iters, procs)[0] for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = f"{procs} chains x {iters} iterations per step, one single-threaded process per chain"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "chain-iter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "chain-iter/s", "cores": procs, "kind": "port", "sample": sample,
                         "note": "reference-algorithm CPU restatement (NumPy/SciPy), not R; R is not installed"},
        "e2e": {"value": value, "unit": "chain-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_gpus):
    return {"workload": "cfg4: synthetic SAR, N=100000, kNN(k=10) row-stochastic W (nnz=1e6), K=20, "
                        "Chebyshev log-det (order 50, 64 probes)",
            "chains_per_gpu": CHAINS_PER_GPU, "global_chains": CHAINS_PER_GPU * n_gpus,
            "iters_per_step": ITERS_PER_STEP, "parallelism": f"chains sharded over {n_gpus} GPU(s), no collective",
            "stats_mode": "stream (W, y, X swept every iteration; persistent cooperative kernel, one launch per step)",
            "l2": "512 MiB L2 flush between timed steps; inside a step the 29 MB working set is L2-resident by "
                  "construction of the Gibbs loop (same operands every iteration)"}


# ----------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from bsreg_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the bsreg_b200 hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, peak_src = load_peaks()

    p = make_cfg4()
    n, k = p.X.shape
    # pinned host copies of the inputs (column-major X as R would hand it over)
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(np.zeros(1, a.dtype)).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    ty, tXt = pinned(p.y), pinned(np.ascontiguousarray(p.X.T))
    trp, tci, tva = pinned(p.W.indptr.astype(np.int32)), pinned(p.W.indices.astype(np.int32)), pinned(p.W.data)
    import scipy.sparse as sp
    Wp = sp.csr_matrix((tva.numpy(), tci.numpy(), trp.numpy()), shape=p.W.shape)
    yp, Xp = ty.numpy(), tXt.numpy().T          # Xp is an F-ordered (n, k) view of pinned memory
    h2d_bytes = yp.nbytes + Xp.nbytes + trp.numpy().nbytes + tci.numpy().nbytes + tva.numpy().nbytes

    def create(seed=2024, stats=capi.STATS_STREAM):
        return capi.Handle(yp, Xp, model=capi.MODEL_SAR, prior=capi.PRIOR_INDEPENDENT, W=Wp,
                           ldet=capi.LDET_CHEBYSHEV, cheb_order=CHEB_ORDER, cheb_probes=CHEB_PROBES,
                           stats_mode=stats, n_chains=CHAINS_PER_GPU, chain_offset=rank * CHAINS_PER_GPU, seed=seed)

    h = create()
    stream = torch.cuda.ExternalStream(h.stream())
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(handle, k_steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k_steps)]
        for a, b in ev:
            flush.zero_()
            torch.cuda.synchronize()
            a.record(stream)
            handle.run_into(0, ITERS_PER_STEP, None, None)     # draws stay on the device
            b.record(stream)
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in ev]

    # ---- device-resident throughput ----
    clocks = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        clocks.start()
        time.sleep(0.3)                                       # nvidia-smi is up before the warm-up steps begin
    timed_steps(h, args.warmup)
    barrier()
    clocks.mark_begin()
    l0 = h.kernel_launches()
    ms = timed_steps(h, args.steps)
    launches = h.kernel_launches() - l0
    clocks.mark_end()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    total_ms = float(sum(ms))
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    chain_iters = args.steps * ITERS_PER_STEP * CHAINS_PER_GPU * world
    value = chain_iters / (total_ms * 1e-3)

    # same loop with the sweep hoisted out (sufficient statistics cached at create)
    hc = create(stats=capi.STATS_CACHED)
    sc = torch.cuda.ExternalStream(hc.stream())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hc.run_into(0, 2 * ITERS_PER_STEP, None, None)      # warm-up with the timed shape (draw buffer, graphs)
    torch.cuda.synchronize()
    a.record(sc)
    hc.run_into(0, 2 * ITERS_PER_STEP, None, None)
    b.record(sc)
    torch.cuda.synchronize()
    cached_value = 2 * ITERS_PER_STEP * CHAINS_PER_GPU * world / (a.elapsed_time(b) * 1e-3)
    hc.close()

    # ---- kernel-level roofline for the dominant N-scale kernel (fused SpMM + Gram sweep) ----
    def time_launches(fn, reps, cold):
        if not cold:
            fn(20)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(reps); b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(1); b.record(stream)
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps

    roof = cpu = e2e = big = more = None
    traffic, traffic_src = ncu_traffic()
    big_p = None
    if not args.skip_big:
        from bsreg_b200.synthetic import make_problem, with_model
        big_p = make_problem("sem", n=CFG5_N, k_reg=K_REG, knn=KNN)          # cfg5's data; rank 0 reuses W, X for the SAR check
    if rank == 0 and not args.skip_big:
        big = hbm_resident_check(capi, torch, hbm_peak, with_model(big_p, "sar"))
        more = []
        for nc in (128, 256):                     # the same sweep serves more resident chains (not the headline config)
            with capi.Handle(yp, Xp, model=capi.MODEL_SAR, prior=capi.PRIOR_INDEPENDENT, W=Wp, ldet=capi.LDET_CHEBYSHEV,
                             cheb_order=CHEB_ORDER, cheb_probes=8, n_chains=nc, seed=7) as hm:
                sm_ = torch.cuda.ExternalStream(hm.stream())
                hm.run_into(0, 2 * ITERS_PER_STEP, None, None)      # warm-up with the timed shape (draw buffer, caches)
                torch.cuda.synchronize()
                tm = []
                for _ in range(3):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(sm_); hm.run_into(0, 2 * ITERS_PER_STEP, None, None); b.record(sm_)
                    torch.cuda.synchronize()
                    tm.append(a.elapsed_time(b))
                mt = float(np.mean(tm))
                more.append({"chains_per_gpu": nc, "value": nc * 2 * ITERS_PER_STEP / (mt * 1e-3),
                             "us_per_iteration": 1e3 * mt / (2 * ITERS_PER_STEP), "launches_timed": 3})
    if rank == 0:
        sweep_warm = time_launches(h.launch_sweep, 400, False)
        sweep_cold = time_launches(h.launch_sweep, 20, True)
        chain_ms = time_launches(h.launch_chain_step, 400, False)
        bytes_sweep = h.sweep_bytes()
        step_ms = total_ms / args.steps                       # one launch of the persistent kernel = one step
        in_loop = bytes_sweep * ITERS_PER_STEP / (step_ms * 1e-3) / 1e9
        alone = bytes_sweep / (sweep_warm * 1e-3) / 1e9
        roof = {"kernel": "gibbs_persistent_kernel<4, mma>: 126 sweep CTAs (TMA-fed fused CSR SpMV + Gram reduction on the FP64 "
                          "tensor cores, one pass over W, y, X per Gibbs iteration) + 22 chain CTAs (3 chains each), one "
                          "cooperative launch per step",
                "bound": "l2/issue (operands L2-resident): the chain step binds, the sweep side runs at 0.38 of the L2 roof",
                "achieved": in_loop, "peak": hbm_peak, "unit": "GB/s",
                "frac": in_loop / hbm_peak,
                "frac_is": "ALGORITHMIC bytes (W, y, X once per Gibbs iteration, SURVEY 8d) / launch time / measured HBM copy "
                           "peak.  At cfg4 the 29 MB of operands stay in L2 inside a launch, so this is NOT a fraction of an "
                           "HBM roofline (it could exceed 1): the iteration is max(sweep side, chain step).  The HBM-bound "
                           "evidence is hbm_bound_check (same kernel, N = 1e6)",
                "inloop_l2": {"l2_to_sm_peak_GBps": 18400.0, "l2_to_sm_peak_source": "profiles/r2_micro_l2_bandwidth.log "
                              "(bulk copies, 29 MB working set; 16-byte loads: 16900)",
                              "l2_traffic_GBps": 6950.0, "l2_to_sm_read_MB_per_iteration": 34.0, "frac_of_l2_peak": 0.38,
                              "source": "profiles/r2_persistent_inloop_metrics.txt: ncu single-pass counters of one "
                                        "500-iteration launch, build b564a8a (the persistent kernel has not changed since); "
                                        "recorded figures, not measured in this run"},
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": bytes_sweep * ITERS_PER_STEP,
                "algorithmic_bytes_per_iteration": bytes_sweep, "peak_source": peak_src,
                "launch_us": step_ms * 1e3, "iterations_per_launch": ITERS_PER_STEP,
                "note": "in-loop: the 29 MB of operands are L2-resident after the first iteration (same operands every "
                        "iteration), so DRAM traffic per launch is one pass, not iterations_per_launch passes; the "
                        "algorithmic figure is W, y, X streamed once per Gibbs iteration (SURVEY 8d).  The sweep CTAs and "
                        "the chain CTAs are balanced (sweep-bound and chain-bound times within 15 % of each other); "
                        "hbm_resident_check runs the same kernel at N = 1e6, where the operands (292 MB) exceed L2",
                "hbm_bound_check": big,
                "sweep_alone": {"kernel": "sweep_ring_kernel<4>, one launch per sweep on 148 SMs", "launch_us": sweep_warm * 1e3,
                                "achieved": alone, "frac": alone / hbm_peak,
                                "cold": {"launch_us": sweep_cold * 1e3, "achieved": bytes_sweep / (sweep_cold * 1e-3) / 1e9,
                                         "frac": bytes_sweep / (sweep_cold * 1e-3) / 1e9 / hbm_peak,
                                         "note": "512 MiB L2 flush before every launch (operands come from HBM)"}},
                "chain_step_alone_us": chain_ms * 1e3}
    h.close()

    # ---- end to end through the C ABI from pinned host buffers ----
    draws_host = torch.empty((CHAINS_PER_GPU, k + 2, E2E_SAVE), dtype=torch.float64, pin_memory=True)
    d2h_bytes = draws_host.numel() * 8

    def one_fit(seed):
        with create(seed=seed) as hh:
            hh.run_into(E2E_BURN, E2E_SAVE, None, draws_host.data_ptr())

    one_fit(1)
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        one_fit(100 + s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = e2e_steps * (E2E_BURN + E2E_SAVE) * CHAINS_PER_GPU * world / e2e_s
    lam_mean = float(draws_host.numpy()[:, -1, :].mean())

    # ---- cfg5 (BASELINE.json configs[4]): SEM, N = 1e6, 512 chains sharded over the GPUs ----
    cfg5 = cfg5_block(capi, torch, dist, rank, world, hbm_peak, big_p, barrier) if big_p is not None else None

    # everything below runs on rank 0 alone; the other ranks leave the GPUs idle and wait on a HOST barrier (an NCCL
    # barrier would keep their SMs spinning beside the one-handle fit and the CPU sample)
    hostg = dist.new_group(backend="gloo") if dist is not None else None
    if rank == 0:
        e2e = {"value": e2e_val, "unit": "chain-iter/s", "h2d_bytes_per_step": int(h2d_bytes),
               "d2h_bytes_per_step": int(d2h_bytes), "ms_per_fit": 1e3 * e2e_s / e2e_steps,
               "what": f"bsreg_create (H2D + setup + Chebyshev moments) + bsreg_run({E2E_BURN} burn, {E2E_SAVE} save) "
                       "+ D2H of draws, per fit, wall clock incl. host work; one process and one handle per GPU",
               "posterior_mean_lambda": lam_mean}
        if world > 1:
            e2e["one_handle"] = one_handle_fit(capi, torch, world, yp, Xp, Wp, k, h2d_bytes)
        # bounded CPU sample on this box's host cores
        cores = os.cpu_count() or 1
        procs = max(1, min(cores, CHAINS_PER_GPU))
        nodes_cpu = cpu_nodes(p)
        cpu_iters = 40 if args.skip_cpu else 2000                                # about 12-15 s of CPU work on every core used
        cpu_val, cpu_wall = cpu_throughput(p, nodes_cpu, cpu_iters, procs)
        cpu = {"value": cpu_val, "unit": "chain-iter/s", "cores": procs, "kind": "port",
               "sample": f"{procs} chains x {cpu_iters} iterations ({cpu_wall:.1f} s), one single-threaded process per chain",
               "note": "reference-algorithm CPU restatement (NumPy/SciPy), not R; R is not installed"}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "chain-iter/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world), "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clk,
            "cached_stats_value": cached_value, "more_resident_chains": more, "cfg5": cfg5,
            "hbm_roofline_chain_iter_s": hbm_peak * 1e9 / h_bytes_per_chain_iter(bytes_sweep) * world,
        }))
    if dist is not None:
        dist.barrier(group=hostg)
        dist.destroy_process_group()


def cfg5_block(capi, torch, dist, rank, world, hbm_peak, p, barrier):
    """This rank's share of cfg5: SEM, N = 1e6, kNN(10), K = 20; 512 chains over the GPUs at N >= 2 (256 / 128 / 64 per
    GPU), one 8-GPU share (64 chains) at N = 1.  Streamed (W, y, X swept every iteration: Gram matrix of
    [y | Wy | X | WX], d = 42) and cached statistics; device time of CFG5_ITERS iterations, mean of three launches of the
    warmed-up shape, MAX over ranks."""
    c_g = CFG5_CHAINS // world if world >= 2 else 64
    out = {"workload": f"cfg5: synthetic SEM, N={CFG5_N}, kNN(k=10), K=20, Chebyshev log-det (order 50, 8 probes); "
                       f"{c_g} chains per GPU" + ("" if world >= 2 else " (one GPU's share of the 8-GPU configuration)"),
           "chains_per_gpu": c_g, "global_chains": c_g * world, "iterations_timed": CFG5_ITERS, "launches_timed": 3,
           "timing": "CUDA events on the handle's stream, mean of 3 launches, MAX over ranks"}
    for mode, name in ((capi.STATS_STREAM, "stream"), (capi.STATS_CACHED, "cached")):
        with capi.Handle(p.y, p.X, model=capi.MODEL_SEM, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=CHEB_ORDER,
                         cheb_probes=8, n_chains=c_g, chain_offset=rank * c_g, seed=11, stats_mode=mode) as h:
            st = torch.cuda.ExternalStream(h.stream())
            h.run_into(0, CFG5_ITERS, None, None)                 # warm-up with the timed shape
            barrier()
            tm = []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st); h.run_into(0, CFG5_ITERS, None, None); b.record(st)
                torch.cuda.synchronize()
                tm.append(a.elapsed_time(b))
            ms = float(np.mean(tm))
            if dist is not None:
                t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            bytes_iter = h.sweep_bytes()
            lam = float(h.fetch_draws(0, min(CFG5_ITERS, 50))[:, -1, :].mean())
        gbs = bytes_iter * CFG5_ITERS / (ms * 1e-3) / 1e9
        out[name] = {"us_per_iteration": 1e3 * ms / CFG5_ITERS, "value": c_g * world * CFG5_ITERS / (ms * 1e-3),
                     "unit": "chain-iter/s", "lambda_mean_last_launch": lam}
        if name == "stream":
            out[name].update({"algorithmic_bytes_per_iteration": bytes_iter, "achieved": gbs, "peak": hbm_peak,
                              "frac": gbs / hbm_peak, "bound": "hbm (292 MB of operands per iteration exceed the 126 MB L2)"})
        barrier()
    return out


def one_handle_fit(capi, torch, world, yp, Xp, Wp, k, h2d_bytes):
    """North star: "draws are gathered to the host once at the end".  ONE process, ONE handle with devices = 0..N-1
    (bsreg_options.devices): 64 N chains, every shard set up by its own host thread, all draws written into one pinned
    host buffer.  Also checks that the draws of a chain id do not depend on where it ran: the last device's block equals
    a single-device run with chain_offset = 64 (N - 1) bit for bit."""
    nch = CHAINS_PER_GPU * world
    draws_all = torch.empty((nch, k + 2, E2E_SAVE), dtype=torch.float64, pin_memory=True)

    def fit(seed, **kw):
        kw = dict(n_chains=nch, devices=list(range(world))) | kw
        with capi.Handle(yp, Xp, model=capi.MODEL_SAR, prior=capi.PRIOR_INDEPENDENT, W=Wp, ldet=capi.LDET_CHEBYSHEV,
                         cheb_order=CHEB_ORDER, cheb_probes=CHEB_PROBES, seed=seed, **kw) as hh:
            hh.run_into(E2E_BURN, E2E_SAVE, None, draws_all.data_ptr())

    fit(1)
    fits = 3
    t0 = time.perf_counter()
    for s in range(fits):
        fit(100 + s)
    dt = time.perf_counter() - t0
    multi = draws_all.numpy()[CHAINS_PER_GPU * (world - 1):].copy()
    fit(100 + fits - 1, n_chains=CHAINS_PER_GPU, chain_offset=CHAINS_PER_GPU * (world - 1), devices=[0])
    single = draws_all.numpy()[:CHAINS_PER_GPU]
    return {"value": fits * (E2E_BURN + E2E_SAVE) * nch / dt, "unit": "chain-iter/s", "ms_per_fit": 1e3 * dt / fits,
            "devices": world, "chains": nch, "h2d_bytes_per_step": int(h2d_bytes) * world,
            "d2h_bytes_per_step": int(draws_all.numel() * 8),
            "draws_invariant": bool(np.array_equal(multi, single)),
            "what": "one process, one bsreg handle over all devices (host thread per shard at create), draws of all "
                    "chains into one pinned host buffer"}


def hbm_resident_check(capi, torch, hbm_peak, p):
    """The same persistent kernel at N = 1e6 (SAR, K = 20, 64 chains): 292 MB of operands per iteration do not fit the
    126 MB L2, so every iteration streams them from HBM.  Mean of three launches of the warmed-up shape."""
    with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=CHEB_ORDER,
                     cheb_probes=8, n_chains=CHAINS_PER_GPU, seed=5) as hb:
        st = torch.cuda.ExternalStream(hb.stream())
        hb.run_into(0, 400, None, None)                       # warm-up (clocks, instruction caches)
        torch.cuda.synchronize()
        iters, tm = 400, []
        for _ in range(3):                                    # mean of three launches of 400 iterations
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); hb.run_into(0, iters, None, None); b.record(st)
            torch.cuda.synchronize()
            tm.append(a.elapsed_time(b))
        ms = float(np.mean(tm))
        gbs = hb.sweep_bytes() * iters / (ms * 1e-3) / 1e9
        return {"workload": "synthetic SAR, N=1000000, kNN(k=10), K=20, 64 chains", "bound": "hbm",
                "timing": "mean of 3 launches x 400 iterations, warmed up with the same shape",
                "us_per_iteration": 1e3 * ms / iters, "us_per_iteration_min": 1e3 * min(tm) / iters,
                "algorithmic_bytes_per_iteration": hb.sweep_bytes(), "achieved": gbs, "unit": "GB/s", "frac": gbs / hbm_peak,
                "value": CHAINS_PER_GPU * iters / (ms * 1e-3)}


def h_bytes_per_chain_iter(bytes_sweep):
    return bytes_sweep / CHAINS_PER_GPU


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-big", action="store_true", help="skip the N = 1e6 HBM-resident check (saves ~25 s)")
    ap.add_argument("--skip-cpu", action="store_true", help="shrink the CPU baseline sample to 40 iterations (profiler runs)")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi (runs under ncu: a launch list, never a bench value)")
    ap.add_argument("--ref-iters", type=int, default=300, help="iterations per chain and step of the reference arm")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
