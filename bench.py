"""Benchmark of the hepmc hot path on B200 (contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[2], the one the north-star target is quoted on):
16-D camel VEGAS integration, divisions = 15, c = 3, 1e9 points per adaptation iteration
for the whole job (strong scaling: the 1e9 points are sharded over the N GPUs by Philox
counter range; one all-gather of the 8 slot partials per iteration).  A "step" is one
VEGAS iteration = fused sample kernel + ordered slot reduction + grid update.

  value   : points/s, device-timed (CUDA events, max over ranks), nothing memory-resident
            is needed as input (counter-based RNG), the grid lives in HBM.
  e2e     : the same metric through the public API `VegasMC(16, 15)(Camel(16), n, K)` from
            host state: host->device copy of the grid, all launches, device->host read of the
            per-iteration estimates; wall clock around the call.
  roofline: dominant kernel vegas_sample_kernel, timed alone with CUDA events.
  others  : RAMBO PS-points/s, banana RWM chain-steps/s, camel-8D HMC+ELM chain-steps/s.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_POINTS = 10 ** 9          # per VEGAS iteration, whole job
NDIM, DIVISIONS, DAMPING = 16, 15, 3
FLOP_PER_POINT = 12 * NDIM + 13          # 205 for n = 16
METRIC = "camel-16D VEGAS points/s"


# ------------------------------------------------------------------ reference arm / CPU baseline
def _cpu_vegas_worker(args):
    """One process = one independent VEGAS run of the NumPy port (the way the reference
    parallelises: multiprocessing.Pool over independent seeds, src/make_sample.py:156-157)."""
    seed, n, iters = args
    from oracle import hepmc_oracle as O
    rs = np.random.RandomState(seed)
    t0 = time.perf_counter()
    bounds = O.grid_uniform_bounds(NDIM, DIVISIONS)
    for j in range(iters):
        idx = np.stack([rs.randint(0, DIVISIONS, n) for _ in range(NDIM)])      # stratified_volume.py:110-111
        u = np.stack([rs.rand(n) for _ in range(NDIM)])                         # :119-121
        _, _, bounds, _ = O.vegas_iteration(O.camel_pdf, bounds, idx, u, DAMPING, j)
    return time.perf_counter() - t0


def cpu_baseline(n_per_step, steps, procs):
    import multiprocessing as mp
    if procs == 1:
        dt = _cpu_vegas_worker((1234, n_per_step, steps))
    else:
        ctx = mp.get_context("fork")
        t0 = time.perf_counter()
        with ctx.Pool(procs) as pool:
            pool.map(_cpu_vegas_worker, [(1234 + i, n_per_step, steps) for i in range(procs)])
        dt = time.perf_counter() - t0
    return procs * n_per_step * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    n_step = 200_000                      # bounded sample of the same workload per step and process
    cpu_baseline(n_step, 1, procs)        # warm-up (imports, page-in)
    value, dt = cpu_baseline(n_step, args.steps, procs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "camel-16D VEGAS, divisions=15, c=3 (BASELINE.json configs[2])",
                   "points_per_step_per_process": n_step, "processes": procs},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": procs, "kind": "port",
                         "sample": "%d processes x %d iterations x %d points of the NumPy port of "
                                   "VegasMC(16,15)(Camel(16)) (oracle/hepmc_oracle.py); the reference is pure "
                                   "Python and /root/reference does not exist on the GPU box" % (procs, args.steps, n_step)},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------ our arm
def run_b200(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import hepmc_b200 as H
    from hepmc_b200 import _lib, parallel, rng

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the harness's NCCL_DEBUG stays as it is: NCCL (a C library writing to fd 1) is pointed at stderr
        # for the life of the process, and the one JSON line goes to the saved stdout
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    H.set_outputs("torch")
    dev = torch.device("cuda", local_rank)
    steps, warmup = args.steps, max(args.warmup, 0)
    n = args.points

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    camel = H.densities.Camel(NDIM)
    mc = H.VegasMC(NDIM, divisions=DIVISIONS, c=DAMPING)
    H.seed(1234)
    if warmup:
        mc(camel, n, warmup, keep_samples=False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    res = mc(camel, n, steps, apriori=False, keep_samples=False)   # exactly K steps, one sync at its end
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = n * steps / (ms_total * 1e-3)

    # ---- e2e: public API from host state, wall clock, includes H2D of the grid + D2H of results
    barrier()
    t0 = time.perf_counter()
    mc_e2e = H.VegasMC(NDIM, divisions=DIVISIONS, c=DAMPING)
    r_e2e = mc_e2e(camel, n, steps, keep_samples=False)
    integral_host = float(r_e2e.integral)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = n * steps / float(dt.item())
    grid_bytes = NDIM * (2 * DIVISIONS + 1) * 8

    # ---- roofline of the dominant kernel: vegas_sample_kernel alone, CUDA events on its stream
    chunk = _lib.default_chunk(n)
    first, n_local, begins = parallel.local_span(n, chunk)
    nc = max(begins[-1], 1)
    sp = torch.zeros((nc, 3), dtype=torch.float64, device=dev)
    hp = torch.zeros((nc, NDIM * DIVISIONS), dtype=torch.float64, device=dev)
    blk = camel._block()
    grid = mc.volumes.device_grid()
    kms = []
    for i in range(3 + min(steps, 5)):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        _lib.call("hepmc_vegas_sample", ctypes.byref(blk), _lib.ptr(grid), NDIM, DIVISIONS, n_local, first, chunk,
                  1234, 1000 + i, rng.mode_code(), None, None, None, None, None, None, _lib.ptr(sp), _lib.ptr(hp),
                  _lib.stream_ptr())
        k1.record()
        torch.cuda.synchronize()
        if i >= 3:
            kms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kms))
    fp64_peak = _lib.fp64_probe(1 << 15) / 1e12          # measured DFMA TFLOP/s on this GPU, live
    achieved = FLOP_PER_POINT * n_local / (kernel_ms * 1e-3) / 1e12
    # the same kernel in the other two draw modes (the headline runs in the library default)
    mode_ms = {rng.get_mode(): kernel_ms}
    for mode in ("u52r10", "u32r10", "u32r7"):
        if mode in mode_ms:
            continue
        ts = []
        for i in range(4):
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            _lib.call("hepmc_vegas_sample", ctypes.byref(blk), _lib.ptr(grid), NDIM, DIVISIONS, n_local, first, chunk,
                      1234, 2000 + i, rng.MODES[mode], None, None, None, None, None, None, _lib.ptr(sp), _lib.ptr(hp),
                      _lib.stream_ptr())
            k1.record()
            torch.cuda.synchronize()
            if i >= 1:
                ts.append(k0.elapsed_time(k1))
        mode_ms[mode] = float(np.mean(ts))

    others = {}
    # the other configs: EVERY rank runs them (the library's reductions look for their peers); the
    # collective-free ones (RAMBO, RWM, HMC) are index-range sharded, whole-job size fixed
    if not args.no_others:
        others = other_workloads(H, _lib, torch, fp64_peak, world == 1 and not args.no_cpu, rank, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    cpu = None
    if world == 1 and not args.no_cpu:
        v, cdt = cpu_baseline(1_000_000, 2, 1)
        cpu = {"value": v, "unit": "points/s", "cores": 1, "kind": "port",
               "sample": "2 iterations x 1e6 points of the NumPy port of VegasMC(16,15)(Camel(16)) "
                         "(oracle/hepmc_oracle.py), single process, %.1f s" % cdt}
    wf = _ncu_wavefronts()
    clk = sampler.summary()
    smem_view = None
    if wf and n_local == N_POINTS and world == 1 and clk.get("sm_mhz"):
        n_sm = torch.cuda.get_device_properties(0).multi_processor_count
        peak_w = n_sm * clk["sm_mhz"] * 1e6
        smem_view = {"bound": "shared-memory data pipe", "unit": "wavefronts/s", "achieved": wf / (kernel_ms * 1e-3),
                     "peak": peak_w, "frac": wf / (kernel_ms * 1e-3) / peak_w, "wavefronts_per_launch": wf,
                     "wavefronts_per_point": wf / N_POINTS}
    line = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "camel-16D VEGAS, divisions=15, c=3, %.3g points/iteration whole job "
                               "(BASELINE.json configs[2]); keep_samples=False" % n,
                   "rng_mode": "%s (library default: 32 random bits per axis, Philox4x32-10; hepmc_b200.rng)" % rng.get_mode(),
                   "parallelism": "philox-counter sharding x%d, all-gather of 8 slot partials per iteration" % world,
                   "l2": "no memory-resident input (counter RNG in registers); partial rows are written, never re-read hot",
                   "integral": res.integral, "integral_err": res.integral_err},
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": grid_bytes / steps,
                "d2h_bytes_per_step": 16, "integral": integral_host},
        # per iteration: vegas_fast_kernel, reduce_slots_kernel, vegas_update_kernel (+ one NCCL all-gather at N > 1)
        "gpu_launches": 3 * steps,
        "clocks": clk,
        "roofline": {"bound": "fp64_simt", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size, read from the
                     # committed ncu --set full summary of the same command (the 59 MB of partial rows mostly
                     # stay in L2 until the reduction reads them)
                     "traffic": _ncu_traffic() if (n_local == N_POINTS and world == 1) else None,
                     "kernel": "vegas_fast_kernel", "kernel_ms": kernel_ms, "points_per_launch": n_local,
                     "kernel_ms_by_rng_mode": mode_ms,
                     "points_per_s_by_rng_mode": {k: n_local / (v * 1e-3) for k, v in mode_ms.items()},
                     "flop_per_point": FLOP_PER_POINT,
                     "peak_source": "live DFMA probe (hepmc_fp64_probe); MEASURED_PEAKS.json has no FP64 entry "
                                    "(hbm_gbs=%s)" % peaks.get("hbm_gbs"),
                     "kernel_share_of_step": kernel_ms / (ms_total / steps),
                     # the same launch read against HBM, to show which bound applies: the kernel has no
                     # memory-resident input and writes one partial row (3 + ndim*K doubles) per chunk
                     "hbm_view": {"bound": "hbm", "unit": "GB/s", "peak": peaks.get("hbm_gbs"),
                                  "achieved": nc * (3 + NDIM * DIVISIONS) * 8 / (kernel_ms * 1e-3) / 1e9,
                                  "bytes_per_launch": nc * (3 + NDIM * DIVISIONS) * 8},
                     # the binding resource: one 128-byte shared-memory wavefront per cycle and SM is the hardware rate;
                     # wavefronts per launch from the committed ncu capture (source page), SM clock sampled during the run
                     "smem_view": smem_view,
                     "note": "535 thread-instructions per point (round 1: 897), 134 of them FP64; issue slots 64 %, FP64 pipe "
                             "32 %; 168 shared-memory wavefronts per 32 points (64 grid look-ups, 64 histogram read-modify-"
                             "writes, 30 staging, 10 exponential-table look-ups; every access at its ideal count) -- the "
                             "shared-memory data pipe is the binding resource (ncu, profiles/" + NCU_SUMMARY + ")"},
        "cpu_baseline": cpu,
        "others": others,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _emit(line):
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


NCU_SUMMARY = "r02b_vegas_fast_kernel_ncu_summary.txt"


def _ncu_wavefronts():
    """Shared-memory wavefronts of one launch of the headline kernel (sum of `L1 Wavefronts Shared` over the source page
    of the committed ncu capture, 1e9 points)."""
    try:
        txt = open(os.path.join(ROOT, "profiles", NCU_SUMMARY)).read()
        line = [l for l in txt.splitlines() if l.startswith("shared-memory wavefronts per launch")][0]
        return float(line.split(":")[1].split()[0])
    except Exception:
        return None


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the headline kernel from the committed ncu summary."""
    try:
        txt = open(os.path.join(ROOT, "profiles", NCU_SUMMARY)).read()
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            line = [l for l in txt.splitlines() if key in l][0].split()
            tot += float(line[1]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[line[2]]
        return int(tot)
    except Exception:
        return None


def _best_of(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_ports():
    """Single-core NumPy-port timings of the other configs at CPU-feasible sizes."""
    from oracle import hepmc_oracle as O
    rs = np.random.RandomState(1234)
    out = {}
    n = 1_000_000
    t = _best_of(lambda: O.plain_mc(O.camel_pdf, rs.random_sample((n, 2))))
    out["plain2"] = {"value": n / t, "unit": "points/s", "cores": 1, "kind": "port", "sample": "PlainMC(2)(Camel(2), 1e6)"}

    def vegas2():
        idx = rs.randint(0, 50, (10, 2, 100_000))
        u = rs.rand(10, 2, 100_000)
        O.vegas(O.camel_pdf, 2, 50, idx, u)
    t = _best_of(vegas2, 2)
    out["vegas2"] = {"value": 1e6 / t, "unit": "points/s", "cores": 1, "kind": "port",
                     "sample": "VegasMC(2, divisions=50)(Camel(2), 1e5, 10)"}
    xs = rs.random_sample((n, 16))
    t = _best_of(lambda: O.rambo_map(xs, 100., 4))
    out["rambo"] = {"value": n / t, "unit": "PS-points/s", "cores": 1, "kind": "port", "sample": "Rambo(100, 4).map on 1e6 x 16"}
    C, T = 2000, 51
    z, u = rs.standard_normal((C, T - 1, 2)), rs.rand(C, T - 1)
    t = _best_of(lambda: O.metropolis_chains(O.banana_pdf, np.tile([0., 10.], (C, 1)), z, u, np.sqrt([10., 10.])), 2)
    out["rwm"] = {"value": C * (T - 1) / t, "unit": "chain-steps/s", "cores": 1, "kind": "port",
                  "sample": "banana-2D RWM, NumPy port vectorised over 2000 chains x 50 steps (the reference itself "
                            "steps ONE chain in a Python loop: 6.2e3 steps/s, BASELINE.md)"}
    for nodes in (100, 1024):
        centers, widths, w = rs.random_sample((nodes, 8)), 0.01 + 0.99 * rs.random_sample(nodes), 0.01 * rs.standard_normal(nodes)
        C, T = 200, 3
        p, uu = rs.standard_normal((C, T - 1, 8)), rs.rand(C, T - 1)
        grad = lambda q: O.elm_gauss_eval_gradient(q, centers, widths, w)
        t = _best_of(lambda: O.hmc_chains(O.camel_pdf, grad, np.full((C, 8), 1 / 3), p, uu, 1.0, 0.01, 20), 1)
        out["hmc_elm%d" % nodes] = {"value": C * (T - 1) / t, "unit": "chain-steps/s", "cores": 1, "kind": "port",
                                    "sample": "camel-8D HMC + GaussianBasis(%d) gradient, NumPy port vectorised over 200 chains x 2 "
                                              "steps x 20 leapfrog (reference, one chain: 770 / 311 steps/s, BASELINE.md)" % nodes}
    return out


def other_workloads(H, _lib, torch, fp64_peak, with_cpu, rank=0, world=1):
    """The other configs BASELINE.json names, device-timed.  Called by every rank: the
    independent-item workloads are split by global index range (parallel.item_span), timed
    between barriers, max over ranks; values are whole-job throughputs."""
    import torch.distributed as dist
    from hepmc_b200 import parallel
    out = {}

    def timed(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3)
        t = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    hbm = None
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    cpu = cpu_ports() if with_cpu else {}
    if world == 1:
        # C1: 2-D camel, 1M points: plain MC and VEGAS 10 x 1e5 (launch-latency dominated at this
        # size, so only quoted on one GPU)
        camel2 = H.densities.Camel(2)
        t = timed(lambda: H.PlainMC(2)(camel2, 1_000_000, keep_samples=False))
        out["c1_plain2"] = {"metric": "camel-2D plain MC points/s (1e6 points, whole call)", "value": 1e6 / t,
                            "cpu_baseline": cpu.get("plain2")}
        mc2 = H.VegasMC(2, divisions=50)
        t = timed(lambda: mc2(camel2, 100_000, 10, keep_samples=False))
        out["c1_vegas2"] = {"metric": "camel-2D VEGAS points/s (10 x 1e5 points, whole call)", "value": 1e6 / t,
                            "cpu_baseline": cpu.get("vegas2")}
        # the reference's default mode keeps every point: data + function values (+ weights) are
        # written out, so these two are read against HBM (outputs of 7 GB / 14 GB, larger than L2)
        camel16 = H.densities.Camel(16)
        nk = 50_000_000
        t = timed(lambda: H.PlainMC(16)(camel16, nk, keep_samples=True))
        out["plain16_keep"] = {"metric": "camel-16D plain MC points/s, keep_samples=True (5e7 points)", "value": nk / t,
                               "roofline": {"bound": "hbm", "achieved": nk * 17 * 8 / t / 1e9, "peak": hbm, "unit": "GB/s",
                                            "frac": nk * 17 * 8 / t / 1e9 / hbm, "bytes_per_point": 136}}
        mck = H.VegasMC(16, divisions=15)
        t = timed(lambda: mck(camel16, nk, 2, keep_samples=True))
        out["vegas16_keep"] = {"metric": "camel-16D VEGAS points/s, keep_samples=True (2 x 5e7 points)", "value": 2 * nk / t,
                               "roofline": {"bound": "hbm", "achieved": 2 * nk * 18 * 8 / t / 1e9, "peak": hbm, "unit": "GB/s",
                                            "frac": 2 * nk * 18 * 8 / t / 1e9 / hbm, "bytes_per_point": 144}}
        del mck
        # the sample kernels alone (C ABI, CUDA events, 1e8 points per launch): what the whole-call figures above add
        # on top is the allocation of the returned arrays and the host side of the call
        import ctypes
        nkk = 100_000_000
        chunk = _lib.default_chunk(nkk)
        ncl = (nkk + chunk - 1) // chunk
        blk = camel16._block()
        xk = torch.empty((16, nkk), dtype=torch.float64, device="cuda")
        fk = torch.empty(nkk, dtype=torch.float64, device="cuda")
        wk = torch.empty(nkk, dtype=torch.float64, device="cuda")
        sp = torch.zeros((ncl, 3), dtype=torch.float64, device="cuda")
        hp = torch.zeros((ncl, 16 * 15), dtype=torch.float64, device="cuda")
        gk = H.GridVolumes(ndim=16, divisions=15).device_grid()
        st = _lib.stream_ptr()
        t = timed(lambda: _lib.call("hepmc_vegas_sample", ctypes.byref(blk), _lib.ptr(gk), 16, 15, nkk, 0, chunk, 1234, 7,
                                    H.rng.mode_code(), None, None, _lib.ptr(xk), _lib.ptr(fk), _lib.ptr(wk), None,
                                    _lib.ptr(sp), _lib.ptr(hp), st))
        out["vegas16_keep"]["kernel_roofline"] = {"bound": "hbm", "achieved": nkk * 18 * 8 / t / 1e9, "peak": hbm, "unit": "GB/s",
                                                  "frac": nkk * 18 * 8 / t / 1e9 / hbm, "points_per_s": nkk / t,
                                                  "what": "hepmc_vegas_sample with x, f, w outputs, 1e8 points per launch"}
        xr = xk.view(-1)[:nkk * 16]
        t = timed(lambda: _lib.call("hepmc_plain_mc_sample", ctypes.byref(blk), nkk, 0, chunk, 1234, 7, H.rng.mode_code(), None,
                                    _lib.ptr(xr), _lib.ptr(fk), _lib.ptr(sp), st))
        out["plain16_keep"]["kernel_roofline"] = {"bound": "hbm", "achieved": nkk * 17 * 8 / t / 1e9, "peak": hbm, "unit": "GB/s",
                                                  "frac": nkk * 17 * 8 / t / 1e9 / hbm, "points_per_s": nkk / t,
                                                  "what": "hepmc_plain_mc_sample with data and f outputs, 1e8 points per launch",
                                                  "note": "the peak is the measured COPY bandwidth (a read and a write stream); "
                                                          "a write-only stream can exceed it"}
        del xk, fk, wk, sp, hp, xr
    # C2: RAMBO 2->4, sqrt(s) = 100 GeV.  BASELINE.json configs[1]: 1e9 points sharded over 8 GPUs (16 GB of output
    # per GPU); on fewer GPUs 2^27 points per launch whole job (16 GiB; larger than L2 at every N)
    m = H.phase_space.Rambo(100., 4)
    nr = 10 ** 9 if world == 8 else 1 << 27
    first, cnt = parallel.item_span(nr, rank, world)
    buf = torch.empty((cnt, 16), dtype=torch.float64, device="cuda")
    t = timed(lambda: m.generate(cnt, first_index=first, out=buf))
    del buf
    out["rambo"] = {"metric": "RAMBO 2->4 PS-points/s", "value": nr / t, "points": nr, "rng_mode": H.rng.get_mode(),
                    "roofline": {"bound": "hbm", "achieved": nr * 128 / t / 1e9 / world, "peak": hbm, "unit": "GB/s",
                                 "frac": nr * 128 / t / 1e9 / world / hbm, "bytes_per_point": 128, "per": "GPU",
                                 "fp64_tflops_algorithmic": 183 * nr / t / 1e12 / world,
                                 "note": "ncu (profiles/r02b_rambo_kernel_ncu_summary.txt): 620 thread-instructions per point of "
                                         "which 284 FP64; FP64 pipe 54 %, issue slots 58 %, DRAM 53 % -- compute bound below the "
                                         "write roofline, as SURVEY 8(d) expected"},
                    "cpu_baseline": cpu.get("rambo")}
    # C4: banana 2-D RWM: 1e6 chains x 1e4 steps in ONE launch (BASELINE.json configs[3] at full size)
    C, T = 1_000_000, 10_001
    first, cnt = parallel.item_span(C, rank, world)
    met = H.DefaultMetropolis(2, H.densities.Banana(2), H.proposals.Gaussian(2, cov=10 * np.eye(2)))
    x0 = torch.tensor([[0., 10.]], device="cuda", dtype=torch.float64).repeat(cnt, 1)
    t = timed(lambda: met.sample(T, x0, store_chain=False, first_chain=first), reps=2)
    out["rwm"] = {"metric": "banana-2D RWM chain-steps/s", "value": C * (T - 1) / t, "chains": C, "steps": T - 1,
                  "roofline": {"bound": "fp64_simt", "achieved": 23 * C * (T - 1) / t / 1e12 / world, "peak": fp64_peak,
                               "unit": "TFLOP/s", "frac": 23 * C * (T - 1) / t / 1e12 / world / fp64_peak,
                               "flop_per_step": 23, "per": "GPU",
                               "note": "rwm_fast_kernel (ncu, profiles/r02b_rwm_kernels_ncu_summary.txt): 189 thread-instructions "
                                       "per step of which 72 FP64 (Philox 45, Box-Muller log / sqrt / sincos, the target's exp); "
                                       "issue slots 67 %, FP64 pipe 51 % busy -- the 23-flop convention counts each "
                                       "transcendental as one"},
                  "cpu_baseline": cpu.get("rwm")}
    # C5: camel-8D HMC, 1e5 chains x 20 leapfrog, ELM Gaussian-basis surrogate gradient TRAINED by the SURVEY 8(d)
    # recipe: 4000 points (2000 per mode, N(mu, 0.005 I)), values = UnconstrainedCamel(8).pot(xs), random centres
    # U[0,1)^(I x 8) and widths U[0.01, 1) (extreme_learning_train's own draw), plain pseudo-inverse (ridge = 0: the
    # reference, interfaces/mc3_hmc_el.py:24-29) and, as a second line, the ridge = 1e-6 extension.
    # Rooflines: float64 -> FP64 SIMT by algorithmic flops; f32 / tc32 / tc -> one MUFU.EX2 per (chain, node,
    # gradient evaluation) against the live MUFU probe (the exponential, not the MMA, bounds this path).
    mufu_peak = _lib.mufu_probe(1 << 14)
    rs = np.random.RandomState(1234)
    xs_train = np.concatenate([1 / 3 + np.sqrt(0.005) * rs.standard_normal((2000, 8)),
                               2 / 3 + np.sqrt(0.005) * rs.standard_normal((2000, 8))])
    v_train = H.densities.UnconstrainedCamel(8).pot(xs_train)
    if isinstance(v_train, torch.Tensor):
        v_train = v_train.cpu().numpy()
    for nodes in (100, 1024):
        for ridge in (0.0, 1e-6):
            basis = H.surrogate.GaussianBasis(8)
            np.random.seed(4321 + nodes)
            H.seed(4321 + nodes)
            params, bias, w = basis.extreme_learning_train(xs_train, v_train, nodes, ridge=ridge)
            for prec in ("f64", "f32", "tc32", "tc"):
                if ridge and prec != "tc":
                    continue                       # the extension only changes which kernel 'tc' may keep
                tgt = H.densities.Camel(8)
                grad = H.surrogate.attach_surrogate_gradient(tgt, basis, params, bias, w)
                upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision=prec)
                Ch, Th = 100_000, 3
                first, cnt = parallel.item_span(Ch, rank, world)
                x0h = torch.full((cnt, 8), 1 / 3, device="cuda", dtype=torch.float64)
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    t = timed(lambda: upd.sample(Th, x0h, store_chain=False, first_chain=first, total_chains=Ch), reps=2)
                    acc = upd.sample(11, x0h[:20000], store_chain=False, total_chains=Ch).accepted
                acc = float(acc.double().mean() if isinstance(acc, torch.Tensor) else np.mean(acc)) / 10
                rate = Ch * (Th - 1) / t
                flop = 21 * (4 * nodes * 8 + 5 * nodes + 4 * 8) + 20 * 64 + 72 + 48
                ran = upd.kernel_precision
                if ran == "f64":
                    roof = {"bound": "fp64_simt", "achieved": flop * rate / 1e12 / world, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": flop * rate / 1e12 / world / fp64_peak, "flop_per_chain_step": flop, "per": "GPU"}
                else:
                    ex = 21 * nodes * rate / world
                    roof = {"bound": "mufu_ex2", "achieved": ex / 1e12, "peak": mufu_peak / 1e12, "unit": "Texp/s",
                            "frac": ex / mufu_peak, "exp_per_chain_step": 21 * nodes, "per": "GPU",
                            "peak_source": "live MUFU.EX2 probe (hepmc_mufu_probe)",
                            "tensor_pipe": "sm__pipe_tensor_cycles_active 30.5 % in the ncu capture of hmc_tc_kernel "
                                           "(profiles/r01_hmc_tc_kernel_ncu_summary.txt)" if ran in ("tc", "tc32") else None}
                key = "hmc_elm%d_%s%s" % (nodes, prec, "_ridge" if ridge else "")
                out[key] = {"metric": "camel-8D HMC+ELM chain-steps/s", "value": rate, "nodes": nodes,
                            "surrogate": "trained: extreme_learning_train on 4000 points, ridge=%g" % ridge,
                            "gradient_cancellation": grad.cancellation(x0h[:256]) , "precision_requested": prec,
                            "kernel_ran": ran, "acceptance": acc, "roofline": roof,
                            "cpu_baseline": cpu.get("hmc_elm%d" % nodes)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=float, default=N_POINTS)
    ap.add_argument("--no-others", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.points = int(args.points)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
