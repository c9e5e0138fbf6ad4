#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Fourier option-pricing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], the configuration the metric / north-star target is quoted
on): Heston (CIR-time-changed Brownian motion, generateCharts.cpp:231-250) COS calibration sweep,
8 maturities x 66 strikes (64 quoted + the two synthetic far strikes of generateCharts.cpp:43-48),
numU = 256, weak scaling with 125 000 parameter sets per GPU (8 GPUs = the 1e6-set sweep).
One step = one pass over the rank's shard: every strike price is computed, written to HBM,
reduced to the per-set weighted squared error against market quotes (fop_cos_loss), and for N > 1
the per-set losses are all-gathered through the library's own collective (fop_gather_losses_async:
NCCL on a side stream, overlapped with the next step).  A price counts once per (set, maturity, quoted
strike): 64 per row, although 66 are evaluated.

Prints ONE JSON line (rank 0).  `value` = prices/s with inputs resident in HBM; `e2e` = the same
through the C ABI with HOST buffers (H2D of the parameter sets and D2H of the losses inside the
timed region); `roofline` = live CUDA-event timing of the dominant kernel against the fp64 peak
measured in the same process; `cpu_baseline` = the reference's own headers (oracle/_ref) on the
host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import cases as CASES  # noqa: E402  (synthetic inputs of SURVEY.md 8d; no oracle code)

METRIC = "option prices/sec (fp64, batched param sets x strikes)"
UNIT = "prices/s"
SETS_PER_GPU = 125_000
NUM_U = 256
S0, RATE = 100.0, 0.0
COUNTED_STRIKES = 64


def workload_config(sets_per_gpu, n_gpus):
    return {
        "workload": "C5 Heston COS calibration sweep (BASELINE.json configs[4]): "
                    f"{sets_per_gpu} parameter sets/GPU x 8 maturities x 64 quoted strikes "
                    "(+2 synthetic end strikes evaluated), numU=256, fused per-set loss, "
                    "NCCL all-gather of losses for N>1",
        "sets_per_gpu": sets_per_gpu,
        "global_sets": sets_per_gpu * n_gpus,
        "maturities": len(CASES.C5_T),
        "strikes_counted": COUNTED_STRIKES,
        "strikes_evaluated": 66,
        "numU": NUM_U,
        "model": "GAUSS+CIR time change (the reference's Heston)",
        "parallelism": f"parameter sets sharded over {n_gpus} GPU(s), no data-path collective; "
                       "loss all-gather only",
        "l2_policy": "working set per step (4 GiB coefficient matrix + 528 MB prices) >> 126 MB L2",
    }


def market_and_weights(K, T):
    # synthetic "market": Black-Scholes prices at a fixed vol (inputs only; identical bytes for
    # every arm).  End strikes carry weight 0.
    import math
    mkt = np.array([[CASES.bs_call(S0, math.exp(-RATE * t), k, 0.25, t) for k in K] for t in T])
    return mkt, CASES.c5_weights()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons of one GPU sampled DURING the timed regions: an NVML polling
    thread (every ~2 ms; the timed region of a short run lasts only tens of ms), falling back to
    `nvidia-smi -lms` when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, uuid=None):
        import threading
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []
        self.stop_flag = False
        self.thread = None
        self.p = None
        self.mode = None
        try:
            import pynvml as N
            N.nvmlInit()
            h = None
            if uuid:
                try:
                    h = N.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
                except Exception:
                    h = None
            if h is None:
                h = N.nvmlDeviceGetHandleByIndex(gpu_index)
            self.N, self.h = N, h
            self.mx.append(float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)))
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.p = subprocess.Popen(
                    ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                     "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                    stderr=subprocess.DEVNULL)
                self.mode = "nvidia-smi"
            except Exception:
                self.p = None

    def _poll(self):
        N, h = self.N, self.h
        bits = {"hw_slowdown": getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(N, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
                self.power.append(N.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.mode}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=float(max(self.mx)),
                           samples=len(self.sm), sm_min_mhz=float(min(self.sm)),
                           power_w_max=float(max(self.power)) if self.power else None)
            out["reasons"] = sorted(self.reasons)
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def load_ncu_summary():
    """Per-launch figures of the two hot kernels from the committed `ncu --set full` capture of this
    very command (profiles/README.md): DRAM bytes per launch, and for the coefficient kernel the
    executed fp64 operations per characteristic-function evaluation.  {} when absent."""
    path = os.path.join(ROOT, "profiles", "r2_bench_top_kernels_ncu.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r1_bench_top_kernels_ncu.json")
    out = {}

    def num(s):
        v, u = (s.split() + [""])[:2]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    try:
        for k in json.load(open(path)):
            name = "cos_gemm_kernel" if "cos_gemm" in k["kernel"] else \
                "cos_coeff_kernel" if "cos_coeff" in k["kernel"] else None
            if not name:
                continue
            e = {"dram_bytes": num(k["dram__bytes_read.sum"]) + num(k["dram__bytes_write.sum"])}
            if "fp64_flops_per_cf_eval" in k:
                e["fp64_flops_per_cf_eval"] = float(k["fp64_flops_per_cf_eval"])
            out[name] = e
    except Exception:
        pass
    return out


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(seconds_target, sets_hint=None):
    """Times the reference CPU implementation (oracle/_ref = reference headers + shims, else the
    port) on a bounded sample of the workload.  The only place bench.py touches oracle/."""
    from oracle import build as obuild
    from oracle import oracle as O
    obuild.build()
    ora = O.load("best")
    K, T = CASES.c5_strikes(S0), CASES.C5_T
    mkt, w = market_and_weights(K, T)
    nthreads = ora.num_threads
    probe = max(8, 2 * nthreads)
    prm = CASES.heston_batch(probe)
    t0 = time.perf_counter()
    ora.cos_loss(0, 1, prm, K, S0, RATE, T, NUM_U, mkt, w)
    dt = time.perf_counter() - t0
    nsets = sets_hint or int(max(probe, min(200_000, probe * seconds_target / max(dt, 1e-6))))
    prm = CASES.heston_batch(nsets)
    t0 = time.perf_counter()
    ora.cos_loss(0, 1, prm, K, S0, RATE, T, NUM_U, mkt, w)
    dt = time.perf_counter() - t0
    rate = nsets * len(T) * COUNTED_STRIKES / dt
    return {"value": rate, "unit": UNIT, "cores": nthreads,
            "kind": "reference" if ora.flavor == "reference" else "port",
            "sample": f"{nsets} parameter sets x 8 maturities x 66 strikes of the same workload "
                      f"(FangOostCallPrice + loss per set, OpenMP over sets, {dt:.1f} s)",
            "seconds": dt, "sets": nsets}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is meant to use every
    # host core (set before libgomp is loaded with the oracle library)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    steps, warm = args.steps, args.warmup
    first = cpu_reference_rate(seconds_target=2.0)
    # size each step for ~2 s so the whole run stays within a few minutes
    sets = max(16, int(first["sets"] * 2.0 / max(first["seconds"], 1e-6)))
    sets = min(sets, 50_000)
    for _ in range(warm):
        cpu_reference_rate(0, sets_hint=sets)
    t0 = time.perf_counter()
    res = None
    for _ in range(steps):
        res = cpu_reference_rate(0, sets_hint=sets)
    dt = time.perf_counter() - t0
    value = steps * sets * len(CASES.C5_T) * COUNTED_STRIKES / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(SETS_PER_GPU, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                         "sample": f"each step = {sets} parameter sets x 8 maturities x 66 strikes "
                                   "of the workload on the host cores (reference headers compiled "
                                   "against oracle/shims, OpenMP over sets)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def extra_configs(eng, torch, dev, peak, stream):
    """The other BASELINE.json configs (C2, C3 at P = 4096, C4 at P = 16384) on one GPU, device
    resident, timed per kernel with the library's CUDA-event profiling: driver-observed lines for
    the kernels the headline workload does not touch.  Roofline per SURVEY.md 8d."""
    import fftoptionpricing_b200 as F
    out = {}

    def timed(fn, reps):
        """(wall ms per call by CUDA events on the handle's stream, per-kernel ms from a separate
        profiled pass -- kernels of one call may overlap on side streams, so their sum can exceed the wall)"""
        fn()
        fn()
        eng.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        wall = e0.elapsed_time(e1) / reps
        eng.profile_enable(True)
        for _ in range(reps):
            fn()
        prof = eng.profile_read()
        eng.profile_enable(False)
        return wall, {k: v[0] / reps for k, v in prof.items()}

    # C1: Black-Scholes Carr-Madan, N = 2^10, ONE parameter set, host in / host out: a latency case
    # (the reference's own CPU-runnable configuration, test.cpp:619-648)
    eng.carr_madan(F.GAUSS, 0, [.3], 1024, .25, 1.5, 50.0, .05, 1.0)
    t0 = time.perf_counter()
    for _ in range(200):
        eng.carr_madan(F.GAUSS, 0, [.3], 1024, .25, 1.5, 50.0, .05, 1.0)
    dt = (time.perf_counter() - t0) / 200
    out["C1"] = {"workload": "Black-Scholes Carr-Madan 2^10 points x 1 set, host -> host call (latency)",
                 "ms": dt * 1e3, "value": 1024 / dt, "unit": UNIT,
                 "roofline": {"bound": "latency", "frac": None,
                              "what": "one 24 us single-CTA kernel + H2D / D2H / synchronise per call: launch-latency bound by construction"}}
    # C2: Heston COS, 256 u x 1024 strikes, one maturity (batch of 4096 sets)
    P = 4096
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K = CASES.fang_oost_strikes(-5.0, 5.0, 100.0, 1024)
    o = torch.empty((P, 1, 1024), dtype=torch.float64, device=dev)
    ms, k = timed(lambda: eng.cos(F.GAUSS, F.TC_CIR, prm, K, 100.0, 0.0, [1.0], 256, 0, out=o), 5)
    fl = P * 1024 * 4.0 * 256
    out["C2"] = {"workload": f"Heston COS 256 u x 1024 strikes x {P} sets", "ms": ms, "kernels_ms": k,
                 "value": P * 1024 / (ms * 1e-3), "unit": UNIT,
                 "roofline": {"bound": "tensor", "algorithmic_flops": fl, "achieved": fl / (ms * 1e-3) / 1e12,
                              "peak": peak, "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / peak,
                              "what": "4 numU flops per price over the whole call (coefficient + contraction kernels) "
                                      "against the fp64 peak; fp64-EQUIVALENT (may exceed 1) when the contraction "
                                      "runs as exact int8 digit planes on tcgen05",
                              "contraction": eng.last_cos_contraction}}
    del o
    # C3: Merton American put FSTS, N = 4096, 256 steps, 4096 sets
    P = 4096
    prm = torch.tensor(CASES.merton_batch(P), device=dev)
    o = torch.empty((P, 4096), dtype=torch.float64, device=dev)
    ms, k = timed(lambda: eng.fsts(F.MERTON, 0, prm, 4096, 7.5, 100.0, .05, .25, steps=256, american=True,
                                   payoff=1, out=o), 3)
    fl = P * (256 * (2 * 5 * 4096 * 12 + 8 * 4096.0))
    out["C3"] = {"workload": f"Merton American put FSTS 2^12 nodes x 256 steps x {P} sets", "ms": ms,
                 "kernels_ms": k, "value": P * 4096 / (ms * 1e-3), "unit": UNIT,
                 "roofline": {"bound": "tensor", "algorithmic_flops": fl, "achieved": fl / (ms * 1e-3) / 1e12,
                              "peak": peak, "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / peak,
                              "what": "steps x (2 x 5 N log2 N + 8 N) flops per set (SURVEY 8d)"}}
    del o
    # C4: CGMY Carr-Madan, N = 2^16, 16384 sets, all strikes (8 GiB of prices)
    P = 16384
    prm = torch.tensor(CASES.cgmy_batch(P), device=dev)
    o = torch.empty((P, 65536), dtype=torch.float64, device=dev)
    ms, k = timed(lambda: eng.carr_madan(F.CGMY, 0, prm, 65536, .25, 1.5, 500.0, .4, .25, False, out=o), 3)
    by = P * 65536 * 8.0
    hbm = measured_peaks().get("hbm_gbs", 6554.6)
    out["C4"] = {"workload": f"CGMY Carr-Madan 2^16 points x {P} sets, all strikes", "ms": ms, "kernels_ms": k,
                 "value": P * 65536 / (ms * 1e-3), "unit": UNIT,
                 "roofline": {"bound": "hbm", "algorithmic_bytes": by, "achieved": by / (ms * 1e-3) / 1e9,
                              "peak": hbm, "unit": "GB/s", "frac": by / (ms * 1e-3) / 1e9 / hbm,
                              "what": "8 B per price written once (SURVEY 8d); 5 N log2 N accounting: %.2f TFLOP/s"
                                      % (P * 5 * 65536 * 16 / (ms * 1e-3) / 1e12)}}
    del o
    torch.cuda.empty_cache()
    return out


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import fftoptionpricing_b200 as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # torch.distributed is the launcher-side plumbing (rendezvous, barrier, max-over-ranks of
        # the timings); the data-path collective is the library's own fop_gather_losses*
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    eng = F.Engine(local)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    if world > 1:
        box = [F.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], rank, world)

    P = args.sets_per_gpu
    K, T = CASES.c5_strikes(S0), CASES.C5_T
    M, J = len(T), len(K)
    mkt, w = market_and_weights(K, T)
    # rank r prices the r-th contiguous slice of the global sweep (seeded synthetic inputs)
    prm_host = torch.from_numpy(CASES.heston_batch(P, seed=CASES.SEED + 1000 * rank)).pin_memory()
    loss_host = torch.empty(P, dtype=torch.float64).pin_memory()
    prm_dev = prm_host.to(dev)
    prices = torch.empty((P, M, J), dtype=torch.float64, device=dev)
    # two (loss, gathered) buffer pairs alternate: the gather of step i runs beside step i + 1
    loss = [torch.empty(P, dtype=torch.float64, device=dev) for _ in range(2)]
    gathered = [torch.empty(P * world, dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    mkt_d = torch.from_numpy(mkt).to(dev)
    w_d = torch.from_numpy(w).to(dev)

    def step_device(i):
        b = i & 1
        eng.cos_loss(F.GAUSS, F.TC_CIR, prm_dev, K, S0, RATE, T, NUM_U, mkt_d, w_d, out=loss[b],
                     prices_out=prices)
        if world > 1:
            eng.gather_losses_async(loss[b], gathered[b])   # side stream, overlaps step i + 1

    def finish_device():
        if world > 1:
            eng.comm_wait()

    d2h_done = torch.cuda.Event()

    def step_host(i=0):
        # host buffers in, host losses out: H2D of the parameter sets (and of the strike / maturity /
        # market / weight tables) + D2H of this rank's losses inside the timed region; prices are stored
        # to HBM exactly as in the device-resident step.
        if world == 1:
            # host loss pointer: the call stages, copies back and returns after the stream is idle
            eng.cos_loss(F.GAUSS, F.TC_CIR, prm_host.numpy(), K, S0, RATE, T, NUM_U, mkt, w,
                         out=loss_host.numpy(), prices_out=prices)
            return
        # N > 1: the losses stay on the device for the library's gather (side stream, the gathered
        # vector is what the next optimiser step consumes on the GPU); this rank's losses are read
        # back to pinned host memory every step and the host waits for them
        b = i & 1
        eng.cos_loss(F.GAUSS, F.TC_CIR, prm_host.numpy(), K, S0, RATE, T, NUM_U, mkt, w, out=loss[b],
                     prices_out=prices)
        loss_host.copy_(loss[b], non_blocking=True)
        d2h_done.record(stream)
        eng.gather_losses_async(loss[b], gathered[b])
        d2h_done.synchronize()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_device(nsteps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(nsteps):
            step_device(i)
        finish_device()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    uuid = str(getattr(torch.cuda.get_device_properties(local), "uuid", "") or "")
    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, 3)):
            step_device(i)
        finish_device()
        barrier()
        # ---- timed region: device-resident inputs
        sampler = ClockSampler(local, uuid) if rank == 0 else None
        l0 = eng.launch_count
        ms = timed_device(args.steps)
        launches = eng.launch_count - l0

        # ---- end to end through the C ABI with host buffers
        for i in range(2):
            step_host(i)
        finish_device()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_host(i)
        finish_device()
        barrier()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None   # sampled over both timed regions

        # ---- sustained leg: the same device-resident step for >= args.sustain_seconds
        sustained = None
        if args.sustain_seconds > 0:
            # the same step count on every rank (each step contains a collective)
            ms_all = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms_all, op=dist.ReduceOp.MAX)
            n_sus = max(args.steps, int(args.sustain_seconds * 1e3 / max(float(ms_all[0]) / args.steps, 1e-3)) + 1)
            sampler2 = ClockSampler(local, uuid) if rank == 0 else None
            ms_sus = timed_device(n_sus)
            clocks2 = sampler2.stop() if sampler2 else None
            sustained = (n_sus, ms_sus, clocks2)

        # ---- per-kernel timing for the roofline (separate pass so events do not perturb `value`)
        eng.profile_enable(True)
        for _ in range(args.steps):
            eng.cos_loss(F.GAUSS, F.TC_CIR, prm_dev, K, S0, RATE, T, NUM_U, mkt_d, w_d, out=loss[0],
                         prices_out=prices)
        prof = eng.profile_read()
        eng.profile_enable(False)
        contraction = eng.last_cos_contraction   # (before the extra configurations run other shapes)
        # ---- the same step with EVERY kernel in fp64 (FOP_GEMM_I8=0: fp64 tensor-core contraction), for a
        # reader who wants the all-fp64 figure next to the default (exact int8 digit planes, DESIGN.md
        # decision 12); untimed warm-up, then `steps` steps, device resident, this rank
        all_fp64 = None
        if contraction == "int8-tcgen05":
            os.environ["FOP_GEMM_I8"] = "0"
            for i in range(3):
                step_device(i)
            finish_device()
            torch.cuda.synchronize(dev)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(args.steps):
                step_device(i)
            finish_device()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            all_fp64 = e0.elapsed_time(e1) / args.steps
            os.environ.pop("FOP_GEMM_I8", None)
        peaks = eng.microbench_fp64() if rank == 0 else None
        i8_peak = eng.microbench_int8() if rank == 0 else None

        # ---- N = 1 only: strong-scaling anchor (the fixed 1e6-set sweep on one GPU) + other configs
        strong = None
        extra = None
        if world == 1 and not args.no_extra:
            del prices
            torch.cuda.empty_cache()
            Pg = args.sets_per_gpu * 8
            prm_g = torch.tensor(CASES.heston_batch(Pg), device=dev)
            loss_g = torch.empty(Pg, dtype=torch.float64, device=dev)
            prices_g = torch.empty((Pg, M, J), dtype=torch.float64, device=dev)

            def big():
                eng.cos_loss(F.GAUSS, F.TC_CIR, prm_g, K, S0, RATE, T, NUM_U, mkt_d, w_d, out=loss_g,
                             prices_out=prices_g)
            big()
            big()
            torch.cuda.synchronize(dev)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                big()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms_g = e0.elapsed_time(e1) / 5
            strong = {"what": f"the fixed {Pg}-set sweep (the 8-GPU job) on ONE GPU, device resident: the "
                              "strong-scaling anchor for the per-N weak-scaling values",
                      "global_sets": Pg, "ms_per_step": ms_g, "value": Pg * M * COUNTED_STRIKES / (ms_g * 1e-3),
                      "unit": UNIT}
            del prm_g, loss_g, prices_g
            torch.cuda.empty_cache()
            extra = extra_configs(eng, torch, dev, max(peaks[0], peaks[1]), stream)

    # max over ranks
    sus_ms = sustained[1] if sustained else 0.0
    t = torch.tensor([ms, e2e_s * 1e3, float(launches), sus_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_ms, launches, sus_ms = float(tmax[0]), float(tmax[1]), int(tsum[2]), float(tmax[3])
    else:
        ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        prices_per_step = world * P * M * COUNTED_STRIKES
        value = prices_per_step * args.steps / (ms * 1e-3)
        e2e_value = prices_per_step * args.steps / (e2e_ms * 1e-3)
        rows = P * M
        gemm_ms, gemm_n = prof.get("cos_gemm_kernel", (0.0, 0))
        coef_ms, coef_n = prof.get("cos_coeff_kernel", (0.0, 0))
        peak = max(peaks[0], peaks[1])
        step_ms = ms / args.steps
        step_kernel_ms = max(sum(v[0] for v in prof.values()), 1e-9)
        ncu = load_ncu_summary()
        pipe_note = ("fp64 pipe, 64 FMA/clk/SM: mma.sync.m8n8k4.f64 (DMMA) and DFMA share it; "
                     "tcgen05/TMEM have no f64 kind")
        peak_note = ("fp64 throughput measured live in this process (fop_microbench_fp64: DFMA %.1f, "
                     "DMMA m8n8k4 %.1f TFLOP/s); MEASURED_PEAKS.json has no fp64 entry"
                     % (peaks[0], peaks[1]))
        # SURVEY 8d algorithmic flops of the step: 4 numU per evaluated price (the contraction); the
        # characteristic-function evaluations carry none by that definition and are reported through
        # their executed fp64 flop count (ncu instruction counts, committed under profiles/)
        step_flops = rows * J * 4.0 * NUM_U
        evals_per_step = P * M * NUM_U
        coef_flops = ncu.get("cos_coeff_kernel", {}).get("fp64_flops_per_cf_eval")
        gemm_launches_per_step = max(gemm_n / args.steps, 1)
        coef_launches_per_step = max(coef_n / args.steps, 1)
        gemm_avg_s = gemm_ms * 1e-3 / max(gemm_n, 1)
        coef_avg_s = coef_ms * 1e-3 / max(coef_n, 1)
        gemm_ach = step_flops / gemm_launches_per_step / gemm_avg_s / 1e12 if gemm_n else None
        coef_ach = (coef_flops * evals_per_step / coef_launches_per_step / coef_avg_s / 1e12
                    if coef_flops and coef_n else None)
        if contraction == "int8-tcgen05":
            # int8 digit-plane contraction (cos_gemm_i8.cuh): per launch it streams the 6 coefficient
            # planes once from HBM (rows x 2 numU x 6 bytes -- the dominant, unavoidable traffic) and
            # issues 12 N=144 + 2 N=80 tcgen05.mma (M 128, K 32) per 32 kk.  Bound: HBM (0.47 ms of
            # traffic against 0.43 ms of int8 tensor work at the nominal 4.5 POP/s).
            hbm = measured_peaks().get("hbm_gbs", 6554.6)
            alg_bytes = rows * (2.0 * NUM_U * 6 + J * 8 + 2 * 8) / gemm_launches_per_step   # planes in, prices + row losses out
            i8_ops = rows * 2.0 * NUM_U * 2.0 * (12 * 144 + 2 * 80) / gemm_launches_per_step
            gemm_entry = {
                "kernel": "cos_gemm_kernel (cos_gemm_i8_kernel)", "bound": "hbm",
                "pipe": "tcgen05.mma kind::i8 (exact int8 x int8 -> int32 digit-plane products, accumulators in "
                        "TMEM) fed by TMA; the kernel streams the coefficient digit planes once",
                "achieved": alg_bytes / gemm_avg_s / 1e9 if gemm_n else None, "peak": hbm, "unit": "GB/s",
                "frac": (alg_bytes / gemm_avg_s / 1e9 / hbm) if gemm_n else None,
                "traffic": ncu.get("cos_gemm_kernel", {}).get("dram_bytes"),
                "algorithmic_bytes_per_launch": alg_bytes,
                "int8_tensor": {"ops_per_launch": i8_ops, "achieved_pops": i8_ops / gemm_avg_s / 1e15 if gemm_n else None,
                                "measured_peak_pops": i8_peak, "nominal_peak_pops": 4.5,
                                "frac_of_measured": (i8_ops / gemm_avg_s / 1e15 / i8_peak) if gemm_n and i8_peak else None,
                                "peak_source": "fop_microbench_int8: tcgen05.mma kind::i8 M128 N256 K32 back to back, "
                                               "operands resident in shared memory, measured live in this process"},
                "fp64_equivalent": {"algorithmic_flops_per_launch": step_flops / gemm_launches_per_step,
                                    "tflops": gemm_ach, "x_fp64_peak": (gemm_ach / peak) if gemm_ach else None,
                                    "what": "SURVEY 8d flops (4 numU per price) over the launch time: the fp64 "
                                            "tensor-core kernel this replaces reached 0.835 of the fp64 peak"},
                "avg_launch_ms": gemm_avg_s * 1e3, "share_of_step": gemm_ms / step_kernel_ms}
        else:
            gemm_entry = {
                "kernel": "cos_gemm_kernel", "bound": "tensor", "pipe": pipe_note, "achieved": gemm_ach,
                "peak": peak, "unit": "TFLOP/s", "frac": (gemm_ach / peak) if gemm_ach else None,
                "traffic": ncu.get("cos_gemm_kernel", {}).get("dram_bytes"),
                "algorithmic_flops_per_launch": step_flops / gemm_launches_per_step,
                "algorithmic_bytes_per_launch": rows * (2.0 * NUM_U * 8 + J * 8) / gemm_launches_per_step,
                "avg_launch_ms": gemm_avg_s * 1e3, "share_of_step": gemm_ms / step_kernel_ms}
        coef_entry = {
            "kernel": "cos_coeff_kernel", "bound": "tensor", "pipe": "fp64 pipe (DFMA/DMUL/DADD), dispatch-bound: "
            "an fp64 warp instruction occupies the dispatch port for two cycles (profiles/r2_coeff_variants.json)",
            "achieved": coef_ach, "peak": peak, "unit": "TFLOP/s",
            "frac": (coef_ach / peak) if coef_ach else None,
            "traffic": ncu.get("cos_coeff_kernel", {}).get("dram_bytes"),
            "what": "executed fp64 flops per CF evaluation (ncu instruction counts of the shipped kernel, "
                    "profiles/) x evaluations per launch / launch time; SURVEY 8d assigns CF evaluation no "
                    "algorithmic flops, so the step-level fraction below is the conservative figure",
            "fp64_flops_per_cf_eval": coef_flops, "cf_evals_per_launch": evals_per_step / coef_launches_per_step,
            "avg_launch_ms": coef_avg_s * 1e3, "share_of_step": coef_ms / step_kernel_ms}
        dominant, other = (coef_entry, gemm_entry) if coef_ms >= gemm_ms else (gemm_entry, coef_entry)
        roofline = dict(dominant)
        roofline["peak_source"] = peak_note
        roofline["whole_step"] = {
            "achieved": step_flops / (step_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
            "frac": step_flops / (step_ms * 1e-3) / 1e12 / peak,
            "what": "SURVEY 8d algorithmic flops of the step (4 numU per evaluated price; CF evaluation not "
                    "counted) over the whole step time, against the fp64 peak (fp64-equivalent when the "
                    "contraction runs as exact int8 digit planes on tcgen05)",
            "contraction": contraction}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if contraction != "int8-tcgen05" else
                     "f64 (characteristic functions, price map, loss) + exact int8 x int8 -> int32 on 46-bit "
                     "fixed-point digit planes (contraction)",
            "data": "synthetic", "config": workload_config(P, world),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(prm_host.numel() * 8 + (J + M + 2 * M * J) * 8),
                    "d2h_bytes_per_step": int(P * 8),
                    "ms_per_step": e2e_ms / args.steps,
                    "what": "fop_cos_loss with host parameter sets in / host losses out"
                            + (" (per rank, host waits for them every step) + fop_gather_losses_async of the "
                               "device-resident losses" if world > 1 else "")},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_other": [other],
            "kernels": {
                k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                for k, v in prof.items()},
            "cf_evals_per_s": (P * M * NUM_U * args.steps / (coef_ms * 1e-3)) if coef_n else None,
            "collective": ("fop_gather_losses_async (the library's NCCL all-gather) on a side stream, overlapped "
                           "with the next step" if world > 1 else None),
        }
        if sustained:
            n_sus = sustained[0]
            line["sustained"] = {"seconds": sus_ms * 1e-3, "steps": n_sus, "ms_per_step": sus_ms / n_sus,
                                 "value": prices_per_step * n_sus / (sus_ms * 1e-3), "unit": UNIT,
                                 "clocks": sustained[2]}
        if all_fp64:
            line["all_fp64"] = {
                "ms_per_step": all_fp64, "value": P * M * COUNTED_STRIKES / (all_fp64 * 1e-3), "unit": UNIT,
                "n_gpus": 1,
                "what": "the same step on this rank with FOP_GEMM_I8=0: the contraction on the fp64 tensor cores "
                        "(mma.sync DMMA) like every other kernel.  The default contraction is exact integer "
                        "arithmetic on 46-bit fixed-point operands; the two agree to <= 1.7e-10 over 66 M prices "
                        "(rms 4e-12) -- less than either differs from the CPU reference (3e-10 = 0.03 of the "
                        "1e-8 + 1e-10 |ref| bar, profiles/r2_parity_sweeps_gpu.log)"}
        if strong:
            line["strong_scaling_anchor"] = strong
        if extra:
            line["extra_configs"] = extra
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_rate(args.cpu_seconds).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    eng.close()
    if world > 1:
        # tear down explicitly and leave without running interpreter finalisers: NCCL / CUDA
        # teardown order at interpreter exit is not ours to control and must not turn a finished
        # measurement into a non-zero exit code
        try:
            dist.destroy_process_group()
        except Exception:
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--sets-per-gpu", type=int, default=SETS_PER_GPU)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustain-seconds", type=float, default=3.0,
                    help="length of the sustained leg (0 disables it)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the N=1-only strong-scaling anchor and the C2/C3/C4 lines")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
