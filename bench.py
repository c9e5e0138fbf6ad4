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
the per-set losses are all-gathered over NCCL.  A price counts once per (set, maturity, quoted
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
def run_ours(args):
    import torch
    import torch.distributed as dist

    import fftoptionpricing_b200 as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    eng = F.Engine(local)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)

    P = args.sets_per_gpu
    K, T = CASES.c5_strikes(S0), CASES.C5_T
    M, J = len(T), len(K)
    mkt, w = market_and_weights(K, T)
    # rank r prices the r-th contiguous slice of the global sweep (seeded synthetic inputs)
    prm_host = torch.from_numpy(CASES.heston_batch(P, seed=CASES.SEED + 1000 * rank)).pin_memory()
    loss_host = torch.empty(P, dtype=torch.float64).pin_memory()
    prm_dev = prm_host.to(dev)
    prices = torch.empty((P, M, J), dtype=torch.float64, device=dev)
    loss = torch.empty(P, dtype=torch.float64, device=dev)
    gathered = torch.empty(P * world, dtype=torch.float64, device=dev) if world > 1 else None
    mkt_d = torch.from_numpy(mkt).to(dev)
    w_d = torch.from_numpy(w).to(dev)

    def step_device():
        eng.cos_loss(F.GAUSS, F.TC_CIR, prm_dev, K, S0, RATE, T, NUM_U, mkt_d, w_d, out=loss,
                     prices_out=prices)
        if world > 1:
            dist.all_gather_into_tensor(gathered, loss)

    def step_host():
        # host buffers in, host losses out: H2D of the parameter sets + D2H of the losses happen
        # inside the call (which returns after the stream is idle)
        # inside the call (which returns after the stream is idle); prices are stored to HBM exactly
        # as in the device-resident step
        eng.cos_loss(F.GAUSS, F.TC_CIR, prm_host.numpy(), K, S0, RATE, T, NUM_U, mkt, w,
                     out=loss_host.numpy(), prices_out=prices)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step_device()
        barrier()
        # ---- timed region: device-resident inputs
        sampler = ClockSampler(local, str(getattr(torch.cuda.get_device_properties(local), 'uuid', '') or '')) if rank == 0 else None
        l0 = eng.launch_count
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count - l0

        # ---- end to end through the C ABI with host buffers
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
            if world > 1:
                dist.all_gather_into_tensor(gathered, loss_host.to(dev, non_blocking=True))
        barrier()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None   # sampled over both timed regions

        # ---- per-kernel timing for the roofline (separate pass so events do not perturb `value`)
        eng.profile_enable(True)
        for _ in range(args.steps):
            eng.cos_loss(F.GAUSS, F.TC_CIR, prm_dev, K, S0, RATE, T, NUM_U, mkt_d, w_d, out=loss,
                         prices_out=prices)
        prof = eng.profile_read()
        eng.profile_enable(False)
        peaks = eng.microbench_fp64() if rank == 0 else None

    # max over ranks
    t = torch.tensor([ms, e2e_s * 1e3, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
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
        # algorithmic flops of the contraction: 4*numU per evaluated price
        gemm_flops_per_launch = rows * J * 4.0 * NUM_U / max(gemm_n / args.steps, 1)
        gemm_avg_s = gemm_ms * 1e-3 / max(gemm_n, 1)
        achieved = gemm_flops_per_launch / gemm_avg_s / 1e12 if gemm_n else None
        ncu = load_ncu_summary()
        gemm_traffic = ncu.get("cos_gemm_kernel", {}).get("dram_bytes")
        step_kernel_ms = max(sum(v[0] for v in prof.values()), 1e-9)
        tensor_note = ("fp64 pipe, 64 FMA/clk/SM: mma.sync.m8n8k4.f64 (DMMA) and DFMA share it; "
                       "tcgen05/TMEM have no f64 kind")
        peak_note = ("fp64 throughput measured live in this process (fop_microbench_fp64: DFMA %.1f, "
                     "DMMA m8n8k4 %.1f TFLOP/s); MEASURED_PEAKS.json has no fp64 entry"
                     % (peaks[0], peaks[1]))
        # second hot kernel: CF evaluation has no algorithmic flop count in SURVEY 8d (reported as
        # cf_evals); its fraction is executed fp64 flops (ncu instruction counts per evaluation,
        # committed under profiles/) over the same measured peak
        evals_per_step = P * M * NUM_U
        coef_flops = ncu.get("cos_coeff_kernel", {}).get("fp64_flops_per_cf_eval")
        coef_ach = (coef_flops * evals_per_step * args.steps / (coef_ms * 1e-3) / 1e12
                    if coef_flops and coef_n else None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(P, world),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(prm_host.numel() * 8 + (J + M + 2 * M * J) * 8),
                    "d2h_bytes_per_step": int(P * 8), "ms_per_step": e2e_ms / args.steps,
                    "what": "fop_cos_loss with host parameter sets in / host losses out"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "bound": "tensor", "pipe": tensor_note, "kernel": "cos_gemm_kernel",
                "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "traffic": gemm_traffic,
                "peak_source": peak_note,
                "algorithmic_flops_per_launch": gemm_flops_per_launch,
                "algorithmic_bytes_per_launch": rows * (2.0 * NUM_U * 8 + J * 8) / max(gemm_n / args.steps, 1),
                "avg_launch_ms": gemm_avg_s * 1e3,
                "share_of_step": gemm_ms / step_kernel_ms,
            },
            "roofline_other": [
                {"kernel": "cos_coeff_kernel", "bound": "tensor", "pipe": "fp64 pipe (DFMA/DMUL/DADD)",
                 "achieved": coef_ach, "peak": peak, "unit": "TFLOP/s",
                 "frac": (coef_ach / peak) if coef_ach else None,
                 "what": "executed fp64 flops per CF evaluation (ncu, profiles/) x evaluations / s; "
                         "SURVEY 8d counts no algorithmic flops for CF evaluation",
                 "fp64_flops_per_cf_eval": coef_flops,
                 "avg_launch_ms": coef_ms / max(coef_n, 1), "share_of_step": coef_ms / step_kernel_ms,
                 "traffic": ncu.get("cos_coeff_kernel", {}).get("dram_bytes")},
                {"kernel": "whole step", "bound": "tensor",
                 "achieved": rows * J * 4.0 * NUM_U / (ms / args.steps * 1e-3) / 1e12, "peak": peak,
                 "unit": "TFLOP/s",
                 "frac": rows * J * 4.0 * NUM_U / (ms / args.steps * 1e-3) / 1e12 / peak,
                 "what": "SURVEY 8d algorithmic flops (4 numU per evaluated price, CF evaluation "
                         "not counted) over the whole step time"},
            ],
            "kernels": {
                k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                for k, v in prof.items()},
            "cf_evals_per_s": (P * M * NUM_U * args.steps / (coef_ms * 1e-3)) if coef_n else None,
        }
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
