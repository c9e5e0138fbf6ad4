"""Times every BASELINE.json config on one GPU (device-resident inputs/outputs, CUDA events via
the library's per-kernel profiling) and writes gpurun_out/configs.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fftoptionpricing_b200 as F  # noqa: E402
from tests import cases as CASES  # noqa: E402

eng = F.Engine(0)
dev = torch.device("cuda:0")
peaks = eng.microbench_fp64()
peak = max(peaks[:2])
res = {"fp64_peak_tflops": {"dfma": peaks[0], "dmma_m8n8k4": peaks[1], "dmma_m16n8k8": peaks[2], "mixed": peaks[3]}}
which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]


WALL = {}
_stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)


def timed(fn, reps=3):
    """per-kernel ms (profiled pass); the wall ms per call (CUDA events on the handle's stream) goes to
    WALL['last'] -- kernels of one call may overlap on side streams"""
    fn(); eng.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(_stream)
    for _ in range(reps):
        fn()
    e1.record(_stream)
    torch.cuda.synchronize(dev)
    WALL["last"] = e0.elapsed_time(e1) / reps
    eng.profile_enable(True)
    for _ in range(reps):
        fn()
    prof = eng.profile_read()
    eng.profile_enable(False)
    return {k: v[0] / reps for k, v in prof.items()}


if "c1" in which:  # latency case: one BS Carr-Madan transform, host in / host out
    import time
    eng.carr_madan(0, 0, [.3], 1024, .25, 1.5, 50.0, .05, 1.0)
    t0 = time.perf_counter()
    for _ in range(200):
        eng.carr_madan(0, 0, [.3], 1024, .25, 1.5, 50.0, .05, 1.0)
    dt = (time.perf_counter() - t0) / 200
    res["c1"] = {"what": "BS Carr-Madan N=2^10, one set, host->host call latency", "ms_per_call": dt * 1e3,
                 "prices_per_s": 1024 / dt}
if "c2" in which:
    P = 65536
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K = CASES.fang_oost_strikes(-5.0, 5.0, 100.0, 1024)
    out = torch.empty((P, 1, 1024), dtype=torch.float64, device=dev)
    k = timed(lambda: eng.cos(0, 1, prm, K, 100.0, 0.0, [1.0], 256, 0, out=out))
    ms = sum(k.values())
    res["c2"] = {"what": f"Heston COS 256 u x 1024 strikes, P={P}", "kernels_ms": k, "ms": ms,
                 "prices_per_s": P * 1024 / (ms * 1e-3),
                 "gemm_tflops": P * 1024 * 1024 / (k["cos_gemm_kernel"] * 1e-3) / 1e12,
                 "gemm_frac_of_fp64_peak": P * 1024 * 1024 / (k["cos_gemm_kernel"] * 1e-3) / 1e12 / peak}
if "c3" in which:
    P = 4096
    prm = torch.tensor(CASES.merton_batch(P), device=dev)
    out = torch.empty((P, 4096), dtype=torch.float64, device=dev)
    k = timed(lambda: eng.fsts(1, 0, prm, 4096, 7.5, 100.0, .05, .25, steps=256, american=True, payoff=1, out=out), reps=2)
    ms = sum(k.values())
    flops = P * (256 * (2 * 5 * 4096 * 12 + 8 * 4096))
    res["c3"] = {"what": f"Merton American put FSTS N=4096 x 256 steps, P={P}", "kernels_ms": k, "ms": ms,
                 "prices_per_s": P * 4096 / (ms * 1e-3), "tflops_5NlogN": flops / (ms * 1e-3) / 1e12,
                 "frac_of_fp64_peak": flops / (ms * 1e-3) / 1e12 / peak}
if "c4" in which:
    P = 16384
    prm = torch.tensor(CASES.cgmy_batch(P), device=dev)
    out = torch.empty((P, 65536), dtype=torch.float64, device=dev)
    k = timed(lambda: eng.carr_madan(2, 0, prm, 65536, .25, 1.5, 500.0, .4, .25, False, out=out), reps=2)
    ms = WALL["last"]
    res["c4"] = {"what": f"CGMY Carr-Madan N=65536, P={P}", "kernels_ms": k, "ms": ms, "sum_of_kernels_ms": sum(k.values()),
                 "prices_per_s": P * 65536 / (ms * 1e-3), "cf_evals_per_s": P * 65536 / (k.get("carr_madan_cols_kernel", ms) * 1e-3),
                 "fft_tflops_5NlogN": P * 5 * 65536 * 16 / (ms * 1e-3) / 1e12}
    del out
if "c5" in which:
    P = 125000
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K, T = CASES.c5_strikes(), CASES.C5_T
    out = torch.empty((P, 8, 66), dtype=torch.float64, device=dev)
    k = timed(lambda: eng.cos(0, 1, prm, K, 100.0, 0.0, T, 256, 0, out=out))
    ms = sum(k.values())
    res["c5"] = {"what": f"Heston COS sweep shard P={P} x 8 x 66", "kernels_ms": k, "ms": ms,
                 "prices_per_s": P * 8 * 64 / (ms * 1e-3)}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("FOP_CONFIGS_OUT", "configs.json")), "w"), indent=1)
print(json.dumps(res, indent=1))
