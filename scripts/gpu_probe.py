"""Scratch probe run on the GPU box: fp64 pipe micro-benchmarks + first COS timings."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fftoptionpricing_b200 as F  # noqa: E402
from tests import cases as CASES  # noqa: E402

import torch  # noqa: E402

eng = F.Engine(0)
res = {}
res["microbench_tflops[dfma,dmma884,dmma1688,mixed]"] = eng.microbench_fp64()
print(res, flush=True)


def time_call(fn, reps=5):
    st = torch.cuda.ExternalStream(eng.stream_ptr)
    fn(); eng.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


dev = torch.device("cuda:0")
# C5 shard: P sets x 8 maturities x 66 strikes
for P in (16384, 131072):
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K = CASES.c5_strikes()
    out = torch.empty((P, 8, 66), dtype=torch.float64, device=dev)
    for var in (None, "fused"):
        if var: os.environ["FOP_COS_FUSED"] = "1"
        else: os.environ.pop("FOP_COS_FUSED", None)
        best, med = time_call(lambda: eng.cos(0, 1, prm, K, 100.0, 0.0, CASES.C5_T, 256, 0, out=out))
        res[f"c5 P={P} variant={var}"] = dict(ms=best, ms_med=med, prices_per_s=P * 8 * 64 / (best * 1e-3))
        print(res, flush=True)
os.environ.pop("FOP_COS_FUSED", None)
# C2: P sets x 1024 strikes
for P in (4096, 65536):
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K = CASES.fang_oost_strikes(-5.0, 5.0, 100.0, 1024)
    out = torch.empty((P, 1, 1024), dtype=torch.float64, device=dev)
    for var in (None, "fused"):
        if var: os.environ["FOP_COS_FUSED"] = "1"
        else: os.environ.pop("FOP_COS_FUSED", None)
        best, med = time_call(lambda: eng.cos(0, 1, prm, K, 100.0, 0.0, [1.0], 256, 0, out=out))
        res[f"c2 P={P} variant={var}"] = dict(ms=best, ms_med=med, prices_per_s=P * 1024 / (best * 1e-3),
                                             tflops=P * 1024 * 1024 / (best * 1e-3) / 1e12)
        print(res, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
