#!/usr/bin/env python
"""A/B of the two COS contractions on the C5 workload: fp64 tensor cores (DMMA, FOP_GEMM_I8=0)
against the int8 digit-plane contraction on tcgen05 (cos_gemm_i8.cuh).  Prints the difference of the
prices / losses and per-kernel CUDA-event times.

    python scripts/ab_i8.py [--sets 125000] [--reps 10] [--no-prices]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cases as CASES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=125000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-prices", action="store_true")
    ap.add_argument("--out", default="gpurun_out/ab_i8.json")
    args = ap.parse_args()
    import torch

    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    dev = torch.device("cuda:0")
    P = args.sets
    K, T = CASES.c5_strikes(), CASES.C5_T
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    mkt = torch.tensor(np.array([[CASES.bs_call(100.0, 1.0, k, 0.25, t) for k in K] for t in T]), device=dev)
    w = torch.tensor(CASES.c5_weights(), device=dev)
    res = {}
    keep = {}
    for v in ("0", "1"):
        os.environ["FOP_GEMM_I8"] = v
        loss = torch.zeros(P, dtype=torch.float64, device=dev)
        prices = None if args.no_prices else torch.zeros((P, len(T), len(K)), dtype=torch.float64, device=dev)
        for _ in range(3):
            eng.cos_loss(0, 1, prm, K, 100.0, 0.0, T, 256, mkt, w, out=loss, prices_out=prices)
        eng.synchronize()
        eng.profile_enable(True)
        for _ in range(args.reps):
            eng.cos_loss(0, 1, prm, K, 100.0, 0.0, T, 256, mkt, w, out=loss, prices_out=prices)
        prof = eng.profile_read()
        eng.profile_enable(False)
        keep[v] = (loss.clone(), None if prices is None else prices.clone())
        res[v] = {k: ms / n for k, (ms, n) in prof.items()}
        print("FOP_GEMM_I8=" + v, {k: round(x, 4) for k, x in res[v].items()}, flush=True)
    l0, p0 = keep["0"]
    l1, p1 = keep["1"]
    res["loss_max_rel_diff"] = float(((l1 - l0).abs() / l0.abs().clamp_min(1e-300)).max())
    print("loss: max rel diff", res["loss_max_rel_diff"], "nan", int(torch.isnan(l1).sum()), int(torch.isnan(l0).sum()))
    if p0 is not None:
        d = (p1 - p0).abs()
        res["price_max_abs_diff"] = float(d.max())
        res["price_rms_diff"] = float((d * d).mean().sqrt())
        i = int(d.argmax())
        print("prices: max abs diff", res["price_max_abs_diff"], "rms", res["price_rms_diff"], "at", i,
              float(p0.flatten()[i]), float(p1.flatten()[i]))
        print("sample fp64:", p0[0, 0, :4].tolist(), "\nsample i8  :", p1[0, 0, :4].tolist())
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    eng.close()


if __name__ == "__main__":
    main()
