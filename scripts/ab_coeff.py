#!/usr/bin/env python
"""A/B timing of the COS coefficient-kernel variants on the C5 shard (125 000 Heston sets x 8
maturities, numU 256): per-kernel CUDA-event times through fop_profile_*.

    python scripts/ab_coeff.py [--sets 125000] [--reps 10] [--out gpurun_out/ab_coeff.json]
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
    ap.add_argument("--out", default="gpurun_out/ab_coeff.json")
    ap.add_argument("--variants", default="generic,0,1,2,3,4,5")
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
    loss = torch.empty(P, dtype=torch.float64, device=dev)
    prices = torch.empty((P, len(T), len(K)), dtype=torch.float64, device=dev)
    res = {}
    base = None
    for v in args.variants.split(","):
        os.environ.pop("FOP_COEFF_GENERIC", None)
        os.environ.pop("FOP_COEFF_SHAPE", None)
        if v == "generic":
            os.environ["FOP_COEFF_GENERIC"] = "1"
        else:
            os.environ["FOP_COEFF_SHAPE"] = v
        for _ in range(3):
            eng.cos_loss(0, 1, prm, K, 100.0, 0.0, T, 256, mkt, w, out=loss, prices_out=prices)
        eng.synchronize()
        eng.profile_enable(True)
        for _ in range(args.reps):
            eng.cos_loss(0, 1, prm, K, 100.0, 0.0, T, 256, mkt, w, out=loss, prices_out=prices)
        prof = eng.profile_read()
        eng.profile_enable(False)
        cur = prices.clone()
        if base is None:
            base = cur
        diff = float((cur - base).abs().max())
        res[v] = {k: ms / n for k, (ms, n) in prof.items()}
        res[v]["max_abs_diff_vs_first"] = diff
        print(v, {k: round(x, 4) for k, x in res[v].items()}, flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    eng.close()


if __name__ == "__main__":
    main()
