#!/usr/bin/env python
"""One small fixed workload per kernel family, for ncu (scripts/ncu_all.py drives it):
    python scripts/prof_workloads.py <name> [reps]
names: c5 c2 c3pair c3single c4 c4put c1 objfn cuckoo fft fstslarge samples"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fftoptionpricing_b200 as F  # noqa: E402
from tests import cases as CASES  # noqa: E402

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = F.Engine(0)
dev = torch.device("cuda:0")


def run(fn):
    for _ in range(reps):
        fn()
    eng.synchronize()


if name == "c5":
    P = 16384
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K, T = CASES.c5_strikes(), CASES.C5_T
    mkt = torch.tensor(np.array([[CASES.bs_call(100.0, 1.0, k, 0.25, t) for k in K] for t in T]), device=dev)
    w = torch.tensor(CASES.c5_weights(), device=dev)
    loss = torch.empty(P, dtype=torch.float64, device=dev)
    out = torch.empty((P, 8, 66), dtype=torch.float64, device=dev)
    run(lambda: eng.cos_loss(0, 1, prm, K, 100.0, 0.0, T, 256, mkt, w, out=loss, prices_out=out))
elif name == "c2":
    P = 2048
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K = CASES.fang_oost_strikes(-5.0, 5.0, 100.0, 1024)
    out = torch.empty((P, 1, 1024), dtype=torch.float64, device=dev)
    run(lambda: eng.cos(0, 1, prm, K, 100.0, 0.0, [1.0], 256, 0, out=out))
elif name in ("c3pair", "c3single"):
    if name == "c3single":
        os.environ["FOP_FSTS_PAIR"] = "0"
    P = 592
    prm = torch.tensor(CASES.merton_batch(P), device=dev)
    out = torch.empty((P, 4096), dtype=torch.float64, device=dev)
    run(lambda: eng.fsts(1, 0, prm, 4096, 7.5, 100.0, .05, .25, steps=256, american=True, payoff=1, out=out))
elif name in ("c4", "c4put"):
    P = 2048
    prm = torch.tensor(CASES.cgmy_batch(P), device=dev)
    out = torch.empty((P, 65536), dtype=torch.float64, device=dev)
    run(lambda: eng.carr_madan(2, 0, prm, 65536, .25, 1.5, 500.0, .4, .25, name == "c4put", out=out))
elif name == "c1":
    run(lambda: eng.carr_madan(0, 0, [.3], 1024, .25, 1.5, 50.0, .05, 1.0))
elif name == "objfn":
    P = 200000
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    u = np.linspace(-20, 20, 15)
    ph = np.exp(-0.02 * u * u) * (1 + 0.1j)
    out = torch.empty(P, dtype=torch.float64, device=dev)
    run(lambda: eng.objfn_cf(0, 1, prm, u, ph, 0.0, 1.0, out=out))
elif name == "cuckoo":
    u = np.linspace(-20, 20, 15)
    ph = np.exp(-0.02 * u * u) * (1 + 0.1j)
    run(lambda: eng.calibrate_cf(0, 1, [.05, .1, .1, -.95, .05], [.6, 2, 4, .95, 2], u, ph, 0.0, 1.0, nests=25,
                                 iterations=300, restarts=148))
elif name == "fft":
    x = torch.randn((512, 65536), dtype=torch.complex128, device=dev)
    run(lambda: eng.fft(x))
    y = torch.randn((16384, 4096), dtype=torch.complex128, device=dev)
    run(lambda: eng.fft(y))
elif name == "fstslarge":
    prm = torch.tensor(CASES.merton_batch(64), device=dev)
    out = torch.empty((64, 32768), dtype=torch.float64, device=dev)
    run(lambda: eng.fsts(1, 0, prm, 32768, 7.5, 100.0, .05, .25, steps=2, american=True, payoff=1, out=out))
elif name == "samples":
    K, T = CASES.c5_strikes(), CASES.C5_T
    smp = np.exp(-0.01 * np.arange(256)[None, None, :] * np.ones((512, 8, 1))) * (1 + 0j)
    run(lambda: eng.cos_cf_samples(smp, K, 100.0, 0.0, T, 256, 0))
else:
    raise SystemExit("unknown workload " + name)
print("ok", name)
