"""Small fixed workload for ncu: one fop_cos call (C5 shard by default, or C2 with `c2`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fftoptionpricing_b200 as F  # noqa: E402
from tests import cases as CASES  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = F.Engine(0)
dev = torch.device("cuda:0")
if which == "c5":
    P = 16384
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K, T = CASES.c5_strikes(), CASES.C5_T
    out = torch.empty((P, 8, 66), dtype=torch.float64, device=dev)
else:
    P = 2048
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K, T = CASES.fang_oost_strikes(-5.0, 5.0, 100.0, 1024), [1.0]
    out = torch.empty((P, 1, 1024), dtype=torch.float64, device=dev)
for _ in range(reps):
    eng.cos(0, 1, prm, K, 100.0, 0.0, T, 256, 0, out=out)
eng.synchronize()
print("ok", float(out[0, 0, 1]))
