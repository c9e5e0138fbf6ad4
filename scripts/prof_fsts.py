"""Small fixed workload for ncu: config 3 (Merton American put, N = 4096, 256 steps) at P = 592."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fftoptionpricing_b200 as F  # noqa: E402
from tests import cases as CASES  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 592
eng = F.Engine(0)
dev = torch.device("cuda:0")
prm = torch.tensor(CASES.merton_batch(P), device=dev)
out = torch.empty((P, 4096), dtype=torch.float64, device=dev)
for _ in range(2):
    eng.fsts(1, 0, prm, 4096, 7.5, 100.0, .05, .25, steps=256, american=True, payoff=1, out=out)
eng.synchronize()
print("ok", float(out[0, 2048]))
