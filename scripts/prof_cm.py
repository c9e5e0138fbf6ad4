"""Small fixed workload for ncu: one fop_carr_madan call, CGMY, N = 65536 (config 4 at P = 1024)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fftoptionpricing_b200 as F  # noqa: E402
from tests import cases as CASES  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = F.Engine(0)
dev = torch.device("cuda:0")
prm = torch.tensor(CASES.cgmy_batch(P), device=dev)
out = torch.empty((P, 65536), dtype=torch.float64, device=dev)
for _ in range(2):
    eng.carr_madan(2, 0, prm, 65536, .25, 1.5, 500.0, .4, .25, False, out=out)
eng.synchronize()
print("ok", float(out[0, 32768]))
