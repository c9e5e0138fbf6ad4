"""Worker for tests/test_dist_gloo.py: world_size-2 gloo run of the sharded loss sweep.  The local
"compute" is the CPU oracle (test infrastructure) because the container has no GPU; what is under
test is the host-side partition + gather logic of fftoptionpricing_b200.calibration."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fftoptionpricing_b200 import calibration  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import cases as CASES  # noqa: E402


def main():
    dist.init_process_group(backend="gloo")
    rank = dist.get_rank()
    ora = O.load("port")
    P = int(sys.argv[1])
    prm = CASES.heston_batch(P)
    K, T = CASES.c5_strikes(), CASES.C5_T[:2]
    mkt = ora.cos(0, 1, prm[:1], K, 100.0, 0.0, T, 64, 0)[0] * 1.01
    w = CASES.c5_weights()[:2]
    calls = []

    def compute(p):
        calls.append(p.shape[0])
        return ora.cos_loss(0, 1, p, K, 100.0, 0.0, T, 64, mkt, w)

    full = calibration.sharded_losses(prm, compute)
    out = {"rank": rank, "local_sets": calls, "losses": full.tolist()}
    with open(sys.argv[2] + f".{rank}", "w") as f:
        json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
