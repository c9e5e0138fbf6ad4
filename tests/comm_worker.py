"""Worker for tests/test_comm_gpu.py: one process per GPU, C-ABI NCCL gather of the per-set losses
(fop_comm_* / fop_gather_losses), no torch.distributed involved."""
import json
import os
import sys
import time

import numpy as np
import torch  # before the library binds NCCL: torch must load ITS bundled libnccl.so.2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fftoptionpricing_b200 as F  # noqa: E402
from fftoptionpricing_b200 import calibration  # noqa: E402
from tests import cases as CASES  # noqa: E402


def main():
    rank, world, P, base = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    idfile = base + ".id"
    if rank == 0:
        uid = F.Engine.comm_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise RuntimeError("no NCCL id from rank 0")
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    eng = F.Engine(rank)
    eng.comm_init(uid, rank, world)
    prm = CASES.heston_batch(P)
    K, T, w = CASES.c5_strikes(), CASES.C5_T, CASES.c5_weights()
    mkt = np.array([[CASES.bs_call(100.0, 1.0, k, 0.25, t) for k in K] for t in T])
    cap = -(-P // world)                                   # equal counts: pad the last shard
    lo, hi = calibration.shard_bounds(P, rank, world)
    local = np.zeros(cap)
    local[: hi - lo] = eng.cos_loss(F.GAUSS, F.TC_CIR, prm[lo:hi], K, 100.0, 0.0, T, 256, mkt, w)
    gathered = eng.gather_losses(local).reshape(world, cap)
    full = np.concatenate([gathered[r, : calibration.shard_bounds(P, r, world)[1] - calibration.shard_bounds(P, r, world)[0]]
                           for r in range(world)])
    # overlapped form: device buffers, side stream, two alternating buffer pairs
    dev = torch.device(f"cuda:{rank}")
    lt = [torch.tensor(local, device=dev) * (k + 1) for k in range(2)]
    gt = [torch.empty(world * cap, dtype=torch.float64, device=dev) for _ in range(2)]
    torch.cuda.synchronize(dev)      # the inputs were produced on torch's stream, the gather runs on the handle's
    for k in range(4):
        eng.gather_losses_async(lt[k & 1], gt[k & 1])
    eng.comm_wait()
    eng.synchronize()
    for k in range(2):
        assert np.array_equal(gt[k].cpu().numpy().reshape(world, cap), gathered * (k + 1)), "async gather mismatch"
    out = {"rank": rank, "full": full.tolist()}
    if rank == 0:
        out["single"] = eng.cos_loss(F.GAUSS, F.TC_CIR, prm, K, 100.0, 0.0, T, 256, mkt, w).tolist()
    with open(f"{base}.{rank}.json", "w") as f:
        json.dump(out, f)
    eng.close()


if __name__ == "__main__":
    main()
