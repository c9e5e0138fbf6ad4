"""CPU test of the N > 1 path: parameter sets are sharded contiguously, priced independently and
only the per-set losses are gathered (gloo, world_size 2)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from fftoptionpricing_b200 import calibration
from tests import cases as CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    for total in (0, 1, 7, 8, 125000, 1000003):
        for world in (1, 2, 3, 8):
            edges = [calibration.shard_bounds(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            sizes = [hi - lo for lo, hi in edges]
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("P", [7, 16])
def test_sharded_losses_gloo_world2(P, tmp_path, oracles):
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + P),
           os.path.join(ROOT, "tests", "dist_worker.py"), str(P), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300,
                       env={**os.environ, "OMP_NUM_THREADS": "2"})
    assert r.returncode == 0, r.stderr[-3000:]
    res = [json.load(open(f"{out}.{k}")) for k in range(2)]
    ora = oracles["port"]
    prm = CASES.heston_batch(P)
    K, T = CASES.c5_strikes(), CASES.C5_T[:2]
    mkt = ora.cos(0, 1, prm[:1], K, 100.0, 0.0, T, 64, 0)[0] * 1.01
    ref = ora.cos_loss(0, 1, prm, K, 100.0, 0.0, T, 64, mkt, CASES.c5_weights()[:2])
    for k in range(2):
        lo, hi = calibration.shard_bounds(P, k, 2)
        assert res[k]["local_sets"] == [hi - lo]           # each rank priced only its own slice
        np.testing.assert_array_equal(np.array(res[k]["losses"]), ref)  # full vector, set order
