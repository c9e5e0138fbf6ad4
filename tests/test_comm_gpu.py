"""GPU test (needs >= 2 GPUs, skipped otherwise) of the C-ABI multi-GPU path: parameter sets sharded
over two processes / GPUs, per-set losses all-gathered with fop_gather_losses (NCCL bound at run
time) -- every rank ends with the full loss vector in set order, identical to the single-GPU one."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("P", [64, 37])
def test_c_abi_loss_gather_two_gpus(P, tmp_path):
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    base = str(tmp_path / "comm")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "comm_worker.py"), str(r), "2",
                               str(P), base], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-3000:]
    res = [json.load(open(f"{base}.{r}.json")) for r in range(2)]
    single = np.array(res[0]["single"])
    for r in range(2):
        np.testing.assert_array_equal(np.array(res[r]["full"]), single)


def test_c_abi_loss_gather_world_1_host_and_device_pointers():
    """Single-GPU coverage of the library's own collective: dlopen of libnccl,
    ncclCommInitRank with world = 1, fop_gather_losses through the host-staging path and the
    device-pointer path; the gathered vector is the input."""
    import torch

    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    try:
        eng.comm_init(F.Engine.comm_unique_id(), 0, 1)
        x = np.random.default_rng(3).random(1000)
        np.testing.assert_array_equal(eng.gather_losses(x), x)
        xd = torch.tensor(x, device="cuda:0")
        yd = eng.gather_losses(xd)
        eng.synchronize()
        assert torch.equal(yd, xd)
    finally:
        eng.close()


def test_c_abi_async_gather_world_1_double_buffered():
    """fop_gather_losses_async / fop_comm_wait on one GPU: side-stream gathers of alternating
    buffer pairs while the handle's stream keeps producing new losses."""
    import torch

    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    try:
        eng.comm_init(F.Engine.comm_unique_id(), 0, 1)
        stream = torch.cuda.ExternalStream(eng.stream_ptr, device=torch.device("cuda:0"))
        src = [torch.empty(4096, dtype=torch.float64, device="cuda:0") for _ in range(2)]
        dst = [torch.empty(4096, dtype=torch.float64, device="cuda:0") for _ in range(2)]
        seen = []
        with torch.cuda.stream(stream):
            for i in range(6):
                b = i & 1
                src[b].fill_(float(i))              # "loss kernel" of step i on the handle's stream
                eng.gather_losses_async(src[b], dst[b])
                if i >= 1:
                    seen.append(dst[(i - 1) & 1].clone())   # previous gather is complete by now
            eng.comm_wait()
            seen.append(dst[5 & 1].clone())
        torch.cuda.synchronize()
        for i, t in enumerate(seen):
            assert float(t.min()) == float(t.max()) == float(i), (i, float(t.min()), float(t.max()))
    finally:
        eng.close()
