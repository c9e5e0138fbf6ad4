"""Mirror of namespace ``optioncal``'s hot function (reference OptionCalibration.h:185-201) and the
price-space sweep of BASELINE.json config 5, plus the multi-GPU sharding of parameter sets.

Parameter sets are independent, so N ranks each price a contiguous slice and the only exchange
is an all-gather of the per-set losses (torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import engine as E
from .models import CF

largeNumber = 5000.0  # OptionCalibration.h:185


def getObjFn_arr(phiHat, model: CF, uArray, r: float, T: float, engine: Optional[E.Engine] = None):
    """Returns ``objFn(params)``: for a batch of parameter vectors [P][np] the vector of
    mean_l |phiHat_l - logCF_theta(1 + i u_l)|^2 with NaN -> 5000 (OptionCalibration.h:186-201).
    `model.params` is ignored (only the model family is used)."""
    phiHat = np.ascontiguousarray(np.asarray(phiHat, dtype=np.complex128))
    uArray = np.ascontiguousarray(np.asarray(uArray, dtype=np.float64))

    def objFn(params):
        from .pricing import default_engine
        eng = engine or default_engine()
        return eng.objfn_cf(model.base, model.time_change, params, uArray, phiHat, r, T)

    objFn.model, objFn.phiHat, objFn.uArray, objFn.r, objFn.T, objFn.engine = model, phiHat, uArray, r, T, engine
    return objFn


def cuckoo_optimize(objFn, constraints, nestSize=25, numSims=1500, tol=1e-12, seed=42, restarts=1):
    """Mirror of the reference's ``cuckoo::optimize(objFn, constraints, nestSize, numSims, tol, seed)``
    call (generateCharts.cpp:115) for an ``objFn`` made by :func:`getObjFn_arr`: the whole search
    runs on the device (fop_calibrate_cf; ``restarts`` independent searches, one CTA each).
    ``constraints``: sequence of (lower, upper).  Returns ``(params, fnval)`` of the best restart."""
    from .pricing import default_engine
    eng = objFn.engine or default_engine()
    lo = [c[0] for c in constraints]
    hi = [c[1] for c in constraints]
    prm, fv, _, best = eng.calibrate_cf(objFn.model.base, objFn.model.time_change, lo, hi, objFn.uArray,
                                        objFn.phiHat, objFn.r, objFn.T, nests=nestSize,
                                        iterations=numSims, tol=tol, seed=seed, restarts=restarts)
    return prm[best], float(fv[best])


def generateFOEstimate(strikes, options, stock, r, t, minStrike, maxStrike):
    """Mirror of optioncal::generateFOEstimate (OptionCalibration.h:157-184): returns
    ``estimate(N, uArray) -> phiHat`` (complex128 [len(uArray)]).  Host computation (spline +
    Simpson DFT), done once per market snapshot; feed the result to getObjFn_arr."""
    import ctypes as C

    from . import _lib
    L = _lib.load_library()
    k = np.ascontiguousarray(np.asarray(strikes, dtype=np.float64))
    o = np.ascontiguousarray(np.asarray(options, dtype=np.float64))

    def estimate(N, uArray):
        u = np.ascontiguousarray(np.asarray(uArray, dtype=np.float64))
        out = np.empty(2 * u.size)
        st = L.fop_fo_estimate(k.ctypes.data_as(_lib._dp), o.ctypes.data_as(_lib._dp), k.size,
                               float(stock), float(r), float(t), float(minStrike), float(maxStrike),
                               int(N), u.ctypes.data_as(_lib._dp), u.size, out.ctypes.data_as(_lib._dp))
        if st != 0:
            raise E.FopError(st, "fop_fo_estimate: invalid arguments")
        return out.view(np.complex128).copy()

    return estimate


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of `total` parameter sets owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_losses(params: np.ndarray, compute: Callable[[np.ndarray], "np.ndarray"],
                   device=None):
    """Evaluates `compute` (e.g. a bound Engine.cos_loss) on this rank's slice of `params` and
    all-gathers the per-set losses; every rank returns the full [P] vector in set order.
    Works with any initialised torch.distributed backend; without one it is a plain call."""
    import torch
    import torch.distributed as dist
    P = params.shape[0]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(compute(params))
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(P, rank, world)
    local = compute(params[lo:hi])
    cap = -(-P // world)  # all_gather needs equal sizes: pad to the largest shard
    dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    buf = torch.zeros(cap, dtype=torch.float64, device=dev)
    buf[: hi - lo] = torch.as_tensor(np.asarray(local), dtype=torch.float64, device=dev)
    out = torch.empty(cap * world, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf) if dev != "cpu" else dist.all_gather(
        list(out.view(world, cap).unbind(0)), buf)
    out = out.view(world, cap).cpu().numpy()
    return np.concatenate([out[r, : shard_bounds(P, r, world)[1] - shard_bounds(P, r, world)[0]]
                           for r in range(world)])
