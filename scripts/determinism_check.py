"""Run-to-run determinism of the kernels with cross-warp communication (mbarrier ring, FSTS partner
exchange, L2-merged strided stores, on-device search): every call repeated, outputs compared bit for
bit.  `gpurun -- python scripts/determinism_check.py` (compute-sanitizer is closed on this pool)."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import fftoptionpricing_b200 as F
from tests import cases as CASES
eng = F.Engine(0)
def rep(fn, n=6):
    first = fn()
    for _ in range(n):
        o = fn()
        if not np.array_equal(o, first):
            return False, float(np.max(np.abs(o - first)))
    return True, 0.0
prm = CASES.merton_batch(301)
print("fsts pair", rep(lambda: eng.fsts(1, 0, prm, 4096, 7.5, 100.0, .05, .25, steps=64, american=True, payoff=1)))
print("fsts pair N=256", rep(lambda: eng.fsts(1, 0, prm, 256, 7.5, 100.0, .05, .25, steps=64, american=True, payoff=1)))
hp = CASES.heston_batch(5000)
K, T = CASES.c5_strikes(), CASES.C5_T
print("cos split", rep(lambda: eng.cos(0, 1, hp, K, 100.0, 0.0, T, 256, 0)))
os.environ["FOP_COS_PATH"] = "fused"
print("cos fused", rep(lambda: eng.cos(0, 1, hp, K, 100.0, 0.0, T, 256, 0)))
del os.environ["FOP_COS_PATH"]
cp = CASES.cgmy_batch(200)
print("carr-madan pruned", rep(lambda: eng.carr_madan(2, 0, cp, 65536, .25, 1.5, 500.0, .4, .25), 3))
U = np.linspace(-20, 20, 15)
from oracle import oracle as O
ora = O.load("best")
ph = np.array([ora.cf(0, 1, CASES.heston0(), 0.0, 1.0, complex(1, u))[1] for u in U])
print("calibrate", rep(lambda: eng.calibrate_cf(0, 1, [0,0,0,-1,0], [.6,2,4,1,2], U, ph, 0.0, 1.0, iterations=200, tol=0.0, restarts=64)[1], 3))
