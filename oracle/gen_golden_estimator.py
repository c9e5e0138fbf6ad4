"""ORACLE / TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/estimator_v1.json from the reference
flavour (optioncal::generateFOEstimate of the reference's own OptionCalibration.h).  Inputs: the
strike list of generateCharts.cpp:225 with call quotes produced by the reference's
FangOostCallPrice for three models (the "synthetic market" recipe of generateCharts.cpp:43-57)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import cases as CASES  # noqa: E402

STRIKES = [95, 100, 130, 150, 160, 165, 170, 175, 185, 190, 195, 200, 210, 240, 250]


def markets():
    return [
        dict(name="heston", base=0, tc=1, params=CASES.heston0(), stock=178.46, r=0.0, T=1.0, mult=3.0, N=4096),
        dict(name="cgmy", base=2, tc=0, params=[.1, .1, 2.4, 10.7, .5], stock=178.0, r=.04, T=1.0, mult=10.0, N=1024),
        dict(name="merton", base=1, tc=0, params=[.2, .5, .05, .05], stock=178.0, r=.04, T=.25, mult=10.0, N=4096),
    ]


def quotes(ora, mk):
    strikes = np.array(STRIKES, dtype=np.float64)
    stock = mk["stock"]
    K = np.concatenate([[stock * 50], strikes[::-1], [stock * .03]])  # generateCharts.cpp:43-48
    prices = ora.cos(mk["base"], mk["tc"], mk["params"], K, stock, mk["r"], [mk["T"]], 256, 0)[0, 0]
    obs = prices[1:-1][::-1]
    maxStrike = strikes[-1] * mk["mult"]
    return strikes, obs, stock / maxStrike, maxStrike


def main():
    ora = O.load("reference")
    u = np.linspace(-20.0, 20.0, 15)
    out = {"generator": "oracle/gen_golden_estimator.py", "u": u.tolist(), "cases": []}
    for mk in markets():
        strikes, obs, minK, maxK = quotes(ora, mk)
        ph = ora.fo_estimate(strikes, obs, mk["stock"], mk["r"], mk["T"], minK, maxK, mk["N"], u)
        out["cases"].append({"name": mk["name"], "strikes": strikes.tolist(), "options": obs.tolist(),
                             "stock": mk["stock"], "r": mk["r"], "T": mk["T"], "minStrike": minK,
                             "maxStrike": maxK, "N": mk["N"], "phiHat_re": ph.real.tolist(),
                             "phiHat_im": ph.imag.tolist()})
    with open(os.path.join(ROOT, "tests", "golden", "estimator_v1.json"), "w") as f:
        json.dump(out, f)
    print("ok", len(out["cases"]))


if __name__ == "__main__":
    main()
