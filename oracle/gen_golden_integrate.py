"""ORACLE / TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/integrate_v1.json from the reference
flavour: the reference's own integrateFFT / genericIntegration(ifft) (fft.hpp:64-110) applied to
(a) a normal density (its Fourier integral is the Gaussian CF) and (b) a damped oscillation."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def cases():
    out = []
    for name, N, lo, hi, omin in [("density", 256, -6.0, 6.0, -10.0), ("density_big", 4096, -8.0, 8.0, 0.0),
                                  ("osc", 1024, 0.0, 40.0, -3.0)]:
        x = lo + (hi - lo) / (N - 1) * np.arange(N)
        if name.startswith("density"):
            f = np.exp(-0.5 * ((x - 0.3) / 0.8) ** 2) / (0.8 * np.sqrt(2 * np.pi)) + 0j
        else:
            f = np.exp(-0.1 * x) * np.exp(1.7j * x) / (1.0 + x)
        out.append(dict(name=name, N=N, inputMin=lo, inputMax=hi, outputMin=omin, fn=f))
    return out


def main():
    ora = O.load("reference")
    res = {"generator": "oracle/gen_golden_integrate.py", "cases": []}
    for c in cases():
        for inverse in (0, 1):
            y = ora.integrate_fft(c["fn"], c["outputMin"], c["inputMin"], c["inputMax"], inverse)
            idx = np.unique(np.linspace(0, c["N"] - 1, 48).astype(int))
            res["cases"].append({"name": c["name"], "inverse": inverse, "N": c["N"],
                                 "inputMin": c["inputMin"], "inputMax": c["inputMax"],
                                 "outputMin": c["outputMin"], "idx": idx.tolist(),
                                 "re": y.real[idx].tolist(), "im": y.imag[idx].tolist()})
    with open(os.path.join(ROOT, "tests", "golden", "integrate_v1.json"), "w") as f:
        json.dump(res, f)
    print("ok", len(res["cases"]))


if __name__ == "__main__":
    main()
