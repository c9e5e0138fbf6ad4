"""ORACLE / TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/golden_v1.json.

Runs every case of tests/cases.py through the "reference" oracle flavour
(oracle/_ref/libfftop_oracle_ref.so = the reference's own headers from /root/reference compiled
against oracle/shims) and stores the outputs.  Outputs longer than MAX_KEEP values are
sub-sampled on a fixed stride (indices are stored).  Must be run in the container that has
/root/reference; the JSON travels with the repo, the reference does not.

    python oracle/gen_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import cases as CASES  # noqa: E402

MAX_KEEP = 3072


def keep_indices(n):
    if n <= MAX_KEEP:
        return None
    stride = -(-n // MAX_KEEP)
    idx = np.unique(np.concatenate([np.arange(0, n, stride), [n - 1]]))
    return idx


def main():
    ora = O.load("reference")
    assert ora.flavor == "reference"
    golden = {"generator": "oracle/gen_golden.py", "oracle": "reference headers + oracle/shims",
              "cases": []}
    for case in CASES.catalogue():
        out = CASES.run_case(ora, case)
        idx = keep_indices(out.size)
        vals = out if idx is None else out[idx]
        golden["cases"].append({
            "name": case["name"], "fn": case["fn"], "ref": case["ref"], "n": int(out.size),
            "idx": None if idx is None else [int(i) for i in idx],
            "values": [float(v) for v in vals],
        })
        print(f"{case['name']:40s} n={out.size:7d} kept={len(vals)}")
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.json")
    with open(path, "w") as f:
        json.dump(golden, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
