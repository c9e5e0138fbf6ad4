"""ORACLE / TEST INFRASTRUCTURE ONLY -- builds the CPU oracles (see Makefile)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(reference: str = "/root/reference", verbose: bool = False) -> dict:
    out = {}
    r = subprocess.run(["make", "-C", HERE, "port"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle port build failed:\n" + r.stdout + r.stderr)
    out["port"] = os.path.join(HERE, "libfftop_oracle_port.so")
    if os.path.isfile(os.path.join(reference, "OptionPricing.h")):
        r = subprocess.run(["make", "-C", HERE, "ref", f"REFERENCE={reference}"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle ref build failed:\n" + r.stdout + r.stderr)
        out["reference"] = os.path.join(HERE, "_ref", "libfftop_oracle_ref.so")
    elif os.path.isfile(os.path.join(HERE, "_ref", "libfftop_oracle_ref.so")):
        out["reference"] = os.path.join(HERE, "_ref", "libfftop_oracle_ref.so")  # prebuilt, travelled
    if verbose:
        print(out)
    return out


if __name__ == "__main__":
    build(verbose=True)
    sys.exit(0)
