#!/usr/bin/env python
"""Static SASS statistics of one kernel: instruction mix of every loop (backward branch) body.

    cuobjdump -sass x.o | python scripts/sass_mix.py [kernel-name-substring]

Used on this GPU-less build machine as the proxy for the issued-instruction count of the
issue-bound characteristic-function kernels (the per-maturity loop body is straight-line code, so
its static count is its dynamic count)."""
import re
import sys
from collections import Counter

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "DMMA")


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else ""
    kernels, cur = {}, None
    for line in sys.stdin:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur is not None:
            kernels[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in kernels.items():
        if want not in name or not ins:
            continue
        tot = Counter(op_of(t) for _, t in ins)
        print(f"== {name}: {len(ins)} instructions, fp64 {sum(v for k, v in tot.items() if k in FP64)}")
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?0x([0-9a-f]+)\)?", t)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                body = ins[addr_index[tgt]: i + 1]
                if len(body) < 40:
                    continue
                c = Counter(op_of(x) for _, x in body)
                f = sum(v for k, v in c.items() if k in FP64)
                print(f"  loop 0x{tgt:x}..0x{a:x}: {len(body)} instr, fp64 {f} ({100.0 * f / len(body):.0f}%)")
                print("     " + " ".join(f"{k}:{v}" for k, v in c.most_common(24)))


def op_of(text):
    t = text
    if t.startswith("@"):
        t = t.split(None, 1)[1] if " " in t else t
    return t.split()[0].split(".")[0]


if __name__ == "__main__":
    main()
