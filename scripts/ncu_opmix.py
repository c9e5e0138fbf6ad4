#!/usr/bin/env python
"""Executed-instruction mix + stall breakdown of every kernel in an .ncu-rep (source + raw pages).
    python scripts/ncu_opmix.py rep.ncu-rep [kernel-regex] [out.json]"""
import csv
import json
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_summary import opcode_mix, summarise  # noqa: E402


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, v in zip(hdr, vals):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
                try:
                    d[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = int(float(v))
                except ValueError:
                    pass
        tot = max(sum(d.values()), 1)
        res.append({k: round(100.0 * v / tot, 1) for k, v in sorted(d.items(), key=lambda kv: -kv[1]) if v * 100 > tot})
    return res


if __name__ == "__main__":
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] else None
    summ = summarise(rep)
    mix = opcode_mix(rep, rx)
    st = stalls(rep)
    out = []
    for i, m in enumerate(mix):
        e = dict(summ[i]) if len(summ) == len(mix) else {}
        e.update(m)
        if len(st) == len(mix):
            e["warp_stall_pct"] = st[i]
        out.append(e)
    if len(sys.argv) > 3:
        json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out, indent=1))
