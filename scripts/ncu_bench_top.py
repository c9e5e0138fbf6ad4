#!/usr/bin/env python
"""End-of-round evidence of the bench command itself (run under gpurun, after `python bench.py` has
exited 0 without ncu):
  1. launch list:   ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv  -> CSV +
                    per-kernel summary (share of the step per kernel);
  2. top kernels:   ncu --set full --import-source on -k regex:"cos_gemm|cos_coeff" (one launch each at
                    bench size) -> JSON with DRAM bytes, opcode mix and the executed fp64 flops per CF
                    evaluation that bench.py's roofline uses.
Outputs go to gpurun_out/ (copy them to profiles/r2_bench_*)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from ncu_opmix import stalls  # noqa: E402
from ncu_summary import opcode_mix, summarise  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
BENCH = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline",
         "--no-extra", "--sustain-seconds", "0"]
SETS = 125000
M, NUMU = 8, 256

# ---- 1. launch list
csv_path = os.path.join(OUT, "r2_bench_launches.csv")
subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "400", "--csv",
                "--log-file", csv_path] + BENCH, capture_output=True, text=True)
per = {}
try:
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].replace("second", "s").replace("usecond", "us"), 1e-6)
        e = per.setdefault(r[ki].split("(")[0], [0, 0.0])
        e[0] += 1
        e[1] += v
    json.dump({k: {"launches": n, "total_ms": t, "avg_ms": t / n} for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1])},
              open(os.path.join(OUT, "r2_bench_launches_summary.json"), "w"), indent=1)
except Exception as ex:  # keep going: the full capture below is the important one
    print("launch list summary failed:", ex)

# ---- 2. one launch of each hot kernel at bench size
rep = os.path.join(OUT, "r2_bench_top")
subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k",
                "regex:cos_gemm|cos_coeff", "-s", "8", "-c", "2", "-f", "-o", rep] + BENCH, capture_output=True, text=True)
summ = summarise(rep + ".ncu-rep")
mix = opcode_mix(rep + ".ncu-rep")
st = stalls(rep + ".ncu-rep")
res = []
for i, e in enumerate(summ):
    e = dict(e)
    if i < len(mix):
        for k in ("warp_instructions", "fp64_flops", "fp64_warp_instructions", "mix_pct", "stall_samples_pct"):
            e[k] = mix[i][k]
    if len(st) == len(summ):
        e["warp_stall_pct"] = st[i]
    if "cos_coeff" in e["kernel"]:
        e["sets_in_launch"] = SETS
        e["cf_evals_in_launch"] = SETS * M * NUMU
        if e.get("fp64_flops"):
            e["fp64_flops_per_cf_eval"] = e["fp64_flops"] / (SETS * M * NUMU)
    else:
        e["rows"] = SETS * M
    res.append(e)
json.dump(res, open(os.path.join(OUT, "r2_bench_top_kernels_ncu.json"), "w"), indent=1)
os.unlink(rep + ".ncu-rep")
print(json.dumps([{k: v for k, v in e.items() if k in ("kernel", "gpu__time_duration.sum", "fp64_flops_per_cf_eval",
                                                       "dram__bytes_read.sum", "dram__bytes_write.sum")} for e in res], indent=1))
