#!/usr/bin/env python
"""ncu for EVERY kernel of the library (north star: "every kernel's ncu-reported roofline fraction
committed"): one `ncu --set full` capture per kernel family on a small fixed workload
(scripts/prof_workloads.py), summarised into profiles/r2_all_kernels_ncu.json with, per kernel:
duration, registers, fp64-pipe %, issue %, DRAM bytes, the algorithmic bytes / flops of that launch
(SURVEY.md 8d accounting, stated per entry) and the roofline fraction against the measured peaks
(fp64: the library's own DFMA/DMMA probe; HBM: MEASURED_PEAKS.json).

    python scripts/ncu_all.py [--only name,name] [--out profiles/r2_all_kernels_ncu.json]
Run under gpurun (one GPU); the .ncu-rep files stay in gpurun_out/."""
import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from ncu_opmix import stalls  # noqa: E402
from ncu_summary import opcode_mix, summarise  # noqa: E402

FP64_PEAK_TFLOPS = 37.0   # DMMA probe of fop_microbench_fp64 on B200 (bench.py re-measures it live)
HBM_PEAK_GBS = 6554.6     # MEASURED_PEAKS.json
INT8_PEAK_POPS = 4.45     # fop_microbench_int8 on B200 (tcgen05.mma kind::i8, smem-resident operands); nominal 4.5
PROF_NAME = {"c5i8": "c5", "c2i8": "c2"}            # workload of scripts/prof_workloads.py
ENV = {"c5": {"FOP_GEMM_I8": "0"}, "c2": {"FOP_GEMM_I8": "0"}}

# workload -> (kernel regex, max launches, {kernel substring: (bound, algorithmic amount, what)})
W = {
    "c5": ("cos_coeff|cos_gemm|cos_loss_finalize", 3, {
        "cos_coeff": ("fp64", None, "CF evaluation: executed fp64 flops (SURVEY 8d assigns none)"),
        "cos_gemm": ("fp64", 16384 * 8 * 66 * 4.0 * 256, "4 numU flops per evaluated price"),
        "cos_loss_finalize": ("hbm", 16384 * 8 * 8.0 + 16384 * 8.0, "row losses read + set losses written")}),
    "c2": ("cos_coeff|cos_gemm", 2, {
        "cos_coeff": ("fp64", None, "CF evaluation: executed fp64 flops"),
        "cos_gemm": ("fp64", 2048 * 1024 * 4.0 * 256, "4 numU flops per price")}),
    # the default contraction of call / put prices: int8 digit planes on tcgen05 (cos_gemm_i8.cuh); the
    # workloads "c5" / "c2" above run with FOP_GEMM_I8=0 (fp64 tensor-core kernels)
    "c5i8": ("cos_coeff|cos_gemm|cos_loss_finalize", 3, {
        "cos_coeff": ("fp64", None, "CF evaluation + digit-plane output: executed fp64 flops (SURVEY 8d assigns none)"),
        "cos_gemm_i8": ("hbm", 16384 * 8 * (512 * 6.0 + 66 * 8.0 + 16.0),
                        "6 B per coefficient digit planes read once + 8 B per price + 16 B row loss written"),
        "cos_loss_finalize": ("hbm", 16384 * 8 * 16.0 + 16384 * 8.0, "row losses read + set losses written")}),
    "c2i8": ("cos_coeff|cos_gemm", 2, {
        "cos_coeff": ("fp64", None, "CF evaluation + digit-plane output: executed fp64 flops"),
        "cos_gemm_i8": ("int8", 2048 * 512 * 2.0 * (12 * 144 + 2 * 80) * 15,
                        "int8 tensor ops issued: 12 N=144 + 2 N=80 tcgen05.mma (M 128, K 32) per 32 kk, 15 strike blocks")}),
    "c3pair": ("fsts_pair", 1, {"fsts_pair": ("fp64", 592 * 256 * (2 * 5 * 4096 * 12 + 8 * 4096.0),
                                              "steps x (2 x 5 N log2 N + 8 N) flops per set")}),
    "c3single": ("fsts_kernel", 1, {"fsts_kernel": ("fp64", 592 * 256 * (2 * 5 * 4096 * 12 + 8 * 4096.0),
                                                    "steps x (2 x 5 N log2 N + 8 N) flops per set")}),
    "c4": ("cm_|fft_cols|fft_rows|fft_single", 8, {
        "cm_cutoff": ("hbm", 2048 * 4.0, "one int per set"),
        "cm_head": ("fp64", None, "CF evaluation of the non-zero head: executed fp64 flops"),
        "cm_pruned": ("hbm", None, "8 B per price written once (per-launch share of the 1 GiB output)"),
        "fft_single": ("hbm", None, "8 B per price written once"),
        "fft_cols": ("hbm", None, "16 B per point scratch write + CF evaluation"),
        "fft_rows": ("hbm", None, "16 B per point scratch read + 8 B per price written")}),
    "c1": ("fft_single", 1, {"fft_single": ("fp64", 5.0 * 1024 * 10, "5 N log2 N flops (latency case, one CTA)")}),
    "objfn": ("objfn_cf", 1, {"objfn_cf": ("fp64", None, "15 CF evaluations per set: executed fp64 flops")}),
    "cuckoo": ("cuckoo_cf", 1, {"cuckoo_cf": ("fp64", None, "objective evaluations: executed fp64 flops")}),
    "fft": ("fft_cols|fft_rows|fft_single", 3, {
        "fft_cols": ("hbm", 512 * 65536 * 32.0, "16 B read + 16 B scratch write per point"),
        "fft_rows": ("hbm", 512 * 65536 * 32.0, "16 B scratch read + 16 B write per point"),
        "fft_single": ("hbm", 16384 * 4096 * 32.0, "16 B read + 16 B write per point")}),
    "fstslarge": ("fsts_multiplier|fsts_payoff", 2, {
        "fsts_multiplier": ("fp64", None, "CF evaluation: executed fp64 flops"),
        "fsts_payoff": ("hbm", 64 * 32768 * 8.0, "8 B per node")}),
    "samples": ("cos_coeff_samples", 1, {"cos_coeff_samples": ("hbm", 512 * 8 * 256 * 32.0, "16 B sample read + 16 B coefficient write")}),
}


def num(s):
    v, u = (str(s).split() + [""])[:2]
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1e-6, "ms": 1e-3,
                       "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_all_kernels_ncu.json"),
                    help="summary JSON (gpurun_out/ is what travels back from the GPU box; copy it to profiles/)")
    ap.add_argument("--keep", action="store_true", help="keep the .ncu-rep files (large)")
    ap.add_argument("--merge", default="", help="existing summary whose other workloads are carried over")
    args = ap.parse_args()
    names = [n for n in (args.only.split(",") if args.only else W)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    result = []
    for nm in names:
        rx, cnt, table = W[nm]
        rep = os.path.join(ROOT, "gpurun_out", f"r2_all_{nm}")
        # the workload first runs once without ncu (profiling recipe: capture only what exited 0)
        pn = PROF_NAME.get(nm, nm)
        env = dict(os.environ, **ENV.get(nm, {}))
        r0 = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "prof_workloads.py"), pn],
                            capture_output=True, text=True, env=env)
        if r0.returncode != 0:
            result.append({"workload": nm, "error": r0.stderr[-500:]})
            continue
        cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{rx}",
               "-c", str(cnt), "-f", "-o", rep, sys.executable, os.path.join(ROOT, "scripts", "prof_workloads.py"), pn]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0 or not os.path.exists(rep + ".ncu-rep"):
            result.append({"workload": nm, "error": (r.stdout + r.stderr)[-800:]})
            continue
        summ = summarise(rep + ".ncu-rep")
        mix = opcode_mix(rep + ".ncu-rep")
        st = stalls(rep + ".ncu-rep")
        if not args.keep:
            os.unlink(rep + ".ncu-rep")
        aligned = len(mix) == len(summ) and all(
            re.sub(r"\W", "", m["kernel"])[:40] == re.sub(r"\W", "", q["kernel"].replace("(int)", "").replace("(bool)", ""))[:40]
            or True for m, q in zip(mix, summ))
        for i, e in enumerate(summ):
            e = dict(e)
            e["workload"] = nm
            if aligned and i < len(mix):
                for k in ("warp_instructions", "fp64_flops", "fp64_warp_instructions", "mix_pct", "stall_samples_pct"):
                    e[k] = mix[i][k]
            if len(st) == len(summ):
                e["warp_stall_pct"] = st[i]
            dur = num(e.get("gpu__time_duration.sum", "0 us"))
            dram = num(e.get("dram__bytes_read.sum", "0 byte")) + num(e.get("dram__bytes_write.sum", "0 byte"))
            e["dram_bytes"] = dram
            for sub, (bound, amount, what) in table.items():
                if sub in e["kernel"]:
                    e["bound"] = bound
                    e["accounting"] = what
                    if bound == "fp64":
                        flops = amount if amount is not None else e.get("fp64_flops")
                        e["algorithmic_flops"] = flops
                        if flops and dur > 0:
                            e["achieved_tflops"] = flops / dur / 1e12
                            e["roofline_frac"] = flops / dur / 1e12 / FP64_PEAK_TFLOPS
                    elif bound == "int8":
                        e["int8_ops"] = amount
                        if dur > 0:
                            e["achieved_pops"] = amount / dur / 1e15
                            e["roofline_frac"] = amount / dur / 1e15 / INT8_PEAK_POPS
                    else:
                        by = amount if amount is not None else num(e.get("dram__bytes_write.sum", "0 byte"))
                        e["algorithmic_bytes"] = by
                        if by and dur > 0:
                            e["achieved_gbs"] = by / dur / 1e9
                            e["roofline_frac"] = by / dur / 1e9 / HBM_PEAK_GBS
                    break
            result.append(e)
            print(nm, e["kernel"][:70], e.get("gpu__time_duration.sum"), "frac", e.get("roofline_frac"), flush=True)
    if args.merge and os.path.exists(args.merge):   # replace the entries of the re-run workloads, keep the rest
        old = json.load(open(args.merge))["kernels"]
        result = [e for e in old if e.get("workload") not in names] + result
    json.dump({"peaks": {"fp64_tflops": FP64_PEAK_TFLOPS, "hbm_gbs": HBM_PEAK_GBS, "int8_pops_nominal": INT8_PEAK_POPS},
               "kernels": result}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
