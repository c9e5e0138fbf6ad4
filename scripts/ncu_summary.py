"""Summarise an .ncu-rep (raw page) into a small JSON kept under profiles/."""
import csv
import json
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'sm__cycles_elapsed.max.per_second',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio',
        # tensor pipe (tcgen05): share of the launch the pipe is active, int8 sub-pipe cycles, TMEM traffic
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg',
        'smsp__sass_inst_executed_op_tmem_ldt.sum', 'smsp__sass_inst_executed_op_tmem_stt.sum',
        'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed']


def summarise(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {'kernel': vals[hdr.index('Kernel Name')]}
        for i, h in enumerate(hdr):
            if h in KEEP:
                d[h] = f"{vals[i]} {units[i]}".strip()
        res.append(d)
    return res


if __name__ == '__main__':
    res = summarise(sys.argv[1])
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], 'w'), indent=1)
    for d in res:
        print(json.dumps(d, indent=1))


def opcode_mix(rep, kernel_regex=None):
    """Executed-instruction mix per kernel from the source page (SASS): warp instructions by
    opcode, the fp64 flop count (2 per DFMA thread instruction, 1 per DMUL / DADD) and the stall
    samples by opcode."""
    import collections
    import re
    cmd = ['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass']
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    res, cur = [], None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == 'Kernel Name':
            cur = {'kernel': row[1], 'rows': []}
            res.append(cur)
        elif row and row[0] == 'Address':
            cur['hdr'] = row
        elif cur is not None and 'hdr' in cur and len(row) == len(cur['hdr']):
            cur['rows'].append(row)
    summ = []
    for k in res:
        if kernel_regex and not re.search(kernel_regex, k['kernel']):
            continue
        ix = {h: i for i, h in enumerate(k['hdr'])}
        ex, th, sm = collections.Counter(), collections.Counter(), collections.Counter()
        for r in k['rows']:
            parts = r[ix['Source']].split()
            op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
            ex[op] += int(r[ix['Instructions Executed']])
            th[op] += int(r[ix['Predicated-On Thread Instructions Executed']])
            sm[op] += int(r[ix['# Samples']])
        tot = sum(ex.values())
        summ.append({'kernel': k['kernel'], 'warp_instructions': tot,
                     'fp64_flops': 2 * th['DFMA'] + th['DMUL'] + th['DADD'],
                     'fp64_warp_instructions': ex['DFMA'] + ex['DMUL'] + ex['DADD'] + ex['DSETP'],
                     'mix_pct': {o: round(100.0 * n / tot, 2) for o, n in ex.most_common(16)},
                     'stall_samples_pct': {o: round(100.0 * n / max(sum(sm.values()), 1), 2)
                                           for o, n in sm.most_common(10)}})
    return summ
