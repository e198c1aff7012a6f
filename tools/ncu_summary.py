#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per launch + top stall lines.  Usage: ncu_summary.py rep [--source N]"""
import csv, subprocess, sys, io, collections

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'local_load', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']

def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('== launch', r[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:90s} {r[i]:>16s} {units[i]}')

def source(rep, top):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # find header
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and any('Sampl' in c for c in r))
    hdr = rows[hi]
    si = hdr.index('Source')
    samp = next(i for i, c in enumerate(hdr) if c.startswith('# Samples') or c == 'Warp Stall Sampling (All Samples)' or 'Sampling (All' in c)
    data = []
    for ln, r in enumerate(rows[hi + 1:]):
        try:
            data.append((float(r[samp] or 0), ln, r[si]))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    print('total samples', tot, 'instructions', len(data))
    for s, ln, src in sorted(data, reverse=True)[:top]:
        print(f'{100*s/tot:6.2f}%  #{ln:5d}  {src[:110]}')

if __name__ == '__main__':
    rep = sys.argv[1]
    raw(rep)
    if '--source' in sys.argv:
        source(rep, int(sys.argv[sys.argv.index('--source') + 1]))
