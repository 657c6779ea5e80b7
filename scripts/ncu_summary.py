#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): key raw metrics per launch + per-region SIMT efficiency from the
source page. usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--regions N] > profiles/x_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg']


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    chunk = int(sys.argv[sys.argv.index('--regions') + 1]) if '--regions' in sys.argv else 0
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        print(f'=== launch {n}: {r[hdr.index("Kernel Name")][:90]}')
        for k in KEYS:
            if k in hdr:
                print(f'  {k:72s} {r[hdr.index(k)]} {units[hdr.index(k)]}')
        st = []
        for i, h in enumerate(hdr):
            if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and r[i]:
                st.append((float(r[i].replace(',', '')), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
        print('  stalls per issue:', ', '.join(f'{h} {v:.2f}' for v, h in sorted(st, reverse=True)[:7]))
    if chunk:
        rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', '--kernel-id', ':::1']))))
        hdr = rows[1]
        data = [r for r in rows[2:] if len(r) > 10 and r[0].startswith('0x')]
        iS, iI, iT, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
        ti = sum(int(r[iI]) for r in data)
        tt = sum(int(r[iT]) for r in data)
        ts = sum(int(r[iSm]) for r in data)
        print(f'=== source page, launch 1: {len(data)} SASS instructions, {ti} warp instructions executed, {tt / ti:.2f} threads active on average')
        print('  region(first SASS index)  share of executed warp instructions  share of stall samples  avg active threads  dominant opcodes')
        for k in range(0, len(data), chunk):
            seg = data[k:k + chunk]
            i = sum(int(r[iI]) for r in seg)
            t = sum(int(r[iT]) for r in seg)
            s = sum(int(r[iSm]) for r in seg)
            if i * 1000 < ti:
                continue
            ops = {}
            for r in seg:
                w = r[iS].split()
                op = (w[1] if w[0].startswith('@') else w[0]).split('.')[0]
                ops[op] = ops.get(op, 0) + 1
            top = ' '.join(f'{o}x{c}' for o, c in sorted(ops.items(), key=lambda x: -x[1])[:4])
            print(f'  {k:5d}  inst {100 * i / ti:5.1f}%  samples {100 * s / ts:5.1f}%  active {t / max(i, 1):5.1f}  {top}')


if __name__ == '__main__':
    main()
