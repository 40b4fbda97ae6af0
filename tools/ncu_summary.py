#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics, stall-reason breakdown and the
hottest SASS lines.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active']


def run(args):
    return subprocess.run(['ncu', '-i', *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    print(f"== {rep}: {len(rows)} launch(es); kernel: {rows[0][hdr.index('Kernel Name')][:90]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {units[i]:14s} {[r[i] for r in rows]}")
    stall = [(h, i) for i, h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
    tot = sum(float(rows[0][i].replace(',', '') or 0) for _, i in stall) or 1
    print("-- stall reasons (launch 0, share of samples)")
    for h, i in sorted(stall, key=lambda t: -float(rows[0][t[1]].replace(',', '') or 0))[:8]:
        print(f"   {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {float(rows[0][i].replace(',', '') or 0) / tot:6.3f}")
    src = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv']))))
    h = next(i for i, r in enumerate(src) if r and r[0] == 'Address')
    cols = src[h]
    si, ii, xi = cols.index('Source'), cols.index('# Samples'), cols.index('Instructions Executed')
    body = []
    for r in src[h + 1:]:
        if len(r) <= max(si, ii, xi) or r[0] == 'Address' or r[0] == 'Kernel Name':
            break
        try:
            body.append((int(r[ii] or 0), int(r[xi] or 0), r[si].strip()))
        except ValueError:
            break
    ts, ti = sum(b[0] for b in body) or 1, sum(b[1] for b in body) or 1
    print(f"-- {len(body)} SASS instructions, {ti} warp-instructions executed, {ts} samples; top {top} by samples")
    for idx, (s, x, t) in sorted(enumerate(body), key=lambda t: -t[1][0])[:top]:
        print(f"   #{idx:4d} samples {s / ts:6.3f} exec {x / ti:6.3f}  {t[:100]}")
    mix = {}
    for s, x, t in body:
        op = t.split()[0] if not t.startswith('@') else t.split()[1]
        op = op.split('.')[0]
        mix[op] = mix.get(op, 0) + x
    print("-- executed instruction mix:", ', '.join(f"{k} {v / ti:.3f}" for k, v in sorted(mix.items(), key=lambda t: -t[1])[:14]))


if __name__ == '__main__':
    main()
