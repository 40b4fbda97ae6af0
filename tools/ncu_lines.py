#!/usr/bin/env python
"""Per-CUDA-source-line aggregation of an .ncu-rep (needs -lineinfo):
   python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ''
lines = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No'): continue
    if r[0] and r[0].isdigit() and len(r) >= 8:
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0), float(r[10] or 0) if len(r) > 10 and r[10] not in ('-', '') else 0.0))
        except ValueError:
            pass
ts = sum(l[3] for l in lines) or 1; ti = sum(l[4] for l in lines) or 1
print(f"{len(lines)} source lines, {ti} warp-instr, {ts} samples")
print("-- by samples")
for f, n, s, sm, ex, th in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{f}:{n:4d} samp {sm/ts:6.3f} exec {ex/ti:6.3f} thr {th:5.1f} | {s[:95]}")
print("-- by executed instructions")
for f, n, s, sm, ex, th in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{f}:{n:4d} samp {sm/ts:6.3f} exec {ex/ti:6.3f} thr {th:5.1f} | {s[:95]}")
