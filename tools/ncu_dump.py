#!/usr/bin/env python
"""Dump SASS lines [a, b) of an .ncu-rep with sample counts and top stall.  Usage: ncu_dump.py rep a b"""
import csv, subprocess, sys, io
rep, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Address')
hdr = rows[hi]
si, sm, ie = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
for ln, r in enumerate(rows[hi + 1:]):
    if ln < a or ln >= b or len(r) <= sm: continue
    stalls = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f'#{ln:5d} {r[sm]:>5s} {r[si][:95]:95s} {stalls[0][1]}:{stalls[0][0]:.0f} {stalls[1][1]}:{stalls[1][0]:.0f}')
