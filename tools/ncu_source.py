#!/usr/bin/env python
"""Top stalled SASS lines of an .ncu-rep (source page).  Usage: ncu_source.py rep [top] [--window k]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Address')
hdr = rows[hi]
si, sm = hdr.index('Source'), hdr.index('# Samples')
stall_cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
ie = hdr.index('Instructions Executed')
data = []
for ln, r in enumerate(rows[hi + 1:]):
    if len(r) <= sm: continue
    try: s = float(r[sm] or 0)
    except ValueError: continue
    stalls = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    data.append((s, ln, r[si], stalls, r[ie]))
tot = sum(d[0] for d in data) or 1
print('total samples', tot, 'instructions', len(data))
agg = {}
for i in stall_cols:
    agg[hdr[i][6:]] = sum(float(r[i] or 0) for r in rows[hi + 1:] if len(r) > i and r[i].replace('.', '').isdigit())
print('stall totals:', ', '.join(f'{k}={100*v/tot:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, ln, src, stalls, n in sorted(data, reverse=True)[:top]:
    print(f'{100*s/tot:6.2f}%  #{ln:5d} x{n:>9s} {src[:90]:90s} {stalls[0][1]}:{stalls[0][0]:.0f} {stalls[1][1]}:{stalls[1][0]:.0f}')
