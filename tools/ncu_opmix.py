#!/usr/bin/env python
"""Executed-instruction mix per opcode (weighted by 'Instructions Executed') from an .ncu-rep source page."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Address')
hdr = rows[hi]; si = hdr.index('Source'); ie = hdr.index('Instructions Executed')
mix = collections.Counter(); static = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= ie: continue
    src = r[si].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    try: mix[op] += float(r[ie] or 0); static[op] += 1
    except ValueError: pass
tot = sum(mix.values())
print('total warp-instructions', tot, 'static', sum(static.values()))
for op, c in mix.most_common(25):
    print(f'{op:12s} {100*c/tot:6.2f}%  static {static[op]}')
