#!/usr/bin/env python
import csv, subprocess, sys, io, collections
rep, lo, hi_ = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Address')
hdr = rows[hi]; si = hdr.index('Source'); ie = hdr.index('Instructions Executed')
mix = collections.Counter(); static = collections.Counter()
for ln, r in enumerate(rows[hi + 1:]):
    if ln < lo or ln >= hi_ or len(r) <= ie: continue
    toks = r[si].split()
    if not toks: continue
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    mix[op] += float(r[ie] or 0); static[op] += 1
tot = sum(mix.values())
print('range', lo, hi_, 'executed', tot, 'static', sum(static.values()))
for op, c in mix.most_common(22): print(f'{op:10s} {100*c/tot:6.2f}%  static {static[op]:4d}  per-warp-plane {c/327680:7.1f}')
