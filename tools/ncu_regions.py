#!/usr/bin/env python
"""Samples and stall reasons per code region (bucketed SASS index) of an .ncu-rep."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 250
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Address')
hdr = rows[hi]; sm = hdr.index('# Samples'); ie = hdr.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
buckets = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter(), 0])
tot = 0
for ln, r in enumerate(rows[hi + 1:]):
    if len(r) <= sm: continue
    try: s = float(r[sm] or 0); e = float(r[ie] or 0)
    except ValueError: continue
    b = buckets[ln // B]
    b[0] += s; b[1] += e; b[3] = max(b[3], e); tot += s
    for i in stall_cols:
        try: b[2][hdr[i][6:]] += float(r[i] or 0)
        except ValueError: pass
for k in sorted(buckets):
    s, e, st, mx = buckets[k]
    top = ', '.join(f'{n}={100*v/max(s,1):.0f}%' for n, v in st.most_common(4))
    print(f'[{k*B:5d},{(k+1)*B:5d}) samples {100*s/tot:5.1f}%  exec {e/1e6:8.1f}M  maxcount {mx:9.0f}  {top}')
