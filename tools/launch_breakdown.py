#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); mu = h.index('Metric Unit')
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(r[mu], 1e-6)
    name = r[kn].split('(')[0].replace('void ', '').replace('dgx::<unnamed>::', '').replace('<unnamed>::', '')[-70:]
    tot[name] += v * scale; cnt[name] += 1
T = sum(tot.values())
print(f"total {T:.2f} ms over {sum(cnt.values())} launches")
for k, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{v:9.2f} ms {100*v/T:5.1f}%  x{cnt[k]:5d}  {v/cnt[k]*1e3:9.1f} us/launch  {k}")
