"""Median per-kernel duration in an ncu launch list (csv from --metrics gpu__time_duration.sum); optional name filter."""
import csv, sys, statistics as st
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
if not rows:
    sys.exit("no launches in " + sys.argv[1])
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size')
flt = sys.argv[2] if len(sys.argv) > 2 else ''
d = defaultdict(list)
for r in rows[1:]:
    name = r[ki].split('(')[0]
    if flt in name:
        d[(name[-60:], r[gi])].append(float(r[vi].replace(',', '')) / 1000)
for (k, g), v in d.items():
    print(f"{k:60s} grid {g:>14s} n={len(v):3d} median {st.median(v):8.2f} us  min {min(v):8.2f}")
