"""Median per-kernel duration of the oodb200 kernels in an ncu launch list (csv from --metrics gpu__time_duration.sum)."""
import csv, sys, statistics as st
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
d = defaultdict(list)
for r in rows[1:]:
    if 'oodb200' in r[ki]:
        d[r[ki].split('(')[0]].append(float(r[vi].replace(',', '')) / 1000)
for k, v in d.items():
    print(f"{k:48s} n={len(v):3d} median {st.median(v):8.2f} us  min {min(v):8.2f}")
