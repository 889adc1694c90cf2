"""Hot spots of one kernel from `ncu -i rep --page source --csv`: share of executed instructions / stall samples per
region of the SASS listing and the most-sampled instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if 'Source' in r and '# Samples' in r)
ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
data = [r for r in rows if len(r) > iex and r[iex].isdigit()]
tot = sum(int(r[iex]) for r in data) or 1; ts = sum(int(r[isamp]) for r in data) or 1
print('sass lines', len(data), 'warp instructions', tot, 'samples', ts)
n = len(data); B = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for b in range(B):
    seg = data[b * n // B:(b + 1) * n // B]
    e = sum(int(r[iex]) for r in seg); s = sum(int(r[isamp]) for r in seg)
    if e * 100 > tot or s * 100 > ts:
        print('%3d @%5d exec %5.1f%% samples %5.1f%%  %s' % (b, b * n // B, 100 * e / tot, 100 * s / ts, seg[len(seg) // 2][ia][:60]))
for r in sorted(data, key=lambda r: -int(r[isamp]))[:24]:
    print(data.index(r), r[isamp], r[iex], r[ia][:100])
