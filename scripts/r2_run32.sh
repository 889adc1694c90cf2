#!/bin/bash
# round 2, GPU call 32: ncu --set full of the seeding distance pass (4 M x 576, 20 segments, 4 candidates)
mkdir -p gpurun_out
MAX_ITER=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqdist_cand4 -s 3 -c 1 -f -o gpurun_out/r2_sqdist python scripts/profile_fit.py 4000000 separated > gpurun_out/r2_ncu_sqdist.log 2>&1
ncu -i gpurun_out/r2_sqdist.ncu-rep --page details > gpurun_out/r2_sqdist_details.txt 2>&1
grep -E "Duration|Registers Per|Issue Slots Busy|Executed Ipc Active|No Eligible|Eligible Warps|Active Warps Per|DRAM Throughput|Mem Busy|Max Bandwidth|Pipe|Theoretical Occ|FP64|fp64" gpurun_out/r2_sqdist_details.txt | head -30
ncu -i gpurun_out/r2_sqdist.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
d=dict(zip(h,r))
out=[]
for k in h:
    if 'smsp__average_warps_issue_stalled' in k and '_per_issue_active' in k and 'not_issued' not in k:
        out.append((float(d[k]),k))
for v,k in sorted(out,reverse=True)[:8]: print(round(v,3),k)
for k in h:
    if ('pipe' in k and 'pct' in k) or 'dram__bytes' in k:
        print(k, d[k])
" | head -60
