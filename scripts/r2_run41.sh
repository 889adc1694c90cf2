#!/bin/bash
mkdir -p gpurun_out
MAX_ITER=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:seed_scan -s 5 -c 1 -f -o gpurun_out/r2_scan python scripts/profile_fit.py 4000000 separated > gpurun_out/r2_ncu_scan.log 2>&1
ncu -i gpurun_out/r2_scan.ncu-rep --page details 2>/dev/null | grep -E "Duration|Issue Slots Busy|Executed Ipc|No Eligible|Registers|Executed Instructions|Elapsed Cycles" | head
ncu -i gpurun_out/r2_scan.ncu-rep --page source --csv 2>/dev/null > gpurun_out/r2_scan_source.csv
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2_scan_source.csv")))
h=rows[0]
print(h[:12])
si=[i for i,c in enumerate(h) if c.startswith("# Samples") or c=="Samples" or "Sampling Data (All)" in c]
ei=[i for i,c in enumerate(h) if "Instructions Executed" == c or c=="# Instructions Executed"]
src=h.index("Source") if "Source" in h else 1
print(si, ei)
if si:
    k=si[0]
    def val(r):
        try: return float(r[k])
        except: return 0
    top=sorted(rows[1:], key=val, reverse=True)[:25]
    for r in top: print(r[k], (r[ei[0]] if ei else ''), r[src][:110])
PY
