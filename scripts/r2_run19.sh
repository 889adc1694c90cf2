#!/bin/bash
# round 2, GPU call 19: per-rank proxy of the N = 8 Lloyd iteration (500 k rows on one GPU), fresh ncu --set full traffic
# capture of the two gather kernels, compute-sanitizer over the small-shape tests
mkdir -p gpurun_out
MAX_ITER=30 timeout 300 python scripts/profile_fit.py 500000 realistic > gpurun_out/r2_fit_phases_500k_n1.json 2> gpurun_out/r2_fit_phases_500k_n1.err
tail -c 1500 gpurun_out/r2_fit_phases_500k_n1.json
for k in items_kernel items_nhwc_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:^$k -s 6 -c 1 -f -o gpurun_out/r2_$k python bench.py --quick --steps 4 --warmup 3 > gpurun_out/r2_ncu_$k.log 2>&1
  ncu -i gpurun_out/r2_$k.ncu-rep --page raw --csv > gpurun_out/r2_${k}_raw.csv 2>/dev/null
  python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2_${k}_raw.csv")))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d.get("Kernel Name"), {m: d.get(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct")}, rows[1][h.index("dram__bytes_read.sum")] if "dram__bytes_read.sum" in h else None)
PY
done
bash scripts/sanitize.sh
