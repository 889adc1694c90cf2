#!/bin/bash
# ncu --set full capture of the small kernels around the gather (outputs in gpurun_out/).
set -x
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:'plan_geo_kernel|score_kernel' --launch-skip 8 --launch-count 2 -o $O/r1b_small -f python bench.py --steps 3 --warmup 3 --fit-n 0 --quick > $O/ncu_s.log 2>&1
ls -la $O/*.ncu-rep
