#!/bin/bash
# round 2, GPU call 23: ncu --set full of the fast vector scorer (D = 576, K = 16, L1)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vec_score_fast -s 2 -c 1 -f -o gpurun_out/r2_vec_fast python scripts/time_vec_score.py 1000000 > gpurun_out/r2_ncu_vec.log 2>&1
ncu -i gpurun_out/r2_vec_fast.ncu-rep --page details > gpurun_out/r2_vec_fast_details.txt 2>&1
grep -E "Duration|Registers Per|Issue Slots Busy|Executed Ipc Active|No Eligible|Eligible Warps|Active Warps Per|Theoretical Occ|Achieved Occ|DRAM Throughput|L1/TEX Hit|Mem Busy|Max Bandwidth|Pipe|Stall|Shared Memory Config|Block Limit" gpurun_out/r2_vec_fast_details.txt | head -50
