#!/bin/bash
# round 2, GPU call 20: the tightened seeding scan (fit tests + phase profile at 4 M)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py -x -q -m gpu 2>&1 | tail -3
MAX_ITER=30 timeout 300 python scripts/profile_fit.py 4000000 realistic > gpurun_out/r2_fit_phases_realistic_n1_scan3.json 2> gpurun_out/r2_fit_phases_scan3.err
tail -c 1400 gpurun_out/r2_fit_phases_realistic_n1_scan3.json
