#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
MAX_ITER=30 timeout 300 python scripts/profile_fit.py 4000000 separated > gpurun_out/r2_fit_phases_separated_n1_v3.json 2> gpurun_out/r2_fit_phases_v3.err
tail -c 1500 gpurun_out/r2_fit_phases_separated_n1_v3.json
