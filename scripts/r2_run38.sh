#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "seed" 2>&1 | tail -15
OODB200_SEED_SCAN=serial timeout 600 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "seed_scan" 2>&1 | tail -3
MAX_ITER=3 timeout 300 python scripts/profile_fit.py 4000000 separated > gpurun_out/r2_fit_phases_parscan.json 2> gpurun_out/r2_fit_phases_parscan.err
tail -c 1300 gpurun_out/r2_fit_phases_parscan.json
