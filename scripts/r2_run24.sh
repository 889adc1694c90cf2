#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_ood_utils.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/time_vec_score.py > gpurun_out/r2_vec_fast_r2final.json 2> gpurun_out/r2_vec_fast_r2final.err; cat gpurun_out/r2_vec_fast_r2final.json; tail -3 gpurun_out/r2_vec_fast_r2final.err
