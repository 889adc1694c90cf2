#!/bin/bash
# round 2, GPU call 22: vec_score fast kernel with bulk-copy row staging
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/time_vec_score.py > gpurun_out/r2_vec_fast_tma.json 2> gpurun_out/r2_vec_fast_tma.err; cat gpurun_out/r2_vec_fast_tma.json; tail -3 gpurun_out/r2_vec_fast_tma.err
