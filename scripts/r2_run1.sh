#!/bin/bash
# round 2, GPU call 1: full GPU test suite, vec_score timing (fast vs one-row-per-warp), quick bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
python scripts/time_vec_score.py 2000000 > gpurun_out/r2_vec_fast.json 2> gpurun_out/r2_vec_fast.err; tail -c 1500 gpurun_out/r2_vec_fast.json
OODB200_VEC_FAST=0 python scripts/time_vec_score.py 2000000 > gpurun_out/r2_vec_old.json 2> gpurun_out/r2_vec_old.err; tail -c 1500 gpurun_out/r2_vec_old.json
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 3000 gpurun_out/r2_bench1.json
