#!/bin/bash
# round 2, GPU call 11: full GPU suite + default bench line on the reverted scoring kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest11.log
tail -8 gpurun_out/r2_pytest11.log
( time timeout 900 python bench.py > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err ) 2>&1 | tail -3; tail -c 5000 gpurun_out/r2_bench11.json; tail -5 gpurun_out/r2_bench11.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref11.json 2> gpurun_out/r2_ref11.err ) 2>&1 | tail -3; tail -c 1500 gpurun_out/r2_ref11.json; tail -3 gpurun_out/r2_ref11.err
