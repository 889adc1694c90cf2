#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_ood_utils.py -x -q > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
tail -15 gpurun_out/r2_pytest9.log
timeout 300 python bench.py --quick --steps 30 --warmup 3 2>&1 | tail -1
OODB200_LIB=ood_in_object_detection_b200/variants/prof.so python scripts/pipe_prof.py > gpurun_out/r2_pipe_prof2.json 2> gpurun_out/r2_pipe_prof2.err; cat gpurun_out/r2_pipe_prof2.json | tr -d '\n '; echo; tail -3 gpurun_out/r2_pipe_prof2.err
