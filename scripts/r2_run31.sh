#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_ood_utils.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python scripts/time_k_search.py > gpurun_out/r2_k_search_batched.json 2> gpurun_out/r2_k_search.err; cat gpurun_out/r2_k_search_batched.json; tail -3 gpurun_out/r2_k_search.err
