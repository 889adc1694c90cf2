#!/bin/bash
# round 2, GPU call 3: v2 scoring path (box-parallel plan, dense plane jobs via TMA, scoring at box completion)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_ood_utils.py -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -25 gpurun_out/r2_pytest3.log
for v in "" "OODB200_FMAP_V1=1" "OODB200_FMAP_NO_DENSE=1"; do
  echo "== $v"; env $v timeout 300 python bench.py --quick --steps 30 --warmup 3 2>&1 | tail -2
done
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_fit.py -x -q > gpurun_out/r2_pytest3b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3b.log
tail -8 gpurun_out/r2_pytest3b.log
