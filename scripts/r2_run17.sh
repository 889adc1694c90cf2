#!/bin/bash
# round 2, GPU call 17: L2 bulk prefetch of the dense strides' maps from the plan kernel
mkdir -p gpurun_out
for m in 0 6 4 2 7; do
  echo "== OODB200_FMAP_PREFETCH=$m"
  OODB200_FMAP_PREFETCH=$m timeout 300 python bench.py --quick --steps 30 --warmup 5 2>&1 | tail -1
done
echo "== parity with prefetch 6"
OODB200_FMAP_PREFETCH=6 timeout 600 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_ood_utils.py -x -q -m gpu 2>&1 | tail -2
