#!/bin/bash
# round 2, GPU call 6: per-box kernel (one launch: plan + pool + score per CTA) vs the plan/gather/score sequence
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_ood_utils.py -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
tail -25 gpurun_out/r2_pytest7.log
for v in "X=1" "OODB200_FMAP_NO_PIPE=1"; do
  echo "== $v"; env $v timeout 300 python bench.py --quick --steps 30 --warmup 3 2>&1 | tail -1
done
env timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:pipe_nhwc|plan|items|score_kernel" -c 40 --csv --log-file gpurun_out/r2_l_pipe.csv python bench.py --quick --steps 2 --warmup 3 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2_l_pipe.csv
