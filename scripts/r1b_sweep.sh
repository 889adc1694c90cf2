#!/bin/bash
# Scoring parity tests on the default library, then the variant sweep and a launch list of the default (gpurun_out/).
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_ood_utils.py -x -q -m gpu > $O/pytest_score.log 2>&1; tail -5 $O/pytest_score.log
timeout 600 bash scripts/sweep_variants.sh --fit-n 0 > $O/sweep.log 2>&1; cat $O/sweep.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_quick.csv python bench.py --steps 3 --warmup 3 --fit-n 0 --quick > $O/ncu_q.log 2>&1
