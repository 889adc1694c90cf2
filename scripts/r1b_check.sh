#!/bin/bash
# Scoring-path check after a kernel change: scoring parity tests, quick bench, launch list (outputs in gpurun_out/).
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_ood_utils.py -x -q -m gpu > $O/pytest_score.log 2>&1; tail -5 $O/pytest_score.log
timeout 300 python bench.py --steps 50 --warmup 5 --fit-n 0 --quick > $O/bench_quick.json 2> $O/bench_quick.err; tail -c 1500 $O/bench_quick.json; tail -3 $O/bench_quick.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_quick.csv python bench.py --steps 3 --warmup 3 --fit-n 0 --quick > $O/ncu_q.log 2>&1
grep -c . $O/launches_quick.csv
