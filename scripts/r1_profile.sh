#!/bin/bash
# Round-1 measurement pass on one B200: tests, bench lines, ncu launch lists and full captures (outputs in gpurun_out/).
set -x
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r1_pytest_gpu.log 2>&1; tail -2 $O/r1_pytest_gpu.log
python bench.py --steps 50 --warmup 5 > $O/r1_bench_c2.json 2> $O/r1_bench_c2.err; tail -c 600 $O/r1_bench_c2.json
python bench.py --workload fit --steps 3 --warmup 1 > $O/r1_bench_fit.json 2> $O/r1_bench_fit.err; tail -c 900 $O/r1_bench_fit.json
python bench.py --impl reference --steps 1 --warmup 0 > $O/r1_bench_ref.json 2> $O/r1_bench_ref.err; tail -c 400 $O/r1_bench_ref.json
python scripts/time_ksearch.py > $O/ksearch_time.log 2>&1; cat $O/ksearch_time.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_launches.csv python bench.py --steps 3 --warmup 3 --fit-n 0 > $O/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:'items_kernel|items_nhwc_kernel|score_kernel|plan_geo_kernel' --launch-skip 24 --launch-count 6 -o $O/r1_fmap -f python bench.py --steps 3 --warmup 3 --fit-n 0 --quick > $O/ncu_f1.log 2>&1
KS_N=8000 ncu --set full --clock-control none --import-source on --kernel-name regex:'pair_cluster_sums' --launch-skip 1 --launch-count 1 -o $O/r1_pair -f python scripts/time_ksearch.py > $O/ncu_f3.log 2>&1
ls -la $O/*.ncu-rep
