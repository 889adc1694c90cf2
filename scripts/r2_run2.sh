#!/bin/bash
# round 2, GPU call 2: GPU tests again, fit phase profile (CUDA events per backend step), fit bench lines, ncu launch list of a fit
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
python scripts/profile_fit.py 4000000 separated > gpurun_out/r2_prof_sep.json 2> gpurun_out/r2_prof_sep.err; tail -c 2500 gpurun_out/r2_prof_sep.json; tail -3 gpurun_out/r2_prof_sep.err
MAX_ITER=30 python scripts/profile_fit.py 4000000 realistic > gpurun_out/r2_prof_real.json 2> gpurun_out/r2_prof_real.err; tail -c 2500 gpurun_out/r2_prof_real.json; tail -3 gpurun_out/r2_prof_real.err
python bench.py --workload fit --fit-variant realistic --steps 2 --warmup 1 > gpurun_out/r2_fit_real.json 2> gpurun_out/r2_fit_real.err; tail -c 3500 gpurun_out/r2_fit_real.json; tail -3 gpurun_out/r2_fit_real.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_fit_launches.csv python scripts/profile_fit.py 4000000 separated > gpurun_out/r2_ncu_fit.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_fit_launches.csv 2>/dev/null | head -30
