#!/bin/bash
# round 2, GPU call 4: per-kernel durations of the scoring path variants (ncu launch lists, quick bench)
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  echo "== $name"
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:plan|items|score_kernel|dense_pool" -c 100 --csv --log-file gpurun_out/r2_l_$name.csv python bench.py --quick --steps 4 --warmup 3 > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/r2_l_$name.csv
}
run v1 OODB200_FMAP_V1=1
run v2_inline X=1
run v2_group OODB200_FMAP_GROUP_SCORE=1
run v2_group_nodense OODB200_FMAP_GROUP_SCORE=1 OODB200_FMAP_NO_DENSE=1
