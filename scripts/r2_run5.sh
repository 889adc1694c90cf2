#!/bin/bash
# round 2, GPU call 5: what bounds dense_pool_kernel (copy issue, pooling from shared memory, job set-up)
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:plan2|items_kernel|dense_pool" -c 40 --csv --log-file gpurun_out/r2_d_$name.csv python bench.py --quick --steps 2 --warmup 3 > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/r2_d_$name.csv dense
}
run lanes OODB200_FMAP_GROUP_SCORE=1
run onethread OODB200_FMAP_GROUP_SCORE=1 OODB200_DENSE_DBG=4
run onecopy OODB200_FMAP_GROUP_SCORE=1 OODB200_DENSE_DBG=1
run lanes_nopool OODB200_FMAP_GROUP_SCORE=1 OODB200_DENSE_DBG=2
run onecopy_nopool OODB200_FMAP_GROUP_SCORE=1 OODB200_DENSE_DBG=3
