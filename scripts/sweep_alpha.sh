#!/bin/bash
for f in ood_in_object_detection_b200/variants/*.so; do for a in 0.5 0.8 100; do
  OODB200_STAGE_ALPHA=$a OODB200_LIB=$PWD/$f python bench.py --quick --steps 30 --warmup 5 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f'.split('/')[-1], 'alpha', $a, round(d['fmap_ms']*1e3,1), 'us fmap;', round(d['ms_per_step']*1e3,1), 'us step; frac', round(d['frac'],3))"
done; done
