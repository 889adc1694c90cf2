#!/bin/bash
# Run bench.py --quick once per variant library (same box, back to back): fused-path time per variant, both map layouts.
for rep in 1 2; do for f in ood_in_object_detection_b200/variants/*.so; do
  OODB200_LIB=$PWD/$f python bench.py --quick --steps 30 --warmup 5 "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
alt = [k for k in d if k.endswith('_fmap_ms') and k != 'fmap_ms']
print('$f'.split('/')[-1], round(d['fmap_ms'] * 1e3, 1), 'us fmap;', round(d['ms_per_step'] * 1e3, 1), 'us step; frac', round(d['frac'], 3), ';', *[(k, round(d[k] * 1e3, 1)) for k in alt])"
done; done
