#!/bin/bash
mkdir -p gpurun_out
for b in 512 1024 2048; do
  OODB200_BLOCK_ROWS=$b MAX_ITER=10 timeout 300 python scripts/profile_fit.py 4000000 realistic 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$b', {k:d['device_ms'][k]['per_call_ms'] for k in ('step','reduce_step','update')}, d['phases_ms'])"
done
