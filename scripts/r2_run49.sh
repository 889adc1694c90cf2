#!/bin/bash
# 2-rank proxy of the N = 8 per-rank share (500 k rows per rank): where does the seeding time go
mkdir -p gpurun_out
MAX_ITER=5 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/profile_fit.py 1000000 realistic > gpurun_out/r2_fit_phases_1m_n2.json 2> gpurun_out/r2_fit_phases_1m_n2.err
tail -c 1500 gpurun_out/r2_fit_phases_1m_n2.json
