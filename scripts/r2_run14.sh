#!/bin/bash
mkdir -p gpurun_out
for peers in 1 0; do
OODB200_KMEANS_PEERS=$peers MAX_ITER=30 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$peers scripts/profile_fit.py 4000000 realistic > gpurun_out/r2_prof_n2_p$peers.json 2> gpurun_out/r2_prof_n2_p$peers.err; tail -c 1800 gpurun_out/r2_prof_n2_p$peers.json; echo; tail -2 gpurun_out/r2_prof_n2_p$peers.err
done
