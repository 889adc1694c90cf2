#!/bin/bash
# round 2, GPU call 12 (2 GPUs): sharded fit over NCCL + peer-memory update, equality with the single-GPU fit, scaling at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest12.log
tail -15 gpurun_out/r2_pytest12.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload fit --fit-variant realistic --steps 2 --warmup 1 > gpurun_out/r2_fit_n2.json 2> gpurun_out/r2_fit_n2.err; tail -c 2500 gpurun_out/r2_fit_n2.json; tail -5 gpurun_out/r2_fit_n2.err
OODB200_KMEANS_PEERS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload fit --fit-variant realistic --steps 2 --warmup 1 > gpurun_out/r2_fit_n2_nccl.json 2> gpurun_out/r2_fit_n2_nccl.err; python - <<PY
import json
for f in ("gpurun_out/r2_fit_n2.json", "gpurun_out/r2_fit_n2_nccl.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])["fit"]
        print(f, d["lloyd_ms_per_iteration"], d["fit_ms"], d["seed_ms"], d["collective"], d.get("matches_single_gpu"))
    except Exception as e:
        print(f, "failed", e)
PY
