#!/bin/bash
# round 2, GPU call 15 (2 GPUs): in-kernel barrier for the fused update, matching kernel, reference driver test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matching.py tests/test_gpu_pipeline.py tests/test_gpu_sharded.py -x -q > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest15.log
tail -30 gpurun_out/r2_pytest15.log
for peers in 1 0; do
OODB200_KMEANS_PEERS=$peers MAX_ITER=30 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$peers scripts/profile_fit.py 4000000 realistic > gpurun_out/r2_prof_n2_p$peers.json 2> gpurun_out/r2_prof_n2_p$peers.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_prof_n2_p$peers.json").read().strip().splitlines()[-1])
    print("peers=$peers", d["collective"], "lloyd ms", d["phases_ms"]["lloyd"], "iters", d["lloyd_iters"], {k: v["per_call_ms"] for k, v in d["device_ms"].items() if k in ("step", "update", "update_peers", "reduce_into")})
except Exception as e:
    print("peers=$peers failed", e); print(open("gpurun_out/r2_prof_n2_p$peers.err").read()[-1500:])
PY
done
