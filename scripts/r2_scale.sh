#!/bin/bash
# scaling run: the default bench line at N GPUs, launched like the driver does
N=$1; mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ) 2>&1 | tail -3
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    f = d["fit"]
    print("N=$N value", round(d["value"] / 1e6, 2), "M det/s; e2e", round(d["e2e"]["value"] / 1e6, 3), "M/s; device-inputs", round(d["e2e"]["device_inputs"]["value"] / 1e6, 3))
    for k in ("", "separated"):
        g = f[k] if k else f
        print("  fit", g["variant"], "lloyd ms/iter", round(g["lloyd_ms_per_iteration"], 3), "iters", g["lloyd_iterations"], "seed ms", round(g["seed_ms"], 1),
              "fit ms", round(g["fit_ms"], 1), "rows on fullest rank", g["rows_on_fullest_rank"], g["collective"])
    print("  matches", f.get("matches_single_gpu"), f.get("matches_single_gpu_realistic"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_bench_n$N.err").read()[-2000:])
PY
