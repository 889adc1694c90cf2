#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
(time timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["cpu_baseline"]["kind"], d["clocks"])
print("e2e", d["e2e"]["value"], "device_inputs", d["e2e"]["device_inputs"]["value"], "from_head", d["e2e"]["from_head"].get("value"))
f=d["fit"]; print("fit", f["fit_ms"], f["lloyd_ms_per_iteration"], f["seed_ms"], "sep", f["separated"]["fit_ms"], f["separated"]["seed_ms"])
PY
tail -2 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; tail -c 600 gpurun_out/r2_bench_reference_arm.json
