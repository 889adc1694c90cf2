"""Where the time of a C3-shaped fit goes: CUDA-event time of every backend step (device time between the call's first
and last launch) next to the wall time of the phases.  Usage: python scripts/profile_fit.py [n_vectors] [variant]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ood_in_object_detection_b200 import kmeans

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
variant = sys.argv[2] if len(sys.argv) > 2 else "separated"
import torch.distributed as dist
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
group = dist.group.WORLD if world > 1 else None
x, gsizes, lsizes = bench.fit_data(n, world, rank, dev, variant=variant)


class Timed(kmeans.CudaBackend):
    def __init__(self, device):
        super().__init__(device)
        self.log = []

for name in ("step", "reduce", "reduce_into", "reduce_step", "update", "update_peers", "converge", "seed_scan", "seed_sqdist", "seed_gather", "seed_pick", "colsum", "center"):
    def wrap(name):
        base = getattr(kmeans.CudaBackend, name)
        def f(self, *a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = base(self, *a, **k); e1.record()
            self.log.append((name, e0, e1))
            return out
        return f
    setattr(Timed, name, wrap(name))

out = {}
for rep in range(2):
    be = Timed(dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if world > 1:
        r = kmeans.kmeans_fit_sharded(x, lsizes, gsizes, bench.FIT_K, world, rank, group, random_state=10, backend=be,
                                      max_iter=int(os.environ.get("MAX_ITER", "300")))
    else:
        r = kmeans.kmeans_fit_predict_single(x, gsizes, bench.FIT_K, random_state=10, backend=be, max_iter=int(os.environ.get("MAX_ITER", "300")))
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    agg = {}
    for name, e0, e1 in be.log:
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
    out = {"n": n, "variant": variant, "wall_ms": 1e3 * wall, "phases_ms": {k: round(1e3 * v, 3) for k, v in r.seconds.items() if isinstance(v, float)},
           "lloyd_iters": r.seconds["lloyd_iters"], "issued": r.seconds.get("lloyd_issued"),
           "device_ms": {k: {"calls": v[0], "total_ms": round(v[1], 3), "per_call_ms": round(v[1] / v[0], 4)} for k, v in agg.items()}}
out["collective"] = r.seconds.get("collective")
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
