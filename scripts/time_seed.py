"""Per-kernel timing of the device seeding at the C3 shape (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ood_in_object_detection_b200 import kmeans

n_seg, per, dim, T = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 200000, 576, 4
dev = torch.device("cuda:0")
be = kmeans.CudaBackend(dev)
n = n_seg * per
x = torch.randn(n, dim, device=dev) * 0.05
off = torch.arange(n_seg + 1, device=dev, dtype=torch.int64) * per
cand = torch.randn(n_seg, T, dim, device=dev) * 0.05
closest = torch.rand(n, device=dev)
pot = torch.full((n_seg,), float(per) * 0.5, device=dev)
uni = torch.rand(n_seg, T, device=dev, dtype=torch.float64)
trials = torch.full((n_seg,), T, dtype=torch.int32, device=dev)
on = torch.ones(n_seg, dtype=torch.int32, device=dev)
piece_off = off[:-1].reshape(n_seg, 1).contiguous()
piece_cnt = torch.full((n_seg, 1), per, dtype=torch.int64, device=dev)
cand_id = torch.zeros((n_seg, T), dtype=torch.int64, device=dev)
cent = torch.zeros(n_seg, 16, dim, device=dev)
shard_first = torch.zeros(n_seg, dtype=torch.int64, device=dev)

def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / reps:.3f} ms")

timeit("seed_sqdist(4 cand)", lambda: be.seed_sqdist(x, off, per, cand, closest))
timeit("sqdist_cand old   ", lambda: be.sqdist_cand(x, off, per, cand, closest))
timeit("seed_scan", lambda: be.seed_scan(closest, piece_off, piece_cnt, uni, pot, trials, on, per, cand_id))
timeit("seed_gather", lambda: be.seed_gather(x, cand_id, off, shard_first))
newd, pots = be.seed_sqdist(x, off, per, cand, closest)
timeit("seed_pick", lambda: be.seed_pick(pots, trials, on, off, per, newd, cand, closest, pot, cent, 3))
print("x bytes", x.numel() * 4 / 1e9, "GB")
