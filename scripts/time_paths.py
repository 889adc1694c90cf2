"""Time the pooling-only path against the fused paths on the C2 batch (tuning aid; run on the GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ood_in_object_detection_b200 import ops, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
wl = synth.CONFIGS[cfg]
dev = torch.device("cuda", 0)
maps = bench.device_maps(wl, 1000, dev)
det = synth.detections(2000, wl.batch, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)
clusters, thr, table, lthr = bench.fit_tables(ops, wl, maps, 3000, dev)
batch = ops.make_batch(maps, det["boxes"], det["strides"], det["cls"], wl.img, dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
cmax = int(batch.map_chw.reshape(3, 3)[:, 0].max())
pooled = torch.zeros((batch.n, cmax), dtype=torch.float32, device=dev)
fout = ops.alloc_fmap_scores(batch.n, dev)
def timeit(fn, reps=20):
    for _ in range(3):
        flush.zero_(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))
print("boxes", batch.n)
print("roi_pool only      ", timeit(lambda: ops.roi_pool(batch, out=pooled)))
for name, mask in (("L1", 1), ("L2", 2), ("cos", 4), ("L1+cos", 5), ("all3", 7)):
    print(f"fmap_score {name:8s}", timeit(lambda: ops.fmap_score(batch, table, mask, True, compat_q1=True, out=fout)))
for s in range(3):   # one stride at a time (boxes of other strides get an invalid stride -> skipped)
    st = [np.where(x == s, x, 7).astype(np.float32) for x in det["strides"]]
    b2 = ops.make_batch(maps, det["boxes"], st, det["cls"], wl.img, dev)
    nb = int(sum((x == s).sum() for x in det["strides"]))
    print(f"stride {s} only ({nb} boxes): pool", timeit(lambda: ops.roi_pool(b2, out=pooled)), "L1+cos", timeit(lambda: ops.fmap_score(b2, table, 5, True, compat_q1=True, out=fout)))
