"""Phase times of pipe_nhwc_kernel per box (needs a library built with -DOODB200_PIPE_PROF; OODB200_LIB points at it).
Stamps: 0 queue fetch, 1 loads + Q1 plan, 2 descriptor slot free, 3 geometry done, 4 pieces issued, 5 consumer got the
descriptor, 6 pooled, 7 scored."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ood_in_object_detection_b200 import ops, synth
wl = synth.CONFIGS["C2"]
dev = torch.device("cuda", 0)
maps = [m.contiguous(memory_format=torch.channels_last) for m in bench.device_maps(wl, 1000, dev)]
det = synth.detections(2000, wl.batch, wl.img, wl.nc, wl.lam)
clusters, thr, table, lthr = bench.fit_tables(ops, wl, maps, 3000, dev)
batch = ops.make_batch(maps, det["boxes"], det["strides"], det["cls"], wl.img, dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    ops.fmap_score(batch, table, 0b101, True, compat_q1=True)
torch.cuda.synchronize()
ws = ops._workspace(batch, table.nc)
lib_layout_items = None
# the item list sits inside the workspace; find it by scanning for the stamps of the last run: stamps are monotone ns values
raw = ws.cpu().numpy().view(np.uint64)
n = batch.n
st = np.concatenate(det["strides"]).astype(int)
best = None
for off in range(0, len(raw) - n * 8, 32):          # 256-byte aligned sections
    blk = raw[off:off + n * 8].reshape(n, 8)
    if blk[:, [0, 1, 3, 6, 7]].min() > 1e15 and (np.diff(blk[:, [0, 1, 3, 6, 7]].astype(np.int64), axis=1) >= 0).all():
        best = blk.astype(np.int64); break
assert best is not None, "stamps not found (library built without -DOODB200_PIPE_PROF?)"
ok = st <= 2
t0 = best[:, 0].min()
names = ["fetch->planned", "planned->geometry", "geometry->pooled", "pooled->scored", "box total"]
out = {}
for s in range(3):
    b = best[st == s]
    d = [b[:, 1] - b[:, 0], b[:, 3] - b[:, 1], b[:, 6] - b[:, 3], b[:, 7] - b[:, 6], b[:, 7] - b[:, 0]]
    out[f"stride{s}"] = {k: round(float(np.mean(v)) / 1e3, 2) for k, v in zip(names, d)}
    out[f"stride{s}"]["boxes"] = int(len(b))
out["kernel_span_us"] = round(float(best[:, 7].max() - t0) / 1e3, 2)
print(json.dumps(out, indent=1))
