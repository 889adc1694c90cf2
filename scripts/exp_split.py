"""Experiment: the C2 scoring step as P sub-batches of images on P streams inside one CUDA graph."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ood_in_object_detection_b200 import ops, synth
dev = torch.device("cuda:0")
wl = synth.CONFIGS["C2"]
maps = bench.device_maps(wl, 1000, dev)
det = synth.detections(2000, wl.batch, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)
clusters, thr, table, lthr = bench.fit_tables(ops, wl, maps, 3000, dev)
fmask = sum(1 << ops.METRIC_SLOT[m] for m in bench.FMAP_METRICS)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for P in (1, 2, 4):
    B = wl.batch
    cuts = [B * i // P for i in range(P + 1)]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        bt = ops.make_batch([m[a:b] for m in maps], det["boxes"][a:b], det["strides"][a:b], det["cls"][a:b], wl.img, dev)
        parts.append((bt, ops.alloc_fmap_scores(bt.n, dev)))
    streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
    def step():
        cur = torch.cuda.current_stream()
        for (bt, fo), st in zip(parts, streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                ops.fmap_score(bt, table, fmask, True, compat_q1=True, out=fo)
        for st in streams:
            cur.wait_stream(st)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    g.replay(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    print(f"P={P}: {tot / 30 * 1e3:.1f} us per step")
