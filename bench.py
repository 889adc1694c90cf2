#!/usr/bin/env python
"""bench.py -- throughput of the OoD-scoring hot path on B200 (contract: see the task prompt / DESIGN.md §5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2]

One "step" = one fused scoring pass over one synthetic batch of the BASELINE.json configuration
(default configs[1] = C2: YOLOv8s 640x640, batch 64, FMap L1 + cosine against K=10 centroids per
(class, stride), Energy / MSP / max-logit scored in the same pass).  N > 1: one process per GPU
(torchrun), image batches sharded, no collective on the scoring path (weak scaling: 64 images per GPU).

Printed JSON (rank 0, one line): `value` = detections scored per second with inputs resident in HBM
(CUDA events, max over ranks); `e2e` = the same through the public class API with HOST inputs
(H2D + kernels + D2H of the decisions inside the timed region); `roofline` for the dominant kernel
(fmap_score) from algorithmic bytes / its own CUDA-event time; `cpu_baseline` = the
reference's per-box CPU path (oracle/cpu_path.py, a loop-for-loop port using the same torchvision /
sklearn calls) timed on a bounded sample on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "detections OoD-scored/sec"
UNIT = "detections/s"
FMAP_METRICS = ("l1", "cosine")
LOGIT_METHODS = ("MSP", "Energy", "MaxLogit")
CPU_SAMPLE_IMAGES = 6


# ----------------------------------------------------------------------------------------- helpers
def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def algorithmic_bytes(det, wl, k, n_tables=1):
    """sum over images/strides of 4*C_s*|union of window cells| + 37 B/box + centroid slices once."""
    total = 0
    upper = 0
    used = set()
    for i in range(len(det["boxes"])):
        for s in range(3):
            H = W = wl.map_hw[s]
            sc = np.float32(W / wl.img)
            sel = det["strides"][i] == s
            if not sel.any():
                continue
            mask = np.zeros((H, W), bool)
            for b in det["boxes"][i][sel]:
                x1, y1, x2, y2 = (float(np.float32(v) * sc) for v in b)
                r0, r1 = int(np.floor(y1)), min(int(np.ceil(y2)), H - 1)
                c0, c1 = int(np.floor(x1)), min(int(np.ceil(x2)), W - 1)
                mask[max(r0, 0):r1 + 1, max(c0, 0):c1 + 1] = True
                upper += 4 * wl.channels[s] * (np.ceil(max(y2 - y1, 1)) + 1) * (np.ceil(max(x2 - x1, 1)) + 1)
            total += 4 * wl.channels[s] * int(mask.sum())
            for c in np.unique(det["cls"][i][sel]):
                used.add((int(c), s))
    n = sum(len(b) for b in det["boxes"])
    total += 37 * n
    total += n_tables * 4 * sum(k * wl.channels[s] for _, s in used)
    return int(total), int(upper), n


# ------------------------------------------------------------------------------------ workload setup
def device_maps(wl, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = []
    for c, hw in zip(wl.channels, wl.map_hw):
        u = torch.randint(0, 65536, (wl.batch, c, hw, hw), device=device, generator=g).to(torch.float64) / 65536.0
        out.append((7.0 * u ** 4 - 0.27).to(torch.float32).contiguous())
    return out


def fit_tables(ops, wl, maps, seed, device):
    """Fit-stage stand-in (untimed setup): centroids = means of K random groups of the pooled train vectors per
    (class, stride); thresholds = TPR-95 'lower' percentile of the train distances / 5th percentile of train logit
    scores.  Pooling and scoring of the train boxes run through the same CUDA kernels."""
    import torch
    from ood_in_object_detection_b200 import synth
    rng = np.random.default_rng(seed)
    tr = synth.detections(seed, wl.batch, wl.img, wl.nc, 300, fixed=True)
    tb = ops.make_batch(maps, tr["boxes"], tr["strides"], tr["cls"], wl.img, device)
    pooled = ops.roi_pool(tb)
    pooled = pooled / pooled.norm(dim=1, keepdim=True).clamp_min(1e-12)
    pooled = pooled.cpu().numpy()
    cls = np.concatenate(tr["cls"]).astype(int)
    st = np.concatenate(tr["strides"]).astype(int)
    clusters = [[np.empty(0)] * 3 for _ in range(wl.nc)]
    for c in range(wl.nc):
        for s in range(3):
            v = pooled[(cls == c) & (st == s)][:, :wl.channels[s]]
            if len(v) > wl.k:
                grp = rng.integers(0, wl.k, size=len(v))
                clusters[c][s] = np.stack([v[grp == j].mean(0) if (grp == j).any() else v[j] for j in range(wl.k)])
    big = [[1e30] * 3 for _ in range(wl.nc)]
    table = ops.pack_centroids(clusters, {0: big, 1: big, 2: big}, list(wl.channels), device)
    res = ops.fmap_score(tb, table, 0b111, compat_q1=False)     # fit path uses the box's own class (ood_utils.py:1769-1776)
    dist = res.dist.cpu().numpy()
    thr = {}
    for m in range(3):
        t = [[[] for _ in range(3)] for _ in range(wl.nc)]
        for c in range(wl.nc):
            for s in range(3):
                d = dist[m][(cls == c) & (st == s)]
                if len(d) > 5 and len(clusters[c][s]):
                    t[c][s] = float(np.percentile(d, 95, method="lower"))
        thr[m] = t
    table = ops.pack_centroids(clusters, thr, list(wl.channels), device)
    z = torch.from_numpy(np.concatenate(tr["logits"])).to(device)
    lg = ops.logit_score(z, tb.cls, 0b11111).scores.cpu().numpy()
    lthr = np.zeros((5, wl.nc))
    for m in range(5):
        for c in range(wl.nc):
            v = lg[m][cls == c]
            if len(v) > 5:
                lthr[m, c] = float(np.percentile(v, 5.000000000000004, method="lower"))
    return clusters, thr, table, lthr


# ------------------------------------------------------------------------------------- CPU reference
def cpu_reference_pass(det, maps_cpu, wl, clusters, thr, lthr, images):
    """One pass of the reference's CPU path (port) over `images` (indices); returns (#boxes, seconds)."""
    import torch
    from oracle import cpu_path
    ims = [dict(maps=[m[i] for m in maps_cpu], boxes=torch.from_numpy(det["boxes"][i]),
                cls=torch.from_numpy(det["cls"][i]), strides=torch.from_numpy(det["strides"][i]),
                logits=torch.from_numpy(det["logits"][i]), img_hw=(wl.img, wl.img)) for i in images]
    t0 = time.perf_counter()
    for metric in FMAP_METRICS:
        cpu_path.distance_decisions(ims, clusters, thr[{"l1": 0, "l2": 1, "cosine": 2}[metric]], metric)
    for name in LOGIT_METHODS:
        if name == "MaxLogit":
            continue                                   # no reference implementation exists
        cpu_path.logit_decisions(ims, name, lthr[{"MSP": 0, "Energy": 1}[name]].tolist(), 1.0)
    dt = time.perf_counter() - t0
    return sum(len(det["boxes"][i]) for i in images), dt


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (loop-for-loop port, same third-party
    calls) on this box's host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from ood_in_object_detection_b200 import synth
    n_img = CPU_SAMPLE_IMAGES
    det = synth.detections(2000, n_img, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)
    maps = [torch.from_numpy(m) for m in synth.feature_maps(1000, n_img, wl.channels, wl.map_hw)]
    rng = np.random.default_rng(0)
    clusters = [[np.abs(rng.standard_normal((wl.k, c))).astype(np.float32) / np.sqrt(c) for c in wl.channels]
                for _ in range(wl.nc)]
    thr = {m: [[1.0] * 3 for _ in range(wl.nc)] for m in range(3)}
    lthr = np.zeros((5, wl.nc))
    imgs = list(range(n_img))
    for _ in range(args.warmup):
        cpu_reference_pass(det, maps, wl, clusters, thr, lthr, imgs[:1])
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        n, dt = cpu_reference_pass(det, maps, wl, clusters, thr, lthr, imgs)
        tot_n += n
        tot_t += dt
    v = tot_n / tot_t
    sample = (f"{n_img} images (~{tot_n // max(args.steps, 1)} boxes) of the {wl.batch}-image batch per step; L1+cosine FMap "
              f"and MSP+Energy logit decisions, per-box sklearn/torchvision calls as in ood_utils.py:2038-2180")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "sample_images": n_img},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from ood_in_object_detection_b200 import ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    maps = device_maps(wl, 1000 + rank, device)
    det = synth.detections(2000 + rank, wl.batch, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)
    clusters, thr, table, lthr = fit_tables(ops, wl, maps, 3000 + rank, device)
    fmask = sum(1 << ops.METRIC_SLOT[m] for m in FMAP_METRICS)
    lmask = sum(1 << ops.LOGIT_SLOT[m] for m in LOGIT_METHODS)
    lthr_d = torch.from_numpy(lthr).to(device)

    batch = ops.make_batch(maps, det["boxes"], det["strides"], det["cls"], wl.img, device)
    logits = torch.from_numpy(np.concatenate(det["logits"])).to(device)
    n = batch.n
    fout = ops.alloc_fmap_scores(n, device)
    lout = ops.LogitScores(scores=torch.zeros((5, n), dtype=torch.float32, device=device), indness=None,
                           decision=torch.ones((5, n), dtype=torch.uint8, device=device),
                           sigmoid_mismatch=torch.zeros(1, dtype=torch.int32, device=device))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)     # > 126 MB L2

    def step(ev=None):
        if ev:
            ev[0].record()
        ops.fmap_score(batch, table, fmask, True, compat_q1=True, out=fout)     # memset + plan_kernel + items_kernel
        if ev:
            ev[1].record()
        ops.logit_score(logits, batch.cls, lmask, thr=lthr_d, out=lout)
    launches_per_step = 3                                                       # plan, items, logit

    for _ in range(args.warmup):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    E = lambda: torch.cuda.Event(enable_timing=True)
    evs = [(E(), E(), E(), E()) for _ in range(args.steps)]
    clk = ClockSampler(local)
    clk.__enter__()                                                      # sampled over the timed steps and the e2e loop
    if True:
        torch.cuda.synchronize()
        t_wall = time.perf_counter()
        for a, b, c, d in evs:
            flush.zero_()                                                # L2 flush between timed iterations
            a.record()
            step((b, c))
            d.record()
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall
    total_ms = sum(a.elapsed_time(d) for a, b, c, d in evs)
    fmap_ms = sum(b.elapsed_time(c) for a, b, c, d in evs) / max(args.steps, 1)
    if world > 1:
        dist.barrier()
        t = torch.tensor([total_ms, float(n)], dtype=torch.float64, device=device)
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms, n_all = float(tm[0]), float(t[1])
    else:
        n_all = float(n)
    value = n_all * args.steps / (total_ms * 1e-3)
    if args.quick:
        clk.__exit__()
        if rank == 0:
            alg, upper, _ = algorithmic_bytes(det, wl, wl.k)
            print(json.dumps({"value": value, "ms_per_step": total_ms / args.steps, "fmap_ms": fmap_ms,
                              "frac": alg / (fmap_ms * 1e-3) / 1e9 / _peaks()[0], "quick": True}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end: host (pinned) inputs -> decisions on the host, every step
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_maps = [m.cpu().pin_memory() for m in maps]
    h_boxes, h_str, h_cls = [pin(b) for b in det["boxes"]], [pin(s) for s in det["strides"]], [pin(c) for c in det["cls"]]
    h_logits = pin(np.concatenate(det["logits"]))
    h2d = sum(m.numel() * 4 for m in h_maps) + sum(b.numel() * 4 for b in h_boxes + h_str + h_cls) + h_logits.numel() * 4
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        b = ops.make_batch(h_maps, h_boxes, h_str, h_cls, wl.img, device)
        z = h_logits.to(device, non_blocking=True)
        fr = ops.fmap_score(b, table, fmask, True, compat_q1=True)
        lr = ops.logit_score(z, b.cls, lmask, thr=lthr_d)
        return fr.decision.cpu(), lr.decision.cpu()
    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fd, ld = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    d2h = fd.numel() + ld.numel()
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = n_all * e2e_steps / e2e_s
    clk.__exit__()

    if rank == 0:
        alg, upper, _ = algorithmic_bytes(det, wl, wl.k)
        peak, how = _peaks()
        achieved = alg / (fmap_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(args.config, {}).get("fmap_dram_bytes_per_launch")
        maps_cpu = [m[:CPU_SAMPLE_IMAGES].cpu() for m in maps]
        cpu_reference_pass(det, maps_cpu, wl, clusters, thr, lthr, [0])        # warm the imports / thread pools
        cpu_n, cpu_t = cpu_reference_pass(det, maps_cpu, wl, clusters, thr, lthr, list(range(CPU_SAMPLE_IMAGES)))
        clocks = clk.summary()
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "images_per_gpu": wl.batch, "boxes_per_gpu": n, "fmap_metrics": FMAP_METRICS,
                       "logit_methods": LOGIT_METHODS, "k_per_class_stride": wl.k, "nc": wl.nc,
                       "l2_flush": "512 MiB memset between timed iterations", "sharding": f"batch x{world}, no collective"},
            "roofline": {"bound": "hbm", "kernel": "plan_kernel+items_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": alg,
                         "upper_bound_bytes": upper, "kernel_ms": fmap_ms, "peak_source": how},
            "cpu_baseline": {"value": cpu_n / cpu_t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{CPU_SAMPLE_IMAGES} of {wl.batch} images ({cpu_n} boxes), L1+cosine FMap and "
                                       f"MSP+Energy logits through oracle/cpu_path.py (per-box sklearn/torchvision calls)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "note": "host pinned feature maps + detections -> decisions on host"},
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C4", "C5"])
    ap.add_argument("--quick", action="store_true", help="kernel timing only: skip the e2e and cpu_baseline legs (tuning sweeps)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    from ood_in_object_detection_b200 import synth
    wl = synth.CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
