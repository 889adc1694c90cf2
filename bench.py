#!/usr/bin/env python
"""bench.py -- throughput of the OoD-scoring hot path on B200 (contract: task prompt / DESIGN.md section 5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2] [--workload score|fit]

workload score (default, BASELINE.json configs[1] = C2: YOLOv8s 640x640, batch 64, FMap L1 + cosine against K=10
centroids per (class, stride), MSP / Energy / max-logit in the same step):
  step    = one scoring pass over one synthetic batch = plan + geo + gather + score + logit kernels, replayed as a CUDA
            graph.  N > 1: one process per GPU (torchrun), image batches sharded, NO collective on the scoring path
            (weak scaling: 64 images per GPU).
  value   = detections scored per second with inputs resident in HBM (CUDA events, max over ranks, L2 flushed between
            timed iterations).
  e2e     = the same through the public class API (ood_utils.compute_ood_decisions_fused: the reference's method
            classes) with pinned HOST inputs: feature maps + detections H2D, decisions D2H inside the timed region.
  roofline for the dominant kernel (items_kernel, the window gather): algorithmic bytes (union of
            window cells per image/stride + 37 B/box + used centroid slices) / CUDA-event time of the fused-path
            launches; `traffic` = dram bytes of the same launches from the committed ncu capture (profiles/).
  fit     = sub-measurement: segmented k-means (K=16, D=576, 20 classes) on --fit-n vectors sharded over the ranks, one
            all-reduce per Lloyd iteration + exact percentile thresholds (strong scaling), vectors x iterations / s.
  cpu_baseline = the reference's per-box CPU path (oracle/cpu_path.py: loop-for-loop port, same torchvision / sklearn
            calls) on a bounded sample of the same batch, on this box's host cores.
workload fit: the C3 fit (4 M vectors) as the main line.
--impl reference: the CPU port alone, same metric / config, bounded sample per step (rank 0 only).
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
FIT_METRIC = "k-means fit vectors/sec"
FIT_UNIT = "vector-iterations/s"
FMAP_METRICS = ("l1", "cosine")
LOGIT_METHODS = ("MSP", "Energy", "MaxLogit")
CPU_SAMPLE_IMAGES = 6
FIT_D, FIT_K, FIT_CLASSES = 576, 16, 20
KERNELS_PER_STEP = 4            # plan_geo_kernel, items_kernel, score_kernel, logit_kernel (+ one small memset)


# ----------------------------------------------------------------------------------------- helpers
def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe: the nvidia-smi query, read through
    NVML in-process when pynvml is there)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _nvml(self):
        """In-process NVML handle (no nvidia-smi process per sample: starting one every 100 ms perturbs the host-side
        legs of the measurement); None when pynvml is unavailable."""
        try:
            import pynvml
            pynvml.nvmlInit()
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            return None

    def _run(self):
        nv = self._nvml()
        while not self._stop.is_set():
            try:
                if nv is not None:
                    pynvml, h = nv
                    sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
                    try:
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    flag = lambda bit: "Active" if r & bit else "Not Active"
                    # NVML bit masks: sw power cap 0x4, hw slowdown 0x8, sw thermal 0x20, hw thermal 0x40
                    self.rows.append([str(sm), str(mx), flag(0x8), flag(0x40), flag(0x20), flag(0x4)])
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05 if nv is not None else 0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def stop(self):
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


def _traffic_channels_last(key):
    """DRAM bytes per launch of the fused path on channels-last maps: the gather's own capture + the plan / score kernels."""
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(tp) as f:
            c = json.load(f)[key]
        b, cl = c["breakdown_bytes"], c["channels_last"]
        return int(cl["items_nhwc_kernel_read"] + cl["items_nhwc_kernel_write"] + b["plan_geo_kernel"] + b["score_kernel_read"])
    except Exception:
        return None


def _traffic(key, field):
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            return json.load(f).get(key, {}).get(field)
    return None


# ------------------------------------------------------------------------------------ workload setup
def device_maps(wl, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = []
    for c, hw in zip(wl.channels, wl.map_hw):          # x = silu(1.2 * randn), fp32 NCHW contiguous
        out.append(torch.nn.functional.silu(1.2 * torch.randn((wl.batch, c, hw, hw), device=device, generator=g)).contiguous())
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
    pooled = ops.normalize_rows(ops.roi_pool(tb)).cpu().numpy()   # rows are zero beyond C_s: the norm is the stride's own
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


FIT_VARIANTS = {
    # "separated": mixture of FIT_K Gaussians, centre spread 6 sigma*sqrt(D): strict convergence in a few Lloyd iterations;
    #              carries the bit-exact-labels claim.
    # "realistic": normalize(silu(1.2 * (low-rank signal + noise))) -- a continuous 8-dimensional structure like pooled
    #              post-SiLU activations, no separable blobs: sklearn needs 200-300 iterations on it, max_iter = 100 caps
    #              every segment at 100 (50-100 iterations); carries the throughput number.
    "separated": dict(sep=6.0, max_iter=300),
    "realistic": dict(rank=8, noise=0.3, max_iter=100),
}


def fit_segment_rows(c, a, cnt, device, variant, seed=77):
    """Rows [a, a + cnt) of class c of the C3-shaped fit set (unit-norm float32 vectors).  Generated per 4096-row unit from
    a seed of (class, unit): the data set does not depend on how many ranks share it."""
    import torch
    from ood_in_object_detection_b200 import kmeans
    unit = kmeans.BLOCK_ROWS * kmeans.SUPER_BLOCKS
    cfg = FIT_VARIANTS[variant]
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000 + c)
    if variant == "separated":
        centers = torch.randn((FIT_K, FIT_D), device=device, generator=g) * cfg["sep"] / FIT_D ** 0.5 + 1.0 / FIT_D ** 0.5
    else:
        basis = torch.randn((cfg["rank"], FIT_D), device=device, generator=g) / cfg["rank"] ** 0.5
    parts = []
    u0, u1 = a // unit, ((a + cnt + unit - 1) // unit if cnt else a // unit)
    for u in range(u0, u1):
        g2 = torch.Generator(device=device)
        g2.manual_seed((seed * 100000 + c) * 4096 + u)
        if variant == "separated":
            lab = torch.randint(0, FIT_K, (unit,), device=device, generator=g2)
            xu = centers[lab] + torch.randn((unit, FIT_D), device=device, generator=g2) / FIT_D ** 0.5
        else:
            z = torch.randn((unit, cfg["rank"]), device=device, generator=g2)
            xu = torch.nn.functional.silu(1.2 * (z @ basis + cfg["noise"] * torch.randn((unit, FIT_D), device=device, generator=g2)))
        lo, hi = max(a, u * unit) - u * unit, min(a + cnt, (u + 1) * unit) - u * unit
        parts.append(xu[lo:hi])
    x = torch.cat(parts) if parts else torch.zeros((0, FIT_D), device=device)
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


def fit_data(n_total, world, rank, device, seed=77, variant="separated"):
    """C3-shaped fit set: FIT_CLASSES segments.  Rank r generates exactly the rows kmeans.shard_rows gives it."""
    import torch
    from ood_in_object_detection_b200 import kmeans
    per = n_total // FIT_CLASSES
    sizes = [per] * FIT_CLASSES
    shard = kmeans.shard_rows(sizes, world, rank)
    parts = [fit_segment_rows(c, a, cnt, device, variant, seed) for c, (a, cnt) in enumerate(shard)]
    return torch.cat(parts).contiguous(), sizes, [cnt for _, cnt in shard]


# ------------------------------------------------------------------------------------- CPU reference
_REF = {}


def load_reference():
    """The reference's own modules (oracle/ref_shim.py: /root/reference in the build container, the byte-compiled
    oracle/_ref on the GPU box) or None when neither is there -- then the loop-for-loop port oracle/cpu_path.py is timed."""
    if "ref" not in _REF:
        _REF["ref"], _REF["shim"] = None, None
        try:
            from oracle import ref_shim
            if ref_shim.available():
                _REF["ref"], _REF["shim"] = ref_shim.load(), ref_shim
        except Exception as e:                          # a broken compiled build must not take the bench down
            print(f"bench: the reference could not be imported ({type(e).__name__}: {e}); timing the port", file=sys.stderr)
    return _REF["ref"], _REF["shim"]


def cpu_reference_pass(det, maps_cpu, wl, clusters, thr, lthr, images, want_decisions=False):
    """One pass of the reference's CPU path over `images` (indices): L1 + cosine FMap decisions and MSP + Energy logit
    decisions, per image / stride / box like ood_utils.py:2038-2180.  Runs the REFERENCE'S OWN classes when they can be
    imported (kind "reference"), else the port.  Returns (#boxes, seconds, kind[, decisions])."""
    import logging
    import torch
    ref, shim = load_reference()
    n_boxes = sum(len(det["boxes"][i]) for i in images)
    slot = {"l1": 0, "l2": 1, "cosine": 2}
    if ref is not None:
        ou = ref.ood_utils
        log = logging.getLogger("bench-ref")
        log.setLevel(logging.CRITICAL)
        b6 = [torch.from_numpy(np.concatenate([det["boxes"][i], det["conf"][i][:, None], det["cls"][i][:, None]], 1).astype(np.float32))
              for i in images]
        res_f = shim.make_results(ref, [[m[i] for m in maps_cpu] for i in images], b6,
                                  strides=[torch.from_numpy(det["strides"][i]) for i in images], batch_hw=(wl.img, wl.img), n_batch=len(images))
        res_l = shim.make_results(ref, None, b6, logits=[torch.from_numpy(det["logits"][i]) for i in images],
                                  batch_hw=(wl.img, wl.img), n_batch=len(images))
        kw = dict(shim.DIST_KW, cluster_method="KMeans_10")
        dm = {"l1": ou.L1DistanceOneClusterPerStride(**kw), "cosine": ou.CosineDistanceOneClusterPerStride(**kw)}
        lm = {"MSP": ou.MSP(**shim.LOGIT_KW), "Energy": ou.Energy(temper=1, **shim.LOGIT_KW)}
        for k, m in dm.items():
            m.clusters, m.thresholds = clusters, thr[slot[k]]
        for k, m in lm.items():
            m.thresholds = lthr[{"MSP": 0, "Energy": 1}[k]].tolist()
        out = {}
        t0 = time.perf_counter()
        for metric in FMAP_METRICS:
            out[metric] = dm[metric].compute_ood_decision_on_results(res_f, log)
        for name in LOGIT_METHODS:
            if name != "MaxLogit":                     # no reference implementation exists
                out[name] = lm[name].compute_ood_decision_on_results(res_l, log)
        dt = time.perf_counter() - t0
        return (n_boxes, dt, "reference", out) if want_decisions else (n_boxes, dt, "reference")
    from oracle import cpu_path
    ims = [dict(maps=[m[i] for m in maps_cpu], boxes=torch.from_numpy(det["boxes"][i]),
                cls=torch.from_numpy(det["cls"][i]), strides=torch.from_numpy(det["strides"][i]),
                logits=torch.from_numpy(det["logits"][i]), img_hw=(wl.img, wl.img)) for i in images]
    out = {}
    t0 = time.perf_counter()
    for metric in FMAP_METRICS:
        out[metric] = cpu_path.distance_decisions(ims, clusters, thr[slot[metric]], metric)
    for name in LOGIT_METHODS:
        if name != "MaxLogit":
            out[name] = cpu_path.logit_decisions(ims, name, lthr[{"MSP": 0, "Energy": 1}[name]].tolist(), 1.0)
    dt = time.perf_counter() - t0
    return (n_boxes, dt, "port", out) if want_decisions else (n_boxes, dt, "port")


def _best_label_agreement(a, b, k):
    """Share of rows on which two clusterings agree after the best one-to-one relabelling (Hungarian on the contingency
    table), and the adjusted Rand index."""
    from scipy.optimize import linear_sum_assignment
    from sklearn.metrics import adjusted_rand_score
    tab = np.zeros((k, k), np.int64)
    np.add.at(tab, (a, b), 1)
    r, c = linear_sum_assignment(-tab)
    return float(tab[r, c].sum() / len(a)), float(adjusted_rand_score(a, b))


def cpu_fit_baseline(n=40000, variant="separated", device=None):
    """sklearn KMeans(K=16, random_state=10) -- the call behind cluster_utils.py:62-73 -- on one bounded segment (class 0 of
    the bench's fit set, first n rows).  With `device`: the same segment through the GPU fit as well, and its agreement /
    ARI / inertia against sklearn next to sklearn's own 1-thread vs all-threads noise floor (its E-step BLAS summation
    order depends on the thread count)."""
    import torch
    from sklearn.cluster import KMeans
    max_iter = FIT_VARIANTS[variant]["max_iter"]
    if device is not None:
        x = fit_segment_rows(0, 0, n, device, variant).cpu().numpy()
    else:                                                # reference arm on a box without touching the GPU
        x = fit_segment_rows(0, 0, n, torch.device("cpu"), variant).numpy()
    t0 = time.perf_counter()
    km = KMeans(n_clusters=FIT_K, random_state=10, max_iter=max_iter).fit(x)
    dt = time.perf_counter() - t0
    out = {"value": n * km.n_iter_ / dt, "unit": FIT_UNIT, "cores": torch.get_num_threads(), "kind": "reference",
           "sample": f"sklearn KMeans(n_clusters={FIT_K}, random_state=10, max_iter={max_iter}).fit (the call of "
                     f"cluster_utils.py:62-73) on one '{variant}' segment of {n} x {FIT_D} vectors "
                     f"({km.n_iter_} Lloyd iterations, {dt:.2f} s incl. k-means++ seeding)",
           "e2e_vectors_per_s": n / dt, "variant": variant}
    if device is not None:
        import threadpoolctl
        from ood_in_object_detection_b200 import kmeans
        with threadpoolctl.threadpool_limits(limits=1):
            km1 = KMeans(n_clusters=FIT_K, random_state=10, max_iter=max_iter).fit(x)
        r = kmeans.kmeans_fit_predict_single(torch.from_numpy(x).to(device), [n], FIT_K, random_state=10, max_iter=max_iter)
        lab = r.labels.cpu().numpy()
        cent = r.centers.cpu().numpy()[0]
        inertia = float(((x.astype(np.float64) - cent[lab].astype(np.float64)) ** 2).sum())
        agree, ari = _best_label_agreement(lab, km.labels_, FIT_K)
        agree0, ari0 = _best_label_agreement(km1.labels_, km.labels_, FIT_K)
        out["parity"] = {"gpu_vs_sklearn": {"label_agreement": agree, "ari": ari, "inertia_ratio": inertia / float(km.inertia_),
                                            "n_iter_gpu": int(r.n_iter[0]), "n_iter_sklearn": int(km.n_iter_)},
                         "sklearn_1_thread_vs_all_threads": {"label_agreement": agree0, "ari": ari0,
                                                            "inertia_ratio": float(km1.inertia_) / float(km.inertia_),
                                                            "n_iter": [int(km1.n_iter_), int(km.n_iter_)]}}
    return out


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (loop-for-loop port, same third-party
    calls) on this box's host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from ood_in_object_detection_b200 import synth
    # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin the reference
    # arm to one thread whenever it is launched the way the N > 1 runs are
    ncpu = os.cpu_count() or 1
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    try:
        import threadpoolctl
        _limits = threadpoolctl.threadpool_limits(limits=ncpu)      # BLAS / OpenMP pools of numpy, scipy, sklearn
    except Exception:                                               # the pools keep their environment defaults
        _limits = None
    if args.workload == "fit":
        vals = [cpu_fit_baseline(variant=args.fit_variant) for _ in range(max(min(args.steps, 3), 1))]
        v = float(np.mean([x["value"] for x in vals]))
        cb = dict(vals[-1], value=v)
        print(json.dumps({"impl": "reference", "metric": FIT_METRIC, "value": v, "unit": FIT_UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"C3 k-means fit ('{args.fit_variant}' set, bounded CPU sample)", "dim": FIT_D, "k": FIT_K},
                          "cpu_baseline": cb, "e2e": {"value": v, "unit": FIT_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    n_img = CPU_SAMPLE_IMAGES
    det = synth.detections(2000, wl.batch, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)     # the GPU arm's detections (rank 0)
    g = torch.Generator()
    g.manual_seed(1000)
    maps = [torch.nn.functional.silu(1.2 * torch.randn((n_img, c, hw, hw), generator=g)) for c, hw in zip(wl.channels, wl.map_hw)]
    rng = np.random.default_rng(0)
    clusters = [[np.abs(rng.standard_normal((wl.k, c))).astype(np.float32) / np.sqrt(c) for c in wl.channels]
                for _ in range(wl.nc)]
    thr = {m: [[1.0] * 3 for _ in range(wl.nc)] for m in range(3)}
    lthr = np.zeros((5, wl.nc))
    imgs = list(range(n_img))
    for _ in range(max(args.warmup, 1)):                # imports and thread pools are not part of the metric
        cpu_reference_pass(det, maps, wl, clusters, thr, lthr, imgs[:1])
    tot_n, tot_t, kind = 0, 0.0, "port"
    for _ in range(args.steps):
        n, dt, kind = cpu_reference_pass(det, maps, wl, clusters, thr, lthr, imgs)
        tot_n += n
        tot_t += dt
    v = tot_n / tot_t
    sample = (f"{n_img} images (~{tot_n // max(args.steps, 1)} boxes) of the {wl.batch}-image batch per step; L1+cosine FMap "
              f"and MSP+Energy logit decisions through " + ("the reference's own compute_ood_decision_on_results (ood_utils.py:2038-2180, "
              ":1195-1208; byte-compiled build oracle/_ref)" if kind == "reference" else "oracle/cpu_path.py (loop-for-loop port, same "
              "sklearn/torchvision calls)"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "sample_images": n_img},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------ fit arm
def _fit_once(x, gsizes, lsizes, world, rank, group, max_iter, reduce="allreduce"):
    """seed + Lloyd + member means + fit scores + exact thresholds of the C3-shaped set (the fit stage of ood_utils)."""
    from ood_in_object_detection_b200 import kmeans, ops, select
    if world > 1:
        r = kmeans.kmeans_fit_sharded(x, lsizes, gsizes, FIT_K, world, rank, group, random_state=10, max_iter=max_iter, reduce=reduce)
    else:
        r = kmeans.kmeans_fit_predict_single(x, gsizes, FIT_K, random_state=10, max_iter=max_iter, reduce=reduce)
    kw = dict(reduce="ordered", global_sizes=gsizes) if reduce == "ordered" else {}
    means, counts = kmeans.member_means(x, lsizes, r.labels, FIT_K, group=group, **kw)
    off = np.concatenate([[0], np.cumsum(lsizes)]).tolist()
    d, _ = ops.vec_score_one(x, off, means.reshape(-1, FIT_D).contiguous(), None, [g * FIT_K for g in range(FIT_CLASSES)],
                             [FIT_K] * FIT_CLASSES, ops.METRIC_SLOT["l2"], normalize=False)      # FP32: the decision path's arithmetic
    ranks = [select.lower_index(n, 95.0) for n in gsizes]
    thr, _, _ = select.segment_select(d[ops.METRIC_SLOT["l2"]].contiguous(), off, ranks, group=group)
    return r, means, thr


def verify_against_single_gpu(device, world, rank, n_total, variant):
    """SURVEY.md section 4 (iii) on hardware: the N-rank fit (reduce="ordered") against the same fit on ONE GPU (rank 0
    holds the whole set): labels identical on every rank's rows, centres / member means / thresholds bit-equal."""
    import torch
    import torch.distributed as dist
    from ood_in_object_detection_b200 import kmeans
    max_iter = FIT_VARIANTS[variant]["max_iter"]
    x, gsizes, lsizes = fit_data(n_total, world, rank, device, variant=variant)
    group = dist.group.WORLD
    r, means, thr = _fit_once(x, gsizes, lsizes, world, rank, group, max_iter, reduce="ordered")
    del x
    full = torch.empty(n_total, dtype=torch.int32, device=device)
    same = torch.zeros(3, dtype=torch.int32, device=device)
    if rank == 0:
        x1, _, _ = fit_data(n_total, 1, 0, device, variant=variant)
        r1, means1, thr1 = _fit_once(x1, gsizes, gsizes, 1, 0, None, max_iter, reduce="ordered")
        del x1
        full.copy_(r1.labels)
        same[0] = int(torch.equal(r1.centers, r.centers))
        same[1] = int(torch.equal(means1, means))
        same[2] = int(all((a == b) or (a is None and b is None) for a, b in zip(thr, thr1)))
        iters1 = [int(v) for v in r1.n_iter]
    dist.broadcast(full, src=0)
    dist.broadcast(same, src=0)
    goff = np.concatenate([[0], np.cumsum(gsizes)])
    shard = kmeans.shard_rows(gsizes, world, rank)
    mine = torch.cat([full[int(goff[g]) + a:int(goff[g]) + a + cnt] for g, (a, cnt) in enumerate(shard)])
    bad = (mine != r.labels).sum().to(torch.int64)
    dist.all_reduce(bad)
    torch.cuda.empty_cache()
    out = {"labels_differing": int(bad), "centres_bit_equal": bool(same[0]), "member_means_bit_equal": bool(same[1]),
           "thresholds_bit_equal": bool(same[2]), "reduce": "ordered", "variant": variant, "n_vectors": n_total,
           "lloyd_iterations": max(int(v) for v in r.n_iter)}
    if rank == 0:
        out["lloyd_iterations_single_gpu"] = max(iters1)
    out["matches"] = out["labels_differing"] == 0 and all(out[k] for k in ("centres_bit_equal", "member_means_bit_equal", "thresholds_bit_equal"))
    return out


def run_fit(device, world, rank, n_total, reps, warm, variant="realistic"):
    """Segmented k-means fit + thresholds; returns a dict (rank-0 meaningful).  The Lloyd loop and the seeding are timed
    inside kmeans_fit (device synchronised on both sides); the end-to-end figure is the wall clock around the whole fit
    (seeding has host decisions), max over ranks."""
    import torch
    import torch.distributed as dist
    x, gsizes, lsizes = fit_data(n_total, world, rank, device, variant=variant)
    max_iter = FIT_VARIANTS[variant]["max_iter"]
    group = dist.group.WORLD if world > 1 else None
    best = None
    for it in range(warm + reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r, means, thr = _fit_once(x, gsizes, lsizes, world, rank, group, max_iter)
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        t = torch.tensor([total, r.seconds["lloyd"], r.seconds["init"], r.seconds["center"]], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cur = dict(total=float(t[0]), lloyd=float(t[1]), init=float(t[2]), center=float(t[3]), seeding=r.seconds.get("seeding"),
                   iters=int(r.seconds["lloyd_iters"]), issued=int(r.seconds.get("lloyd_issued", 0)),
                   n_iter=[int(v) for v in r.n_iter], strict=all(r.strict), thr0=thr[0], collective=r.seconds.get("collective"))
        if it >= warm and (best is None or cur["total"] < best["total"]):
            best = cur
    rows = torch.tensor([x.shape[0]], dtype=torch.int64, device=device)
    rows_max = rows.clone()
    if world > 1:
        dist.all_reduce(rows_max, op=dist.ReduceOp.MAX)
    del x
    torch.cuda.empty_cache()
    peak, _ = _peaks()
    it_bytes = 4.0 * n_total * FIT_D + 4.0 * n_total + 2 * 4.0 * FIT_CLASSES * FIT_K * FIT_D
    # vector-iterations actually computed: segments that converged early drop out of the later iterations
    vec_iters = float(sum(n * i for n, i in zip(gsizes, best["n_iter"])))
    lloyd_per_iter = best["lloyd"] / max(best["iters"], 1)
    return {"metric": FIT_METRIC, "value": vec_iters / best["lloyd"], "unit": FIT_UNIT, "variant": variant,
            "n_vectors": n_total, "dim": FIT_D, "k": FIT_K, "segments": FIT_CLASSES, "lloyd_iterations": best["iters"],
            "lloyd_iterations_per_segment": best["n_iter"], "lloyd_iterations_issued": best["issued"],
            "lloyd_ms_per_iteration": 1e3 * lloyd_per_iter, "seed_ms": 1e3 * best["init"], "seeding": best["seeding"],
            "center_ms": 1e3 * best["center"],
            "means_scores_thresholds_ms": 1e3 * (best["total"] - best["init"] - best["lloyd"] - best["center"]),
            "fit_ms": 1e3 * best["total"], "rows_on_fullest_rank": int(rows_max),
            "e2e_vectors_per_s": n_total / best["total"], "strict_convergence": best["strict"], "scaling": "strong",
            "collective": (f"per Lloyd iteration: {best['collective']}; per radix pass: 1 NCCL all-reduce of histograms") if world > 1 else "none (1 GPU)",
            "roofline": {"bound": "hbm", "kernel": "kmeans_step_tc_kernel (tcgen05 assignment + partial sums; + reduce/update, host loop)",
                         "unit": "GB/s",
                         "achieved": it_bytes * (vec_iters / n_total / max(best["iters"], 1)) / world / lloyd_per_iter / 1e9, "peak": peak,
                         "frac": it_bytes * (vec_iters / n_total / max(best["iters"], 1)) / world / lloyd_per_iter / 1e9 / peak,
                         "algorithmic_bytes_per_iteration": it_bytes,
                         "traffic": (_traffic("C3", "kmeans_step_dram_bytes_per_vector") or 0) * n_total / world or None}}


# ------------------------------------------------------------------------------------- other BASELINE configs
def _graph_kernel_nodes(g):
    """Kernel nodes of a captured CUDA graph (torch keeps the cudaGraph_t with keep_graph=True), counted through the CUDA
    runtime's own graph API: what one replay launches.  None when the bindings are not there."""
    try:
        from cuda.bindings import runtime as rt
        raw = g.raw_cuda_graph()
        err, _, n = rt.cudaGraphGetNodes(raw, 0)
        if int(err) != 0 or not n:
            return None
        err, nodes, n = rt.cudaGraphGetNodes(raw, n)
        if int(err) != 0:
            return None
        kinds = [rt.cudaGraphNodeGetType(nd) for nd in nodes[:n]]
        return sum(1 for e, t in kinds if int(e) == 0 and t == rt.cudaGraphNodeType.cudaGraphNodeTypeKernel)
    except Exception:
        return None


def _time_graph(fn, steps, flush):
    """CUDA-event time per replay of `fn` captured as one CUDA graph, L2 flushed between replays."""
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        g.replay()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps


def other_config(name, device, flush, steps=10):
    """Resident-input scoring of another BASELINE.json config on this GPU (C1: YOLOv8n B=8 vanilla L2, K=1; C4: YOLOv8l
    B=32 per GPU, K=10, Cosine + MSP with and/or/score fusion; C5: YOLOv8x 1280x1280, 300 boxes per image, K=64, FP32 fused
    path next to the tcgen05 cross-term path).  Same timing rules as the main line."""
    import torch
    from ood_in_object_detection_b200 import ops, synth
    wl = synth.CONFIGS[name]
    maps = device_maps(wl, 1100, device)
    det = synth.detections(2100, wl.batch, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)
    clusters, thr, table, lthr = fit_tables(ops, wl, maps, 3100, device)
    batch = ops.make_batch(maps, det["boxes"], det["strides"], det["cls"], wl.img, device)
    n = batch.n
    metrics = {"C1": ("l2",), "C4": ("cosine",), "C5": ("l2", "cosine")}[name]
    mask = sum(1 << ops.METRIC_SLOT[m] for m in metrics)
    out = ops.alloc_fmap_scores(n, device)
    alg, upper, _ = algorithmic_bytes(det, wl, wl.k, n_tables=len(metrics))
    peak, _ = _peaks()
    res = {"workload": wl.name, "boxes": n, "fmap_metrics": list(metrics), "k_per_class_stride": wl.k}
    if name == "C4":                                    # + MSP and the three fusion rules in the same step
        logits = torch.from_numpy(np.concatenate(det["logits"])).to(device)
        lthr_d = torch.from_numpy(lthr).to(device)
        lout = ops.LogitScores(scores=torch.zeros((5, n), dtype=torch.float32, device=device),
                               indness=torch.zeros((5, n), dtype=torch.float32, device=device),
                               decision=torch.ones((5, n), dtype=torch.uint8, device=device),
                               sigmoid_mismatch=torch.zeros(1, dtype=torch.int32, device=device))
        smin = torch.zeros((5, wl.nc), dtype=torch.float64, device=device)
        smax = torch.ones((5, wl.nc), dtype=torch.float64, device=device)
        neg1 = torch.full((n,), -1.0, dtype=torch.float32, device=device)       # the reference's distance INDness (Q2)

        def step():
            ops.fmap_score(batch, table, mask, True, compat_q1=True, out=out)
            ops.logit_score(logits, batch.cls, 1 << ops.LOGIT_SLOT["MSP"], thr=lthr_d, smin=smin, smax=smax, out=lout)
            ops.fuse_decisions(lout.decision[0], out.decision[2], "and")
            ops.fuse_decisions(lout.decision[0], out.decision[2], "or")
            ops.fuse_scores(lout.indness[0], neg1)
        res["fusion"] = "fusion-MSP-Cosine_cl_stride: and / or / score in the same step"
    else:
        def step():
            ops.fmap_score(batch, table, mask, True, compat_q1=True, out=out)
    ms = _time_graph(step, steps, flush)
    res.update({"ms_per_step": ms, "value": n / (ms * 1e-3), "unit": UNIT, "algorithmic_bytes": alg,
                "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak})
    if name == "C5":
        # the dense-contraction variant (BASELINE.json configs[4]): pool -> rows grouped by (stride, class used) -> normalise ->
        # x.c cross-term on tcgen05 (vec_score_tc) for l2 and cosine
        pooled = ops.roi_pool(batch)
        cls_used, out_index = ops.q1_plan(batch)
        st_h = batch.stride_idx.cpu().numpy()
        cu_h = cls_used.cpu().numpy()
        oi = out_index.long()
        groups = []
        for s in range(3):
            idx = np.nonzero(st_h == s)[0]
            order = idx[np.argsort(cu_h[idx], kind="stable")]
            seg_off = np.searchsorted(cu_h[order], np.arange(wl.nc + 1)).tolist()
            cents = [np.asarray(clusters[c][s], np.float32).reshape(-1, wl.channels[s]) for c in range(wl.nc)]
            ks = [len(a) for a in cents]
            crow = np.concatenate([[0], np.cumsum(ks)])[:-1].tolist()
            cent = torch.from_numpy(np.concatenate(cents)).to(device)
            unit = torch.from_numpy(ops._unit_rows(np.concatenate(cents))).to(device)
            tthr = torch.full((3, wl.nc), float("nan"), dtype=torch.float64, device=device)
            for m in metrics:
                tthr[ops.METRIC_SLOT[m]] = torch.tensor([thr[ops.METRIC_SLOT[m]][c][s] if thr[ops.METRIC_SLOT[m]][c][s] != [] else float("nan")
                                                         for c in range(wl.nc)], dtype=torch.float64)
            groups.append((torch.from_numpy(order).to(device), seg_off, cent, unit, crow, ks, tthr, wl.channels[s]))
        tc_dec = torch.zeros((3, n), dtype=torch.uint8, device=device)
        tc_arg = torch.zeros((3, n), dtype=torch.int32, device=device)

        def tensor_step():
            pooled_now = ops.roi_pool(batch, out=pooled)
            for order, seg_off, cent, unit, crow, ks, tthr, c_s in groups:
                x = ops.normalize_rows(pooled_now.index_select(0, order)[:, :c_s].contiguous())
                for m in metrics:
                    d, a, de = ops.vec_score_tc(x, seg_off, unit if m == "cosine" else cent, crow, ks, m, thr=tthr)
                    slot = ops.METRIC_SLOT[m]
                    tc_dec[slot].index_copy_(0, oi.index_select(0, order), de[slot])
                    tc_arg[slot].index_copy_(0, oi.index_select(0, order), a[slot])
        for _ in range(2):
            tensor_step()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()
            a.record()
            tensor_step()
            b.record()
        torch.cuda.synchronize()
        tms = sum(a.elapsed_time(b) for a, b in evs) / steps
        step()
        torch.cuda.synchronize()
        slots = [ops.METRIC_SLOT[m] for m in metrics]
        kmax = max(max(len(np.atleast_2d(clusters[c][s])) if np.size(clusters[c][s]) else 0 for s in range(3)) for c in range(wl.nc))
        flops = float(sum(2.0 * wl.k * wl.channels[int(s)] for s in st_h)) * len(metrics) * 3      # split-float: 3 tensor-core products
        res["tensor_path"] = {"ms_per_step": tms, "value": n / (tms * 1e-3),
                              "what": "roi_pool + per-stride gather / normalise + vec_score_tc (tcgen05 kind::tf32, split-float x, K <= 64) "
                                      "for l2 and cosine, launched eagerly (index_select / index_copy plumbing included)",
                              "decisions_differing_from_fp32": int((tc_dec[slots] != out.decision[slots]).sum()),
                              "argmin_differing_from_fp32": int((tc_arg[slots] != out.argmin[slots]).sum()),
                              "tensor_flops_per_step": flops, "k": kmax,
                              "note": "the fused FP32 pass is the faster one at 4800 boxes: the contraction is 2*K*C flop per box = "
                                      f"{flops / 1e9:.2f} GFLOP per step, far below what keeps the tensor pipe busy; tcgen05 pays off on the "
                                      "fit side (millions of rows: Lloyd step, fit scores)"}
    del maps, batch, out
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from ood_in_object_detection_b200 import ood_utils, ops, synth
    from ood_in_object_detection_b200.results import Results, batch_shape

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    clk = ClockSampler(local).start()

    if args.workload == "fit":
        fit = run_fit(device, world, rank, args.fit_n or 4_000_000, max(min(args.steps, 5), 1), 1, variant=args.fit_variant)
        if world > 1:
            fit["matches_single_gpu"] = verify_against_single_gpu(device, world, rank, args.fit_n or 4_000_000, args.fit_variant)
        clk.stop()
        if rank == 0:
            cb = cpu_fit_baseline(variant=args.fit_variant, device=device)
            out = {"metric": FIT_METRIC, "value": fit["value"], "unit": FIT_UNIT, "n_gpus": world, "steps": args.steps,
                   "warmup": args.warmup, "ms_per_step": fit["fit_ms"], "higher_is_better": True, "scaling": "strong",
                   "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                   "config": {"workload": f"C3 k-means fit N={fit['n_vectors']} D={FIT_D} K={FIT_K} x {FIT_CLASSES} classes, '{args.fit_variant}' set",
                              "sharding": f"rows of every segment over {world} rank(s)"},
                   "roofline": fit["roofline"], "cpu_baseline": cb, "fit": fit,
                   "e2e": {"value": fit["e2e_vectors_per_s"], "unit": "vectors/s (seed + Lloyd + member means + scores + thresholds)",
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "note": "activations are produced on the device by the pooling kernel; nothing crosses PCIe"},
                   # per Lloyd iteration: centroid prep, tcgen05 step, 2 x reduce, update; per seeded centre: scan, gather, distance
                   # pass, potentials, pick; plus centring (4), member means (3), scores (1), 3 radix passes
                   "gpu_launches": (5 * fit["lloyd_iterations"] + 5 * FIT_K + 11) * args.steps, "clocks": clk.summary()}
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return

    maps = device_maps(wl, 1000 + rank, device)
    det = synth.detections(2000 + rank, wl.batch, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)
    clusters, thr, table, lthr = fit_tables(ops, wl, maps, 3000 + rank, device)
    fmask = sum(1 << ops.METRIC_SLOT[m] for m in FMAP_METRICS)
    lmask = sum(1 << ops.LOGIT_SLOT[m] for m in LOGIT_METHODS)
    lthr_d = torch.from_numpy(lthr).to(device)

    maps_cl = [m.contiguous(memory_format=torch.channels_last) for m in maps]    # same values, [B, H, W, C] memory
    if args.layout == "nhwc":
        maps, maps_cl = maps_cl, maps
    batch = ops.make_batch(maps, det["boxes"], det["strides"], det["cls"], wl.img, device)
    batch_alt = ops.make_batch(maps_cl, det["boxes"], det["strides"], det["cls"], wl.img, device)
    assert batch.nhwc == (args.layout == "nhwc") and batch_alt.nhwc != batch.nhwc
    logits = torch.from_numpy(np.concatenate(det["logits"])).to(device)
    n = batch.n
    fout = ops.alloc_fmap_scores(n, device)
    lout = ops.LogitScores(scores=torch.zeros((5, n), dtype=torch.float32, device=device), indness=None,
                           decision=torch.ones((5, n), dtype=torch.uint8, device=device),
                           sigmoid_mismatch=torch.zeros(1, dtype=torch.int32, device=device))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)     # > 126 MB L2

    def fused():
        ops.fmap_score(batch, table, fmask, True, compat_q1=True, out=fout)     # memset + plan + geo + items + score

    fout_alt = ops.alloc_fmap_scores(n, device)

    def fused_alt():                                                             # the same pass on the other map layout
        ops.fmap_score(batch_alt, table, fmask, True, compat_q1=True, out=fout_alt)

    def logit():
        ops.logit_score(logits, batch.cls, lmask, thr=lthr_d, out=lout)

    for _ in range(max(args.warmup, 3)):                                        # also sizes the workspace before capture
        flush.zero_()
        fused()
        fused_alt()
        logit()
    torch.cuda.synchronize()
    # the step = ONE graph: the fused FMap path on the capture stream, the (independent) logit methods on a forked branch
    g_step, g_fused = torch.cuda.CUDAGraph(keep_graph=True), torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=device)
    with torch.cuda.graph(g_step):
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            logit()
        fused()
        cur.wait_stream(side)
    kernels_per_step = _graph_kernel_nodes(g_step)                              # counted, not assumed
    g_step.instantiate()
    with torch.cuda.graph(g_fused):                                             # the fused path alone: roofline timing
        fused()
    g_alt = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_alt):
        fused_alt()
    for _ in range(2):
        g_step.replay()
        g_fused.replay()
        g_alt.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    E = lambda: torch.cuda.Event(enable_timing=True)
    evs = [(E(), E()) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_wall = time.perf_counter()
    for a, c in evs:
        flush.zero_()                                                # L2 flush between timed iterations
        a.record()
        g_step.replay()
        c.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall
    total_ms = sum(a.elapsed_time(c) for a, c in evs)
    evf = [(E(), E()) for _ in range(args.steps)]                    # same launches without the logit branch (not in `value`)
    for a, b in evf:
        flush.zero_()
        a.record()
        g_fused.replay()
        b.record()
    torch.cuda.synchronize()
    fmap_ms = sum(a.elapsed_time(b) for a, b in evf) / max(args.steps, 1)
    eva = [(E(), E()) for _ in range(args.steps)]                    # the fused path on the other map layout (reported beside)
    for a, b in eva:
        flush.zero_()
        a.record()
        g_alt.replay()
        b.record()
    torch.cuda.synchronize()
    alt_ms = sum(a.elapsed_time(b) for a, b in eva) / max(args.steps, 1)
    alt_name = "nchw" if args.layout == "nhwc" else "channels_last"
    slots = [ops.METRIC_SLOT[m] for m in FMAP_METRICS]                # only the requested metric slots are written
    alt_same = int((fout.decision[slots] != fout_alt.decision[slots]).sum()) + int((fout.argmin[slots] != fout_alt.argmin[slots]).sum())
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
        clk.stop()
        if rank == 0:
            alg, upper, _ = algorithmic_bytes(det, wl, wl.k)
            print(json.dumps({"value": value, "ms_per_step": total_ms / args.steps, "fmap_ms": fmap_ms,
                              "frac": alg / (fmap_ms * 1e-3) / 1e9 / _peaks()[0], alt_name + "_fmap_ms": alt_ms,
                              alt_name + "_frac": alg / (alt_ms * 1e-3) / 1e9 / _peaks()[0],
                              "decisions_or_argmin_differing": alt_same, "quick": True}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the class surface: host (pinned) inputs -> decisions on the host, every step
    KW = dict(agg_method="mean", cluster_method=f"KMeans_{wl.k}", cluster_optimization_metric="silhouette",
              ind_info_creation_option="valid_preds_one_stride", which_internal_activations="ftmaps_and_strides",
              iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)
    LKW = dict(per_class=True, per_stride=False, iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15,
               min_conf_threshold_test=0.15, use_values_before_sigmoid=True)
    m_l1, m_cos = ood_utils.L1DistanceOneClusterPerStride(**KW), ood_utils.CosineDistanceOneClusterPerStride(**KW)
    m_l1.clusters = m_cos.clusters = clusters
    m_l1.thresholds, m_cos.thresholds = thr[0], thr[2]
    m_msp, m_en, m_ml = ood_utils.MSP(**LKW), ood_utils.Energy(temper=1, **LKW), ood_utils.MaxLogit(**LKW)
    m_msp.thresholds, m_en.thresholds, m_ml.thresholds = lthr[0].tolist(), lthr[1].tolist(), lthr[4].tolist()
    methods = [m_l1, m_cos, m_msp, m_en, m_ml]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_maps = [m.cpu().pin_memory() for m in maps]
    shape = batch_shape(wl.batch, wl.img, wl.img)
    res_f, res_l = [], []
    for i in range(wl.batch):
        b6 = np.concatenate([det["boxes"][i], det["conf"][i][:, None], det["cls"][i][:, None]], 1).astype(np.float32)
        res_f.append(Results(orig_img=shape, boxes=pin(b6), extra_item=([hm[i] for hm in h_maps], pin(det["strides"][i]))))
        res_l.append(Results(orig_img=shape, boxes=pin(b6), extra_item=pin(det["logits"][i])))
    h2d = sum(m.numel() * 4 for m in h_maps) + sum(r.boxes.data.numel() * 4 + r.extra_item[1].numel() * 4 for r in res_f) \
        + sum(r.extra_item.numel() * 4 + r.boxes.data.shape[0] * 4 for r in res_l)      # logits + classes, uploaded once
    import logging
    log = logging.getLogger("bench")
    log.setLevel(logging.ERROR)
    e2e_steps = args.e2e_steps or max(2, min(args.steps, 20))

    def e2e_step():
        return ood_utils.compute_ood_decisions_fused(methods, res_f, log, logits_results=res_l)
    dec = e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dec = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    d2h = sum(sum(len(v) for v in d) for d in dec.values())
    # the class surface and the resident path agree
    same = np.array_equal(np.array([v for im in dec[m_cos.name] for v in im], np.uint8), fout.decision[2].cpu().numpy())
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = n_all * e2e_steps / e2e_s
    # the same call with the inputs where a detector leaves them: feature maps and detections already on the device (only the
    # decisions cross PCIe) -- the deployment case; the host-input figure above is bound by the 367 MB map upload per step
    res_fd = [Results(orig_img=shape, boxes=r.boxes.data.to(device), extra_item=([m[i] for m in maps], r.extra_item[1].to(device)))
              for i, r in enumerate(res_f)]
    res_ld = [Results(orig_img=shape, boxes=r.boxes.data.to(device), extra_item=r.extra_item.to(device)) for r in res_l]
    ood_utils.compute_ood_decisions_fused(methods, res_fd, log, logits_results=res_ld)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_steps = max(e2e_steps, 20)
    t0 = time.perf_counter()
    for _ in range(dev_steps):
        dec_d = ood_utils.compute_ood_decisions_fused(methods, res_fd, log, logits_results=res_ld)
    torch.cuda.synchronize()
    dev_s = time.perf_counter() - t0
    same_d = dec_d[m_cos.name] == dec[m_cos.name] and dec_d[m_l1.name] == dec[m_l1.name]
    if world > 1:
        t = torch.tensor([dev_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t[0])
    e2e_dev_value = n_all * dev_steps / dev_s
    del res_fd, res_ld
    # the producer side in front of it: from the detector head's raw output [B, 4 + nc, A] and the hooked maps,
    # both on the device, through postprocess (NMS + payload + Results of views) and the same fused class call, to the host
    from_head = None
    if world == 1:
        try:
            from types import SimpleNamespace as NS
            from ood_in_object_detection_b200.postprocess import postprocess
            head = torch.from_numpy(synth.head_output(7100, wl.batch, wl.nc, wl.img)[0]).to(device)
            pred_like = NS(args=NS(conf=0.25, iou=0.45, agnostic_nms=False, max_det=300, classes=None, model="yolov8s.pt"),
                           model=NS(model=NS(extraction_mode="ftmaps_and_strides", model=[NS(output_values_before_sigmoid=False)]),
                                    names={i: str(i) for i in range(wl.nc)}), batch=None)
            img_t = torch.empty((wl.batch, 3, wl.img, wl.img), device="meta")           # only its shape is read

            def head_step():
                results = postprocess(pred_like, ((head,), maps), img_t, img_t)
                return results, ood_utils.compute_ood_decisions_fused([m_l1, m_cos], results, log)
            results, dec_h = head_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(dev_steps):
                results, dec_h = head_step()
            torch.cuda.synchronize()
            head_s = (time.perf_counter() - t0) / dev_steps
            n_head = sum(len(r.boxes) for r in results)
            from_head = {"value": n_head / head_s, "unit": UNIT, "ms_per_batch": 1e3 * head_s, "detections_kept": n_head,
                         "anchors": int(head.shape[2]), "steps": dev_steps,
                         "path": "head output + maps on the device -> postprocess (NMS with payload, Results) -> fused L1 + cosine decisions -> host"}
            del head, results
        except Exception as e:                              # a side measurement must not take the headline down
            from_head = {"error": f"{type(e).__name__}: {e}"}

    fit = None
    if args.fit_n != 0:
        del h_maps, res_f, res_l
        torch.cuda.empty_cache()
        # BASELINE config 3 at its full size on every N (strong scaling: 4 M x 576, K = 16 x 20 classes): the 'realistic' set
        # (tens of Lloyd iterations) carries the throughput and scaling figures, the 'separated' set the bit-exact-labels claim
        n_fit = args.fit_n or 4_000_000
        fit = run_fit(device, world, rank, n_fit, 2, 1, variant="realistic")
        fit["separated"] = run_fit(device, world, rank, n_fit, 2, 1, variant="separated")
        if world > 1:
            fit["matches_single_gpu"] = verify_against_single_gpu(device, world, rank, n_fit, "separated")
            fit["matches_single_gpu_realistic"] = verify_against_single_gpu(device, world, rank, n_fit, "realistic")
        if rank == 0:
            fit["cpu_baseline"] = cpu_fit_baseline(variant="realistic", device=device)
    others = {}
    if rank == 0 and world == 1 and args.config == "C2" and not args.no_other_configs:
        del batch_alt, maps_cl
        torch.cuda.empty_cache()
        for name in ("C1", "C4", "C5"):
            try:
                others[name] = other_config(name, device, flush)
            except Exception as e:                          # a side measurement must not take the headline down
                others[name] = {"error": f"{type(e).__name__}: {e}"}
    clk.stop()
    if world > 1:                                        # every collective is done: only rank 0 has host work left
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return

    if rank == 0:
        alg, upper, _ = algorithmic_bytes(det, wl, wl.k)
        peak, how = _peaks()
        achieved = alg / (fmap_ms * 1e-3) / 1e9
        cpu_baseline = {"value": None, "unit": UNIT, "note": "the CPU arm is timed at N = 1 only"}
        if world == 1:
            maps_cpu = [m[:CPU_SAMPLE_IMAGES].contiguous(memory_format=torch.contiguous_format).cpu() for m in maps]
            cpu_reference_pass(det, maps_cpu, wl, clusters, thr, lthr, [0])        # warm the imports / thread pools
            cpu_n, cpu_t, cpu_kind, cpu_dec = cpu_reference_pass(det, maps_cpu, wl, clusters, thr, lthr, list(range(CPU_SAMPLE_IMAGES)),
                                                                 want_decisions=True)
            # the CPU arm's decisions on its sample against the CUDA path's decisions on the same boxes (same order: per image,
            # stride-major for the FMap methods, box order for the logit methods)
            n6 = sum(len(det["boxes"][i]) for i in range(CPU_SAMPLE_IMAGES))
            flat = lambda d: np.array([v for im in d for v in im], np.uint8)
            cpu_vs_gpu = {m: int((flat(cpu_dec[m]) != fout.decision[ops.METRIC_SLOT[m]][:n6].cpu().numpy()).sum()) for m in FMAP_METRICS}
            cpu_vs_gpu.update({m: int((flat(cpu_dec[m]) != lout.decision[ops.LOGIT_SLOT[m]][:n6].cpu().numpy()).sum())
                               for m in LOGIT_METHODS if m in cpu_dec})
            cpu_baseline = {"value": cpu_n / cpu_t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": cpu_kind,
                            "sample": f"{CPU_SAMPLE_IMAGES} of {wl.batch} images ({cpu_n} boxes), L1+cosine FMap and MSP+Energy logits through "
                                      + ("the reference's own compute_ood_decision_on_results (byte-compiled build oracle/_ref)"
                                         if cpu_kind == "reference" else "oracle/cpu_path.py (loop-for-loop port, same sklearn/torchvision calls)"),
                            "decisions_differing_from_gpu": cpu_vs_gpu}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "map_layout": args.layout, "images_per_gpu": wl.batch, "boxes_per_gpu": n, "fmap_metrics": FMAP_METRICS,
                       "logit_methods": LOGIT_METHODS, "k_per_class_stride": wl.k, "nc": wl.nc,
                       "l2_flush": "512 MiB memset between timed iterations", "sharding": f"batch x{world}, no collective",
                       "launch": "one CUDA graph per step: fused FMap path + logit methods on a forked branch"},
            "roofline": {"bound": "hbm", "kernel": "items_kernel (window gather) within plan_geo+items+score", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(args.config, "fmap_dram_bytes_per_launch") if args.layout == "nchw" else None,
                         "algorithmic_bytes": alg,
                         "upper_bound_bytes": upper, "kernel_ms": fmap_ms, "peak_source": how, "layout": args.layout,
                         "other_layout": {"layout": alt_name, "kernel_ms": alt_ms, "achieved": alg / (alt_ms * 1e-3) / 1e9,
                                          "frac": alg / (alt_ms * 1e-3) / 1e9 / peak, "decisions_or_argmin_differing": alt_same,
                                          "traffic": _traffic_channels_last(args.config) if alt_name == "channels_last" else None,
                                          "note": "the same fused pass over the same values with the maps in the other memory "
                                                  "layout (channels_last = what a detector run in torch.channels_last hands over)"},
                         "note": "HBM moves whole 128-byte lines; NCHW window rows are 8..52 B (DESIGN.md section 4)"},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "ood_utils.compute_ood_decisions_fused([L1, Cosine, MSP, Energy, MaxLogit], results)",
                    "matches_resident_path": bool(same),
                    "note": "host pinned feature maps + detections -> per-image decision lists on the host",
                    "pcie_bound_value": n_all / (h2d / 55e9),
                    "pcie_note": "upper bound of this figure at the 55 GB/s the box's pinned H2D copies reach (profiles/r1_h2d_bw.log): "
                                 "the boxes' windows are scattered over ~all rows of every plane, a partial upload would need one "
                                 "strided copy per box",
                    "from_head": from_head,
                    "device_inputs": {"value": e2e_dev_value, "unit": UNIT, "steps": dev_steps, "same_decisions": bool(same_d),
                                      "note": "same API call with the feature maps and detections already on the device (where the "
                                              "detector leaves them); D2H of the decisions and all host-side packing included"}},
            "gpu_launches": (kernels_per_step or KERNELS_PER_STEP) * args.steps,
            "gpu_launches_source": ("kernel nodes of the timed CUDA graph (cudaGraphGetNodes) x steps" if kernels_per_step
                                    else "constant (graph API unavailable)"),
            "clocks": clk.summary(), "wall_s_timed_region": t_wall,
        }
        if fit is not None:
            out["fit"] = fit
        if others:
            out["other_configs"] = others
        print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C4", "C5"])
    ap.add_argument("--workload", default="score", choices=["score", "fit"])
    ap.add_argument("--layout", default="nchw", choices=["nchw", "nhwc"],
                    help="memory layout of the synthetic feature maps: nchw = what the reference's hooks hand over (default), "
                         "nhwc = torch.channels_last; the other layout is timed beside it (roofline.other_layout)")
    ap.add_argument("--fit-n", type=int, default=None, help="vectors of the k-means fit (0: skip the sub-measurement; default 4 M = "
                    "BASELINE config 3, the same on every N)")
    ap.add_argument("--fit-variant", default="realistic", choices=list(FIT_VARIANTS),
                    help="--workload fit / --impl reference --workload fit: which C3 set (SURVEY.md section 8d)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C1 / C4 / C5 side measurements of the default line")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the end-to-end leg (default: min(steps, 20))")
    ap.add_argument("--quick", action="store_true", help="kernel timing only: skip the e2e, fit and cpu_baseline legs (tuning sweeps)")
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
