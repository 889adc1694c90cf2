"""Timing of K7 (per-cluster pair-distance sums) and of the searched 'KMeans' method on one (class, stride)-sized segment,
next to sklearn's silhouette_score on the host cores for a bounded sample (gpurun_out/ksearch_time.log)."""
import logging, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_in_object_detection_b200 import cluster_utils, ops, synth

dev = ops.default_device()
log = logging.getLogger("k"); log.setLevel(logging.CRITICAL)
n, dim, k = int(os.environ.get("KS_N", 20000)), 576, 8
x, lab = synth.blob_vectors(5, n, dim, k, 6.0)
xd, ld = torch.from_numpy(x).to(dev), torch.from_numpy(lab.astype(np.int32)).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for metric in ("l1", "l2", "cosine"):
    xs = ops.normalize_rows(xd) if metric == "cosine" else xd
    ops.pair_cluster_sums(xs, ld, k, metric); torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        ops.pair_cluster_sums(xs, ld, k, metric)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ops_per_pair = 1 if metric == "cosine" else 2
    print(f"pair_cluster_sums {metric:6s} n={n} D={dim}: {ms:8.2f} ms  {n * n * dim * ops_per_pair / ms / 1e9:7.2f} T instr-ops/s "
          f"({n * n / ms / 1e6:.1f} G pairs/s)")
t0 = time.perf_counter(); s = ops.silhouette_score(xd, ld, "l2"); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"ops.silhouette_score l2: {s:.6f} in {(t1 - t0) * 1e3:.1f} ms")
t0 = time.perf_counter()
labels, scores, ks = cluster_utils.search_number_of_clusters(xd, "l2", "silhouette", log)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"search_number_of_clusters (k = 2..14, l2, silhouette): {(t1 - t0) * 1e3:.1f} ms, best k = {ks[int(np.argmax(scores))]}")
from sklearn.metrics import silhouette_score
m = min(n, 6000)
t0 = time.perf_counter(); sc = silhouette_score(x[:m], lab[:m], metric="l2"); t1 = time.perf_counter()
print(f"sklearn silhouette_score l2 on {m} rows, {os.cpu_count()} host threads: {(t1 - t0) * 1e3:.1f} ms "
      f"({m * m / (t1 - t0) / 1e9:.3f} G pairs/s); device on the same rows: ", end="")
t0 = time.perf_counter(); sd = ops.silhouette_score(xd[:m].contiguous(), ld[:m].contiguous(), "l2"); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"{(t1 - t0) * 1e3:.1f} ms, |diff| = {abs(sc - sd):.2e}")
