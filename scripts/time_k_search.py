"""Time the silhouette-searched KMeans of one segment (cluster_utils.search_number_of_clusters) and its K7 part.
Usage: python scripts/time_k_search.py [n_rows] [dim]; OODB200_PAIR_MATRIX_GB=0 selects the recompute-per-k route."""
import json, logging, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ood_in_object_detection_b200 import cluster_utils, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 576
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)
centres = torch.randn((6, dim), device=dev, generator=g) * 2
x = (centres[torch.randint(0, 6, (n,), device=dev, generator=g)] + torch.randn((n, dim), device=dev, generator=g)).abs().contiguous()
log = logging.getLogger("ks"); log.setLevel(logging.ERROR)
out = {"n": n, "dim": dim, "matrix_gb_budget": os.environ.get("OODB200_PAIR_MATRIX_GB", "8")}
for metric in ("l2", "cosine", "l1"):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        labels, scores, ks = cluster_utils.search_number_of_clusters(x, metric, "silhouette", log)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    lab = labels.to(torch.int32)
    kc = int(lab.max()) + 1
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); pairs = ops.PairDistances(x, metric); e[1].record()
    s1 = pairs.cluster_sums(lab, kc); e[2].record()
    s0 = ops.pair_cluster_sums(pairs.xs, lab, kc, metric); e[3].record()
    torch.cuda.synchronize()
    out[metric] = {"search_s": round(dt, 4), "best_k": int(ks[int(np.argmax(scores))]), "scores": [round(float(v), 6) for v in scores],
                   "matrix_ms": round(e[0].elapsed_time(e[1]), 3), "fold_from_matrix_ms": round(e[1].elapsed_time(e[2]), 3),
                   "recompute_fold_ms": round(e[2].elapsed_time(e[3]), 3),
                   "max_rel_diff_routes": float(((s1 - s0).abs() / s0.abs().clamp_min(1e-30)).max())}
    del pairs
print(json.dumps(out))
