"""Where the host time of the class-surface call goes when the inputs already sit on the device (bench `e2e.device_inputs`):
cProfile over 20 calls of ood_utils.compute_ood_decisions_fused on the C2 batch.  Usage: python scripts/profile_class_api.py"""
import cProfile, io, logging, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ood_in_object_detection_b200 import ood_utils, ops, synth
from ood_in_object_detection_b200.results import Results, batch_shape

dev = torch.device("cuda", 0)
wl = synth.CONFIGS["C2"]
maps = bench.device_maps(wl, 1000, dev)
det = synth.detections(2000, wl.batch, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)
clusters, thr, table, lthr = bench.fit_tables(ops, wl, maps, 3000, dev)
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
shape = batch_shape(wl.batch, wl.img, wl.img)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
res_f, res_l = [], []
for i in range(wl.batch):
    b6 = np.concatenate([det["boxes"][i], det["conf"][i][:, None], det["cls"][i][:, None]], 1).astype(np.float32)
    res_f.append(Results(orig_img=shape, boxes=t(b6), extra_item=([m[i] for m in maps], t(det["strides"][i]))))
    res_l.append(Results(orig_img=shape, boxes=t(b6), extra_item=t(det["logits"][i])))
log = logging.getLogger("p"); log.setLevel(logging.ERROR)
for _ in range(3):
    ood_utils.compute_ood_decisions_fused(methods, res_f, log, logits_results=res_l)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    ood_utils.compute_ood_decisions_fused(methods, res_f, log, logits_results=res_l)
torch.cuda.synchronize()
print("ms per call", 1e3 * (time.perf_counter() - t0) / 20)
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    ood_utils.compute_ood_decisions_fused(methods, res_f, log, logits_results=res_l)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
