"""Where the end-to-end step (host inputs -> decisions on the host) spends its time: torch.profiler table of one
`compute_ood_decisions_fused` call on the C2 batch + wall-clock split (gpurun_out/e2e_profile.log)."""
import logging, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ood_in_object_detection_b200 import ood_utils, ops, synth
from ood_in_object_detection_b200.results import Results, batch_shape

dev = torch.device("cuda", 0)
wl = synth.CONFIGS["C2"]
maps = bench.device_maps(wl, 1000, dev)
det = synth.detections(2000, wl.batch, wl.img, wl.nc, wl.lam, fixed=wl.fixed_boxes)
clusters, thr, table, lthr = bench.fit_tables(ops, wl, maps, 11, dev)
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
log = logging.getLogger("p"); log.setLevel(logging.ERROR)
step = lambda: ood_utils.compute_ood_decisions_fused(methods, res_f, log, logits_results=res_l)
for _ in range(3):
    step()
torch.cuda.synchronize()
# raw copies of the three maps
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d = [torch.empty_like(m) for m in maps]
for rep in range(2):
    e0.record()
    for a, b in zip(d, h_maps):
        a.copy_(b, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
print("3 map copies alone: %.2f ms for %.1f MB" % (e0.elapsed_time(e1), sum(m.numel() * 4 for m in maps) / 1e6))
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("e2e step wall ms:", [round(t * 1e3, 2) for t in ts])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=60))
