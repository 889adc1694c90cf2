"""Producer-side hand-off at the C2 batch shape (64 images, 20 classes, 640 px: 8400 anchors): this package's postprocess against
the reference's `DetectionPredictor.postprocess` on the same CUDA tensors (reference code from oracle/_ref or /root/reference).
Usage: python scripts/time_postprocess.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import nms_inputs, fake_predictor
from ood_in_object_detection_b200 import synth
from ood_in_object_detection_b200.postprocess import postprocess

dev = torch.device("cuda", 0)
B = 64
pred, logits, _ = nms_inputs(seed=5, bs=B, nc=20, img=640)
p = torch.from_numpy(pred).to(dev)
raw = torch.cat([p[:, :4], torch.from_numpy(logits).to(dev)], 1)
maps = [torch.from_numpy(m).to(dev) for m in synth.feature_maps(3, B, (128, 256, 512), (80, 40, 20))]
img = torch.zeros((B, 3, 640, 640), device=dev)


def fp(mode, before):
    f = fake_predictor(mode, before, 0.25, device="cuda:0")
    f.batch = [[f"im{i}.jpg" for i in range(B)]]
    return f


def timed(fn, mode, before, reps):
    head = raw if before else p
    extra = None if mode == "logits" else maps
    for _ in range(2):
        res = fn(fp(mode, before), ((head.clone(),), extra), img, img)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        res = fn(fp(mode, before), ((head.clone(),), extra), img, img)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps, sum(len(r.boxes) for r in res)


out = {"batch": B, "anchors": int(p.shape[2])}
ref = None
try:
    from oracle import ref_shim
    if ref_shim.available():
        ref_shim.load()
        from ultralytics.models.yolo.detect.predict import DetectionPredictor
        ref = DetectionPredictor.postprocess
except Exception as e:                                                     # the timing of this package's side still stands
    out["reference_unavailable"] = repr(e)
for mode, before in (("ftmaps_and_strides", False), ("logits", True), ("roi_aligned_ftmaps", False)):
    ms, n = timed(postprocess, mode, before, 10)
    out[mode] = {"ms": round(ms, 3), "detections": n}
    if ref is not None:
        rms, rn = timed(ref, mode, before, 3)
        out[mode].update({"reference_ms": round(rms, 3), "reference_detections": rn})
print(json.dumps(out))
