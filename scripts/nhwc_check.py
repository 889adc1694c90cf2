import sys, numpy as np, torch
sys.path.insert(0, '.')
from ood_in_object_detection_b200 import ops, synth
wl = synth.CONFIGS["C5"]
maps = synth.feature_maps(3, 2, wl.channels, wl.map_hw)
det = synth.detections(4, 2, wl.img, wl.nc, 150)
a = ops.roi_pool(ops.make_batch([torch.from_numpy(m).cuda() for m in maps], det["boxes"], det["strides"], det["cls"], wl.img)).cpu().numpy()
b = ops.roi_pool(ops.make_batch([torch.from_numpy(m).cuda().contiguous(memory_format=torch.channels_last) for m in maps], det["boxes"], det["strides"], det["cls"], wl.img)).cpu().numpy()
print("max rel diff nchw vs nhwc pooled:", float(np.max(np.abs(a - b) / (np.abs(a).max(1, keepdims=True) + 1e-30))))
