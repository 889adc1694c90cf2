"""Producer side of the hand-off: from the detector head's raw output to the `Results` the OoD
methods consume -- a drop-in for the OoD branch of the reference's `DetectionPredictor.postprocess`
(/root/reference/ultralytics/models/yolo/detect/predict.py:117-363).

The reference, per batch: `non_max_suppression_old` (a python loop over the images, ~10 launches + torchvision.nms each), a
device->host copy of the whole uint8 image batch (`convert_torch2numpy_batch`) whose only later use is its shape, per-image
`scale_boxes` (8 small launches each) and per-image lists of feature maps.  Here: ONE NMS launch for the batch (csrc/nms.cu)
with the payload gathered by the same kernel, box clipping as 4 batched ops on the padded result, and `Results` that hold
VIEWS of the batched maps (no copy of a map or an image is made; `orig_img` is a shape-only placeholder unless
`keep_images=True`).

Binding in the reference: `DetectionPredictor.postprocess = postprocess` -- the function reads the
same attributes of the predictor (`args.conf / iou / max_det / agnostic_nms / classes / model`, `model.names`,
`model.model.extraction_mode`, `model.model.model[-1].output_values_before_sigmoid`, `batch[0]`).
Outside the path: image lists as `orig_imgs` (the OoD pipeline feeds tensors; the reference rescales to `shape[1:3]` of an
HWC image there) and the `classes` filter (raises, like nms.non_max_suppression; `agnostic_nms` is served).
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import nms as _nms
from .results import Results, batch_shape


def _stride_of_anchor(input_size: int, device) -> torch.Tensor:
    """predict.py:162-171 / :250-259: 0 / 1 / 2 for the anchors of the stride-8 / 16 / 32 grids, in head order."""
    n = [(input_size // s) ** 2 for s in (8, 16, 32)]
    return torch.cat([torch.full((n[i],), float(i), device=device) for i in range(3)])


def _xyxy2xywh_rows(pred: torch.Tensor) -> torch.Tensor:
    """v10 heads emit xyxy: the reference converts rows 0..3 to (cx, cy, w, h) before NMS (predict.py:262-270)."""
    b = pred[:, :4]
    out = torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], dim=1)
    return torch.cat([out, pred[:, 4:]], dim=1)


def scale_and_clip_(det: torch.Tensor, net_hw: Sequence[int], img0_hw: Sequence[int]) -> torch.Tensor:
    """`ops.scale_boxes(img1_shape, boxes, img0_shape)` (ultralytics/utils/ops.py:96-130) + `clip_boxes` (:536-555) on the
    padded [B, max_det, 6] result at once, in place: same float32 operations per box as the reference's per-image calls."""
    gain = min(net_hw[0] / img0_hw[0], net_hw[1] / img0_hw[1])
    pad = (round((net_hw[1] - img0_hw[1] * gain) / 2 - 0.1), round((net_hw[0] - img0_hw[0] * gain) / 2 - 0.1))
    det[..., 0] -= pad[0]
    det[..., 1] -= pad[1]
    det[..., 2] -= pad[0]
    det[..., 3] -= pad[1]
    det[..., :4] /= gain
    det[..., 0].clamp_(0, img0_hw[1])
    det[..., 1].clamp_(0, img0_hw[0])
    det[..., 2].clamp_(0, img0_hw[1])
    det[..., 3].clamp_(0, img0_hw[0])
    return det


def postprocess(self, preds, img, orig_imgs, keep_images: bool = False, **kwargs) -> List[Results]:
    """`DetectionPredictor.postprocess(self, preds, img, orig_imgs)` for a model whose `extraction_mode` is set
    (predict.py:139-141).  preds = ((prediction [B, 4 + nc, A], ...), extra) as the patched head returns it."""
    inner = self.model.model
    if not hasattr(inner, "extraction_mode"):
        raise NotImplementedError("postprocess: only the OoD extraction branch (model.extraction_mode set) is served here")
    mode = inner.extraction_mode
    before_sigmoid = bool(inner.model[-1].output_values_before_sigmoid)
    if isinstance(orig_imgs, (list, tuple)):
        raise NotImplementedError("postprocess: image lists are outside the OoD path (the pipeline feeds [B, 3, H, W] tensors)")
    args = self.args
    if args.classes:
        raise NotImplementedError("postprocess: class-filtered NMS is outside the OoD path (see nms.non_max_suppression)")
    # predict.py:143-148
    output_extra = preds[0][0] if (before_sigmoid or mode == "logits") else preds[1]
    pred = preds[0][0]
    if not pred.is_cuda:
        raise RuntimeError("postprocess needs CUDA tensors: there is no CPU fallback")
    is_v10 = "v10" in str(getattr(args, "model", ""))
    B = int(pred.shape[0])
    dev = pred.device
    net_hw = (int(img.shape[2]), int(img.shape[3]))
    img0_hw = (int(orig_imgs.shape[2]), int(orig_imgs.shape[3]))
    maps_per_image = lambda: [[m[i] for m in output_extra] for i in range(B)]          # views of the batched maps, no copy
    strides, payload_in = None, None
    if mode in ("roi_aligned_ftmaps", "ftmaps_and_strides", "ftmaps_and_strides_exact_pos"):
        if mode != "ftmaps_and_strides_exact_pos":
            strides = _stride_of_anchor(net_hw[0], dev)
        if is_v10 and mode == "ftmaps_and_strides":
            pred = _xyxy2xywh_rows(pred)
    elif mode == "logits":                                                              # predict.py:195-220
        payload_in = pred
        if before_sigmoid:
            boxes_only = _xyxy2xywh_rows(pred)[:, :4] if is_v10 else pred[:, :4]
            pred = torch.cat((boxes_only, pred[:, 4:].sigmoid()), dim=1)
    elif mode != "all_ftmaps":
        raise ValueError(f"postprocess: unknown extraction mode {mode!r}")
    det, out_ex, out_st, anchor, counts = _nms.nms_padded(pred, args.conf, args.iou, max_det=args.max_det, extra_item=payload_in,
                                                          strides=strides, max_wh=0 if args.agnostic_nms else 7680)
    if mode == "roi_aligned_ftmaps":                                                    # predict.py:184-193: boxes in network pixels
        from .ood_utils import extract_roi_aligned_features_from_correct_stride
        extra = extract_roi_aligned_features_from_correct_stride(
            ftmaps=output_extra, boxes=[det[i, :k, :4] for i, k in enumerate(counts)], strides=[out_st[i, :k] for i, k in enumerate(counts)],
            img_shape=net_hw, device=dev)
    elif mode == "ftmaps_and_strides":
        extra = list(zip(maps_per_image(), _nms.slice_results(det, None, out_st, counts)[1]))
    elif mode == "ftmaps_and_strides_exact_pos":                                        # predict.py:300-330: the anchor index itself
        a64 = anchor.to(torch.int64)
        extra = list(zip(maps_per_image(), [a64[i, :k] if k else torch.empty(0, device=dev) for i, k in enumerate(counts)]))
    elif mode == "logits":
        extra = [out_ex[i, :k, 4:] if k else torch.empty(0, device=dev) for i, k in enumerate(counts)]
    else:
        extra = maps_per_image()
    # predict.py:342-360: tensor source -> the whole batch stands for every image's `orig_img`; boxes clipped to its (H, W)
    scale_and_clip_(det, net_hw, img0_hw)                                               # every image's rows at once
    if keep_images:
        orig = (orig_imgs.permute(0, 2, 3, 1).contiguous() * 255).clamp(0, 255).to(torch.uint8).cpu().numpy()
    else:
        orig = batch_shape(B, img0_hw[0], img0_hw[1])
    paths = self.batch[0] if getattr(self, "batch", None) else [""] * B
    names = getattr(self.model, "names", None)
    return [Results(orig_img=orig, path=(paths[i] if isinstance(paths, list) else paths), names=names, boxes=det[i, :k],
                    extra_item=extra[i]) for i, k in enumerate(counts)]
