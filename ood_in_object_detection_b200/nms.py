"""NMS with the OoD payload on the GPU: drop-in for the default path of the reference's
`ultralytics.utils.ops.non_max_suppression_old` (/root/reference/ultralytics/utils/ops.py:348-530), the step immediately
before the scoring hot path.

The reference loops over the images in python (~10 small launches + torchvision.nms each) and gathers the payload -- the
per-anchor `extra_item` rows (raw class logits) and `strides` -- with the same index tensors.  Here the whole batch is ONE
launch (csrc/nms.cu, a CTA per image); the result lists have the reference's structure and hold device tensors.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .ops import _ptr, _stream


def nms_padded(prediction: torch.Tensor, conf_thres: float, iou_thres: float, max_det: int = 300, nc: int = 0,
               max_nms: int = 30000, max_wh: int = 7680, extra_item: Optional[torch.Tensor] = None,
               strides: Optional[torch.Tensor] = None):
    """The kernel's own result layout: det [bs, max_det, 6], extra [bs, max_det, E] | None, strides [bs, max_det] | None,
    anchor int32 [bs, max_det] (index of every kept detection's anchor), counts (python list, one host read).  Rows beyond an
    image's count are zero.  postprocess.py works on these padded tensors (batched box clipping) before slicing."""
    if not prediction.is_cuda:
        raise RuntimeError("ood_in_object_detection_b200.nms needs CUDA tensors: there is no CPU fallback")
    lib = _lib.load()
    pred = prediction.to(torch.float32).contiguous()
    bs, rows, A = (int(v) for v in pred.shape)
    nc = nc or rows - 4
    if rows - nc - 4 != 0:
        raise NotImplementedError("mask coefficients after the class columns are not supported")
    dev = pred.device
    ex = extra_item.to(device=dev, dtype=torch.float32).contiguous() if extra_item is not None else None
    st = strides.to(device=dev, dtype=torch.float32).contiguous() if strides is not None else None
    ne = int(ex.shape[1]) if ex is not None else 0
    det = torch.zeros((bs, max_det, 6), dtype=torch.float32, device=dev)
    out_ex = torch.zeros((bs, max_det, max(ne, 1)), dtype=torch.float32, device=dev) if ex is not None else None
    out_st = torch.zeros((bs, max_det), dtype=torch.float32, device=dev) if st is not None else None
    anchor = torch.zeros((bs, max_det), dtype=torch.int32, device=dev)
    count = torch.zeros(bs, dtype=torch.int32, device=dev)
    ws = torch.empty(int(lib.oodb200_nms_workspace_bytes(bs, A)), dtype=torch.uint8, device=dev)
    _lib.check(lib.oodb200_nms_payload_f32(_ptr(pred), _ptr(ex), _ptr(st), bs, nc, ne, A, float(conf_thres), float(iou_thres),
                                           float(max_wh), int(max_det), int(max_nms), _ptr(det), _ptr(out_ex), _ptr(out_st),
                                           _ptr(anchor), _ptr(count), _ptr(ws), int(ws.numel()), _stream()),
               "oodb200_nms_payload_f32")
    return det, (out_ex[:, :, :ne] if out_ex is not None else None), out_st, anchor, count.cpu().tolist()


def slice_results(det, out_ex, out_st, counts):
    """Padded tensors -> the reference's list structure (views)."""
    dev = det.device
    ret = [[det[i, :k] for i, k in enumerate(counts)]]
    if out_ex is not None:
        ret.append([out_ex[i, :k] if k else torch.empty(0, device=dev) for i, k in enumerate(counts)])
    if out_st is not None:
        ret.append([out_st[i, :k] if k else torch.empty(0, device=dev) for i, k in enumerate(counts)])
    return ret[0] if len(ret) == 1 else tuple(ret)


def non_max_suppression(prediction, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None, agnostic: bool = False,
                        multi_label: bool = False, labels=(), max_det: int = 300, nc: int = 0, max_time_img: float = 0.05,
                        max_nms: int = 30000, max_wh: int = 7680, extra_item: Optional[torch.Tensor] = None,
                        strides: Optional[torch.Tensor] = None, v10: bool = False):
    """prediction [bs, 4 + nc, A] (cx, cy, w, h, class confidences) CUDA float32; extra_item [bs, E, A]; strides [A].
    Returns like the reference: `output` (list of [k, 6] tensors: xyxy, confidence, class), followed by the list of [k, E]
    payload rows when `extra_item` is given and the list of [k] strides when `strides` is given.
    `agnostic=True` is the reference's class offset of 0 (ops.py:487: all classes suppress each other).  Options outside the
    path the OoD pipeline uses raise NotImplementedError: multi_label, apriori labels, v10, mask columns, and the class filter
    (the reference's own filter, ops.py:470-474, leaves `extra_item` unfiltered and indexes `strides` with a mask of the wrong
    length: there is no payload behaviour to reproduce).  `max_time_img` has no meaning here (the reference abandons the
    remaining images on timeout)."""
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if multi_label or classes or (labels and len(labels)) or v10:
        raise NotImplementedError("only the best-class path of non_max_suppression_old (no class filter, labels, multi_label, v10) runs on the GPU")
    det, out_ex, out_st, _, counts = nms_padded(prediction, conf_thres, iou_thres, max_det=max_det, nc=nc, max_nms=max_nms,
                                                max_wh=0 if agnostic else max_wh, extra_item=extra_item, strides=strides)
    return slice_results(det, out_ex, out_st, counts)
