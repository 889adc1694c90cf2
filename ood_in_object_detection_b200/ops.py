"""Thin torch-tensor wrappers over the C ABI (include/oodb200.h).

torch is used for device memory and streams only; every computation below happens in the
hand-written CUDA kernels of liboodb200.so.  Nothing here has a CPU path: tensors that arrive
on the host are copied to the device, the kernels run there.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib

METRIC_SLOT = {"l1": 0, "manhattan": 0, "cityblock": 0, "l2": 1, "euclidean": 1, "cosine": 2}
LOGIT_SLOT = {"MSP": 0, "Energy": 1, "ODIN": 2, "Sigmoid": 3, "MaxLogit": 4}
FUSE_SLOT = {"and": 0, "or": 1, "majority_voting": 2}
N_LOGIT = 5
LOGIT_FLAG_POST_SIGMOID = 0x100      # OR-ed into method_mask: Sigmoid slot = input value (inputs are post-sigmoid)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_empty_stub = {}


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    """Device address for the C ABI.  An EMPTY tensor (a rank that owns no row of a sharded fit) has a null data pointer,
    which the entry points reject: it is passed as the address of a small stub buffer that is never dereferenced."""
    if t is None:
        return C.c_void_p(0)
    if t.numel() == 0 and t.is_cuda:
        stub = _empty_stub.get(t.device)
        if stub is None:
            stub = _empty_stub[t.device] = torch.zeros(256, dtype=torch.uint8, device=t.device)
        return C.c_void_p(stub.data_ptr())
    return C.c_void_p(t.data_ptr())


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("ood_in_object_detection_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class _PinnedRing:
    """Reusable pinned staging buffers, one ring per size class.  torch's caching pinned allocator hands a block out again
    only after the copy that read it has completed; a caller that queues several batches without synchronising would get
    a FRESH cudaHostAlloc (~0.5 ms) for every small upload.  A slot is reused after its own copy event has completed
    (waited on only if the ring has wrapped around while that copy is still in flight)."""
    SLOTS = 32

    def __init__(self):
        self.rings = {}

    def stage(self, t: torch.Tensor, device) -> torch.Tensor:
        nbytes = max(int(t.numel()) * t.element_size(), 1)
        size = 1 << max(nbytes - 1, 255).bit_length()
        ring = self.rings.get(size)
        if ring is None:
            ring = self.rings[size] = {"next": 0, "slots": []}
        if len(ring["slots"]) < self.SLOTS:
            ring["slots"].append([torch.empty(size, dtype=torch.uint8, pin_memory=True), None])
            slot = ring["slots"][-1]
        else:
            slot = ring["slots"][ring["next"]]
            ring["next"] = (ring["next"] + 1) % self.SLOTS
            if slot[1] is not None:
                slot[1].synchronize()
        stage = slot[0][:nbytes].view(t.dtype).reshape(t.shape)
        stage.copy_(t)
        out = torch.empty(t.shape, dtype=t.dtype, device=device)
        out.copy_(stage, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        slot[1] = ev
        return out


_pinned = _PinnedRing()
_STAGE_MAX_BYTES = 8 << 20          # larger host tensors are pinned (or not) by their owner


def h2d(t, device, dtype=None) -> torch.Tensor:
    """Host array / tensor -> device through PINNED staging.  A copy from pageable memory makes the host wait for
    everything queued on the stream before it (the 100s of MB of feature maps of the same batch): staged through a ring
    of reusable pinned buffers the call returns at once and the host keeps preparing the next launch under that copy."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.is_cuda:
        return t.to(device).contiguous()
    if t.is_pinned() and t.is_contiguous():
        return t.to(device, non_blocking=True)
    if t.numel() * t.element_size() <= _STAGE_MAX_BYTES and t.numel() > 0:
        return _pinned.stage(t.contiguous(), device)
    return t.contiguous().to(device, non_blocking=True)


def _dev(t, device, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    if t.device != device:
        return h2d(t, device, dtype)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


@dataclass
class DetectionBatch:
    """Flat device-side view of a batch of detections and their hooked feature maps
    (what `Results.boxes` / `Results.extra_item` hold per image in the reference, T1 in SURVEY.md §8a)."""
    map_ptrs: torch.Tensor          # int64 [n_img*3] device pointers to the CHW maps
    map_chw: np.ndarray             # host int32 [9]
    scale: np.ndarray               # host float32 [3]
    n_img: int
    boxes: torch.Tensor             # [n,4] f32
    img_idx: torch.Tensor           # [n] i32
    stride_idx: torch.Tensor        # [n] i32
    cls: torch.Tensor               # [n] i32
    img_start: torch.Tensor         # [n_img+1] i32
    counts: List[int]               # host copy of boxes per image
    keepalive: tuple = ()           # tensors whose storage map_ptrs points into
    nhwc: bool = False              # the maps are channels-last ([H, W, C] per image): the *_nhwc entry points read them

    @property
    def n(self) -> int:
        return int(self.boxes.shape[0])


@dataclass
class StagedMaps:
    """Feature maps of a batch on the device (or on their way: the copies are queued on the current stream)."""
    ptrs: np.ndarray           # int64 [n_img, 3] device addresses of the CHW maps
    chw: list                  # 3 x (C_s, H_s, W_s)
    n_img: int
    keepalive: tuple
    nhwc: bool = False         # channels-last maps: element (c, y, x) at [(y * W + x) * C + c]


def _is_channels_last(t: torch.Tensor) -> bool:
    """4-D [B,C,H,W] or 3-D [C,H,W] tensor whose memory is [.., H, W, C] (torch.channels_last and its per-image slices);
    a map with C == 1 or H == W == 1 is both layouts at once and counts as the default one."""
    if t.dim() == 4:
        c, h, w = t.shape[1:]
        st = t.stride()[1:]
    else:
        c, h, w = t.shape
        st = t.stride()
    if c == 1 or h * w == 1:
        return False
    return tuple(st) == (1, w * c, c)


def stage_maps(maps, n_img: int, device=None) -> StagedMaps:
    """Queue the upload of the feature maps (the bulk of a host batch) BEFORE any other host-side preparation, so that
    the detections are flattened and the launches prepared while the DMA engine works.
    maps: either 3 batched tensors [B,C_s,H_s,W_s] or a per-image list of 3 CHW tensors."""
    device = device or default_device()
    keep = []
    nhwc = False
    if len(maps) == 3 and all(isinstance(m, torch.Tensor) and m.dim() == 4 for m in maps):
        nhwc = all(_is_channels_last(m) for m in maps)
        if nhwc:                                       # keep the channels-last memory: no re-layout, the kernels read it as is
            mt = [m if (m.device == device and m.dtype == torch.float32) else
                  m.to(device=device, dtype=torch.float32, non_blocking=True, memory_format=torch.preserve_format) for m in maps]
            assert all(_is_channels_last(m) for m in mt)
        else:
            mt = [_dev(m, device, torch.float32) for m in maps]
        assert all(m.shape[0] == n_img for m in mt), "batched maps must have one slice per image"
        chw = [tuple(m.shape[1:]) for m in mt]
        ptrs = np.empty((n_img, 3), dtype=np.int64)
        for s, m in enumerate(mt):
            ptrs[:, s] = m.data_ptr() + np.arange(n_img, dtype=np.int64) * (m.stride(0) * 4)
        keep = mt
    else:
        assert len(maps) == n_img, "maps must be 3 batched tensors or one list of 3 CHW tensors per image"
        ptrs = np.empty((n_img, 3), dtype=np.int64)
        chw = None
        for per_img in maps:
            assert len(per_img) == 3 and all(isinstance(t, torch.Tensor) and t.dim() == 3 for t in per_img), \
                "each image needs 3 CHW maps"
            shp = [tuple(t.shape) for t in per_img]
            assert chw is None or shp == chw, "all images must share the map shapes"
            chw = shp
        if chw is None:
            chw = [(1, 1, 1)] * 3
        flat_maps = [t for per_img in maps for t in per_img]
        nhwc = bool(flat_maps) and all(_is_channels_last(t) for t in flat_maps)
        for s in range(3):
            col = [per_img[s] for per_img in maps]
            nbytes = int(np.prod(chw[s])) * 4
            if nhwc:                                   # per-image channels-last views: moved as they are, one by one
                ts = [t if (t.device == device and t.dtype == torch.float32) else
                      t.to(device=device, dtype=torch.float32, non_blocking=True, memory_format=torch.preserve_format) for t in col]
                assert all(_is_channels_last(t) for t in ts), "channels-last maps must keep their layout on the device"
                ptrs[:, s] = [t.data_ptr() for t in ts]
                keep.extend(ts)
                continue
            batched = None
            st0 = col[0].untyped_storage() if col else None
            if col and not col[0].is_cuda and col[0].dtype == torch.float32 and all(
                    t.is_contiguous() and t.dtype == torch.float32 and t.data_ptr() == col[0].data_ptr() + i * nbytes
                    and t.untyped_storage().data_ptr() == st0.data_ptr() for i, t in enumerate(col)) \
                    and col[0].data_ptr() - st0.data_ptr() + len(col) * nbytes <= st0.nbytes():
                # host maps that are consecutive views of one batched tensor (what the detector hook hands out): ONE copy
                batched = torch.as_strided(col[0], (len(col),) + chw[s], (int(np.prod(chw[s])),) + tuple(col[0].stride()))
                dev_b = batched.to(device, non_blocking=True)
                ptrs[:, s] = dev_b.data_ptr() + np.arange(n_img, dtype=np.int64) * nbytes
                keep.append(dev_b)
            else:
                ts = [_dev(t, device, torch.float32) for t in col]
                ptrs[:, s] = [t.data_ptr() for t in ts]
                keep.extend(ts)
    return StagedMaps(ptrs=ptrs, chw=[tuple(int(v) for v in c) for c in chw], n_img=n_img, keepalive=tuple(keep), nhwc=nhwc)


def make_batch(maps, boxes: Sequence, strides: Sequence, cls: Sequence, img_w: int, device=None) -> DetectionBatch:
    """maps: either 3 batched tensors [B,C_s,H_s,W_s], a per-image list of 3 CHW tensors, or `stage_maps(...)` of those.
    boxes/strides/cls: per-image sequences ([M_i,4], [M_i], [M_i]); strides in {0,1,2}.
    img_w: width of the network input (`res.orig_img.shape[2]`, predict.py:68)."""
    device = device or default_device()
    n_img = len(boxes)
    staged = maps if isinstance(maps, StagedMaps) else stage_maps(maps, n_img, device)
    assert staged.n_img == n_img, "one set of maps per image"
    ptrs, chw, keep = staged.ptrs, staged.chw, staged.keepalive
    counts = [int(len(b)) for b in boxes]
    n = sum(counts)
    def _flat(seq, dtype, width=None):
        """per-image sequences -> one device tensor; host inputs are concatenated on the host and copied ONCE"""
        ts = [t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t)) for t in seq]
        shape = (-1, width) if width else (-1,)
        if all(not t.is_cuda for t in ts):
            return h2d(torch.cat([t.reshape(shape) for t in ts]), device, dtype)
        return torch.cat([t.to(device).reshape(shape) for t in ts]).to(dtype).contiguous()

    if n:
        bx = _flat(boxes, torch.float32, 4)
        st = _flat(strides, torch.int32)
        cl = _flat(cls, torch.int32)
    else:
        bx = torch.zeros((0, 4), dtype=torch.float32, device=device)
        st = torch.zeros(0, dtype=torch.int32, device=device)
        cl = torch.zeros(0, dtype=torch.int32, device=device)
    start = np.zeros(n_img + 1, dtype=np.int32)
    np.cumsum(counts, out=start[1:])
    img_idx = np.repeat(np.arange(n_img, dtype=np.int32), counts)
    return DetectionBatch(
        map_ptrs=h2d(ptrs.reshape(-1), device),
        map_chw=np.asarray(chw, dtype=np.int32).reshape(9),
        scale=np.asarray([np.float32(c[2] / img_w) for c in chw], dtype=np.float32),
        n_img=n_img, boxes=bx, img_idx=h2d(img_idx, device),
        stride_idx=st, cls=cl, img_start=h2d(start, device),
        counts=counts, keepalive=tuple(keep), nhwc=staged.nhwc)


@dataclass
class CentroidTable:
    """Device-side packing of `clusters[cls][stride]` / `thresholds[cls][stride]` (SURVEY.md §5 format contract)."""
    cent: torch.Tensor
    cent_unit: torch.Tensor
    cent_off: torch.Tensor     # int64 [3*nc]
    cent_k: torch.Tensor       # int32 [3*nc]
    thr: torch.Tensor          # float64 [3 metrics, 3*nc]
    nc: int
    k_host: np.ndarray         # [3, nc]


def _unit_rows(a: np.ndarray) -> np.ndarray:
    """sklearn.preprocessing.normalize on float32 rows (what cosine_distances applies to the centroids)."""
    a = np.array(a, dtype=np.float32, copy=True)
    nrm = np.sqrt(np.einsum("ij,ij->i", a, a))
    nrm[nrm < 10 * np.finfo(np.float32).eps] = 1.0
    a /= nrm[:, None]
    return a


def pack_thresholds(thresholds_by_metric: dict, nc: int) -> np.ndarray:
    """{metric_slot: thresholds[cls][stride]} -> float64 [3 metrics, 3*nc] (index s*nc + c), NaN = "no threshold":
    python floats are taken, falsy entries ([] / 0 / 0.0 / None) mean OoD like in the reference (ood_utils.py:2173)."""
    thr = np.full((3, 3, nc), np.nan, dtype=np.float64)
    for slot, th in thresholds_by_metric.items():
        if th is None:
            continue
        for c in range(min(nc, len(th))):
            for s in range(3):
                v = th[c][s] if s < len(th[c]) else []
                if v is None or (isinstance(v, (list, tuple, np.ndarray)) and np.size(v) == 0):
                    continue
                if v:                                  # python truthiness, like the reference
                    thr[slot, s, c] = float(v)
    return thr.reshape(3, 3 * nc)


_table_cache: dict = {}


def _table_key(clusters, thresholds_by_metric: dict, dims, device):
    """Identity of a (clusters, thresholds) pair: buffer address, shape and two probe values of every centroid array, the
    threshold values themselves.  Cheap next to packing + uploading the tables for every batch."""
    ck = []
    for per_cls in clusters:
        for a in per_cls:
            if isinstance(a, np.ndarray) and a.size:
                ck.append((a.ctypes.data, a.shape, a.dtype.str, float(a.flat[0]), float(a.flat[-1])))
            else:
                ck.append(np.size(a))
    tk = []
    for slot in sorted(thresholds_by_metric):
        th = thresholds_by_metric[slot]
        tk.append((slot, None if th is None else tuple(tuple(tuple(np.ravel(np.asarray(v, dtype=np.float64)).tolist()) for v in per) for per in th)))
    return (tuple(ck), tuple(tk), tuple(int(d) for d in dims), str(device))


def pack_centroids_cached(clusters, thresholds_by_metric: dict, dims: Sequence[int], device=None) -> "CentroidTable":
    """`pack_centroids` memoised on the content fingerprint of its inputs (scoring calls it once per batch with the same
    fitted tables); holds the 4 most recent tables."""
    device = device or default_device()
    try:
        key = _table_key(clusters, thresholds_by_metric, dims, device)
    except (TypeError, ValueError):
        return pack_centroids(clusters, thresholds_by_metric, dims, device)
    hit = _table_cache.get(key)
    if hit is None:
        hit = pack_centroids(clusters, thresholds_by_metric, dims, device)
        if len(_table_cache) >= 4:
            _table_cache.pop(next(iter(_table_cache)))
        _table_cache[key] = hit
    return hit


def pack_centroids(clusters, thresholds_by_metric: dict, dims: Sequence[int], device=None) -> CentroidTable:
    """clusters[cls][stride] = ndarray [K, C_s] or empty; thresholds_by_metric = {metric_slot: thresholds[cls][stride]}
    with python floats or falsy entries ([] / 0 / 0.0 -> "no threshold", ood_utils.py:2173)."""
    device = device or default_device()
    nc = len(clusters)
    off = np.zeros((3, nc), dtype=np.int64)
    kk = np.zeros((3, nc), dtype=np.int32)
    chunks, units = [], []
    pos = 0
    for s in range(3):
        for c in range(nc):
            a = clusters[c][s] if s < len(clusters[c]) else []
            a = np.asarray(a, dtype=np.float32)
            off[s, c] = pos
            if a.size == 0:
                continue
            a = a.reshape(-1, a.shape[-1]) if a.ndim > 1 else a.reshape(1, -1)
            if a.shape[1] != dims[s]:
                raise ValueError(f"clusters[{c}][{s}] has {a.shape[1]} features, the stride-{s} map has {dims[s]} channels")
            kk[s, c] = a.shape[0]
            chunks.append(a.reshape(-1))
            units.append(_unit_rows(a).reshape(-1))
            pos += a.size
            pos = (pos + 3) & ~3           # keep every slice 16-byte aligned for 128-bit loads
            pad = pos - (int(off[s, c]) + a.size)
            if pad:
                chunks.append(np.zeros(pad, np.float32))
                units.append(np.zeros(pad, np.float32))
    flat = np.concatenate(chunks) if chunks else np.zeros(4, np.float32)
    flat_u = np.concatenate(units) if units else np.zeros(4, np.float32)
    thr = pack_thresholds(thresholds_by_metric, nc)
    t = lambda a: h2d(a, device)
    return CentroidTable(cent=t(flat), cent_unit=t(flat_u), cent_off=t(off.reshape(-1)), cent_k=t(kk.reshape(-1)),
                         thr=t(thr), nc=nc, k_host=kk)


def q1_plan(batch: DetectionBatch, cls_used: Optional[torch.Tensor] = None, out_index: Optional[torch.Tensor] = None):
    """(cls_used, out_index) of the reference's quirk Q1 (ood_utils.py:2152-2154), standalone."""
    lib = _lib.load()
    cls_used = torch.empty_like(batch.cls) if cls_used is None else cls_used
    out_index = torch.empty_like(batch.cls) if out_index is None else out_index
    _lib.check(lib.oodb200_q1_plan_i32(_ptr(batch.img_start), _ptr(batch.stride_idx), _ptr(batch.cls), batch.n_img,
                                       _ptr(cls_used), _ptr(out_index), _stream()), "oodb200_q1_plan_i32")
    return cls_used, out_index


_workspaces = {}


def _workspace(batch: DetectionBatch, nc: int) -> torch.Tensor:
    """Scratch for the pooling kernels, grown on demand: one per (device, stream) -- kernels on one stream reuse it in
    order, passes issued on different streams (sub-batches side by side) must not share work lists."""
    lib = _lib.load()
    need = int(lib.oodb200_fmap_workspace_bytes(batch.n, int(nc), batch.map_chw.ctypes.data_as(C.c_void_p)))
    dev = batch.boxes.device
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def roi_pool(batch: DetectionBatch, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1: [n, Cmax] pooled vectors (row i valid up to C of its stride)."""
    lib = _lib.load()
    cmax = int(batch.map_chw.reshape(3, 3)[:, 0].max())
    if out is None:
        out = torch.zeros((batch.n, cmax), dtype=torch.float32, device=batch.boxes.device)
    ws = _workspace(batch, 0)
    fn, name = (lib.oodb200_roi_pool_nhwc_f32, "oodb200_roi_pool_nhwc_f32") if batch.nhwc else (lib.oodb200_roi_pool_f32, "oodb200_roi_pool_f32")
    _lib.check(fn(
        _ptr(batch.map_ptrs), batch.map_chw.ctypes.data_as(C.c_void_p), batch.scale.ctypes.data_as(C.c_void_p),
        batch.n_img, _ptr(batch.boxes), _ptr(batch.img_idx), _ptr(batch.stride_idx), _ptr(batch.img_start), batch.n,
        _ptr(out), int(out.stride(0)), _ptr(ws), int(ws.numel()), _stream()), name)
    return out


@dataclass
class FmapScores:
    dist: torch.Tensor        # [3, n] f32 (slot = metric)
    argmin: torch.Tensor      # [3, n] i32
    decision: torch.Tensor    # [3, n] u8
    pooled: Optional[torch.Tensor] = None
    cls_used: Optional[torch.Tensor] = None     # [n] i32, per input box
    out_index: Optional[torch.Tensor] = None    # [n] i32, output slot of every input box


def alloc_fmap_scores(n: int, device, cmax: int = 0, want_pooled: bool = False, want_plan: bool = False) -> FmapScores:
    return FmapScores(dist=torch.empty((3, n), dtype=torch.float32, device=device),
                      argmin=torch.empty((3, n), dtype=torch.int32, device=device),
                      decision=torch.zeros((3, n), dtype=torch.uint8, device=device),
                      pooled=torch.zeros((n, cmax), dtype=torch.float32, device=device) if want_pooled else None,
                      cls_used=torch.empty(n, dtype=torch.int32, device=device) if want_plan else None,
                      out_index=torch.empty(n, dtype=torch.int32, device=device) if want_plan else None)


def fmap_score(batch: DetectionBatch, table: CentroidTable, metric_mask: int, normalize: bool = True,
               compat_q1: bool = True, want_pooled: bool = False, want_plan: bool = False,
               out: Optional[FmapScores] = None) -> FmapScores:
    """K1+K2 fused pass over every box of the batch.  compat_q1=True reproduces the reference's class lookup by
    in-stride index and its stride-major output order; False = class of the box itself, box order."""
    lib = _lib.load()
    n = batch.n
    if out is None:
        cmax = int(batch.map_chw.reshape(3, 3)[:, 0].max())
        out = alloc_fmap_scores(n, batch.boxes.device, cmax, want_pooled, want_plan)
    ws = _workspace(batch, table.nc)
    fn, name = (lib.oodb200_fmap_score_nhwc_f32, "oodb200_fmap_score_nhwc_f32") if batch.nhwc else (lib.oodb200_fmap_score_f32, "oodb200_fmap_score_f32")
    _lib.check(fn(
        _ptr(batch.map_ptrs), batch.map_chw.ctypes.data_as(C.c_void_p), batch.scale.ctypes.data_as(C.c_void_p),
        batch.n_img, _ptr(batch.boxes), _ptr(batch.img_idx), _ptr(batch.stride_idx), _ptr(batch.cls),
        _ptr(batch.img_start), int(bool(compat_q1)), n, int(metric_mask), int(bool(normalize)),
        _ptr(table.cent), _ptr(table.cent_unit), _ptr(table.cent_off), _ptr(table.cent_k), table.nc, _ptr(table.thr),
        _ptr(out.dist), _ptr(out.argmin), _ptr(out.decision),
        _ptr(out.pooled), int(out.pooled.stride(0)) if out.pooled is not None else 0,
        _ptr(out.cls_used), _ptr(out.out_index), _ptr(ws), int(ws.numel()), _stream()),
        name)
    return out


@dataclass
class LogitScores:
    scores: torch.Tensor      # [5, n] f32
    indness: Optional[torch.Tensor]
    decision: Optional[torch.Tensor]
    sigmoid_mismatch: torch.Tensor   # int32[1]


def logit_score(logits: torch.Tensor, cls: torch.Tensor, method_mask: int, t_energy: float = 1.0,
                t_odin: float = 1000.0, thr: Optional[torch.Tensor] = None, smin: Optional[torch.Tensor] = None,
                smax: Optional[torch.Tensor] = None, clip: bool = True, out: Optional[LogitScores] = None) -> LogitScores:
    """K3: every requested logit method in one pass. thr/smin/smax: float64 [5, nc] device tensors."""
    lib = _lib.load()
    dev = logits.device
    if logits.dtype != torch.float32 or cls.dtype != torch.int32 or not logits.is_cuda:
        raise TypeError("logit_score: logits must be float32 and cls int32 CUDA tensors")
    logits = logits.contiguous()
    n, nc = int(logits.shape[0]), int(logits.shape[1])
    if out is None:
        out = LogitScores(scores=torch.zeros((5, n), dtype=torch.float32, device=dev),
                          indness=torch.zeros((5, n), dtype=torch.float32, device=dev) if smin is not None else None,
                          decision=torch.ones((5, n), dtype=torch.uint8, device=dev) if thr is not None else None,
                          sigmoid_mismatch=torch.zeros(1, dtype=torch.int32, device=dev))
    _lib.check(lib.oodb200_logit_score_f32(
        _ptr(logits), _ptr(cls), n, nc, int(method_mask), float(t_energy), float(t_odin), _ptr(thr), _ptr(smin),
        _ptr(smax), int(bool(clip)), _ptr(out.scores), _ptr(out.indness), _ptr(out.decision),
        _ptr(out.sigmoid_mismatch), _stream()), "oodb200_logit_score_f32")
    return out


def fuse_decisions(a: torch.Tensor, b: torch.Tensor, strategy: str, c: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty_like(a)
    _lib.check(lib.oodb200_fuse_u8(_ptr(a), _ptr(b), _ptr(c), int(a.numel()), FUSE_SLOT[strategy], _ptr(out), _stream()),
               "oodb200_fuse_u8")
    return out


def fuse_scores(s1: torch.Tensor, s2: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(s1.shape, dtype=torch.uint8, device=s1.device)
    _lib.check(lib.oodb200_fuse_score_f32(_ptr(s1), _ptr(s2), int(s1.numel()), _ptr(out), _stream()),
               "oodb200_fuse_score_f32")
    return out


def vec_score(x: torch.Tensor, seg_off: Sequence[int], cent: torch.Tensor, cent_unit: Optional[torch.Tensor],
              cent_row_off: Sequence[int], cent_k: Sequence[int], metric_mask: int, normalize: bool = True,
              thr: Optional[torch.Tensor] = None):
    """K2 on already pooled vectors: x [n, D] (device), rows of segment g = [seg_off[g], seg_off[g+1]) scored against
    cent rows cent_row_off[g] .. +cent_k[g].  -> (dist [3, n] f32, argmin [3, n] i32), plus decision [3, n] u8 when
    `thr` (float64 [3, n_seg], NaN = no threshold) is given."""
    lib = _lib.load()
    dev = x.device
    for name, ten in (("x", x), ("cent", cent), ("cent_unit", cent_unit)):
        if ten is not None and (ten.dtype != torch.float32 or ten.device != dev):
            raise TypeError(f"vec_score: {name} must be a float32 tensor on {dev}, got {ten.dtype} on {ten.device}")
    x = x.contiguous()
    n, dim = int(x.shape[0]), int(x.shape[1])
    n_seg = len(seg_off) - 1
    dist = torch.empty((3, n), dtype=torch.float32, device=dev)
    arg = torch.empty((3, n), dtype=torch.int32, device=dev)
    dec = None
    if thr is not None:
        if thr.dtype != torch.float64 or tuple(thr.shape) != (3, n_seg) or thr.device != dev:
            raise TypeError("vec_score: thr must be a float64 [3, n_seg] tensor on the device of x")
        thr = thr.contiguous()
        dec = torch.zeros((3, n), dtype=torch.uint8, device=dev)
    t = lambda a, dt: torch.tensor(list(a), dtype=dt, device=dev)
    off_d, crow_d, ck_d = t(seg_off, torch.int64), t(cent_row_off, torch.int64), t(cent_k, torch.int32)   # keep alive
    cent = cent.contiguous()
    _lib.check(lib.oodb200_vec_score_f32(_ptr(x), int(x.stride(0)), dim, _ptr(off_d), n_seg, n,
                                         int(metric_mask), int(bool(normalize)), _ptr(cent),
                                         _ptr(cent_unit), _ptr(crow_d), _ptr(ck_d), _ptr(dist),
                                         _ptr(arg), _ptr(thr), _ptr(dec), _stream()), "oodb200_vec_score_f32")
    if thr is not None:
        return dist, arg, dec
    return dist, arg


def vec_score_tc(x: torch.Tensor, seg_off: Sequence[int], cent: torch.Tensor, cent_row_off: Sequence[int],
                 cent_k: Sequence[int], metric: str, thr: Optional[torch.Tensor] = None, out=None):
    """K2b: like vec_score for ONE of 'l2' / 'cosine' with up to 64 centroids per segment, the x.c cross-term on the
    tensor cores (csrc/kmeans_tc.cu::vec_score_tc_kernel).  x [n, D] float32 contiguous on the device, already
    normalised when the method normalises (ops.normalize_rows); for 'cosine' pass the unit-norm centroid rows.
    -> (dist [3, n], argmin [3, n][, decision [3, n]]) with only the metric's slot written (others: NaN / -1 / 0)."""
    from . import kmeans
    lib = _lib.load()
    dev = x.device
    slot = METRIC_SLOT[metric]
    if metric not in ("l2", "cosine"):
        raise ValueError("vec_score_tc: metric must be 'l2' or 'cosine'")
    for name, ten in (("x", x), ("cent", cent)):
        if ten.dtype != torch.float32 or ten.device != dev:
            raise TypeError(f"vec_score_tc: {name} must be a float32 tensor on {dev}")
    x, cent = x.contiguous(), cent.contiguous()
    n, dim = int(x.shape[0]), int(x.shape[1])
    n_seg = len(seg_off) - 1
    ws_bytes = int(lib.oodb200_vec_score_tc_workspace_bytes(max(n_seg, 1), n, dim))
    if not ws_bytes:
        raise ValueError(f"vec_score_tc: dim must be a multiple of 32 in [128, 2048], got {dim}")
    if out is None:
        dist = torch.full((3, n), float("nan"), dtype=torch.float32, device=dev)
        arg = torch.full((3, n), -1, dtype=torch.int32, device=dev)
    else:
        dist, arg = out
    dec = None
    if thr is not None:
        if thr.dtype != torch.float64 or tuple(thr.shape) != (3, n_seg) or thr.device != dev:
            raise TypeError("vec_score_tc: thr must be a float64 [3, n_seg] tensor on the device of x")
        thr = thr.contiguous()
        dec = torch.zeros((3, n), dtype=torch.uint8, device=dev)
    sizes = [int(seg_off[g + 1] - seg_off[g]) for g in range(n_seg)]
    table, _, _ = kmeans.build_blocks(sizes, 1, 0, dev)
    t = lambda a, dt: torch.tensor(list(a), dtype=dt, device=dev)
    crow_d, ck_d = t(cent_row_off, torch.int64), t(cent_k, torch.int32)
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 255) & ~255
    _lib.check(lib.oodb200_vec_score_tc_f32(_ptr(x), n, dim, n_seg, slot, _ptr(cent), _ptr(crow_d), _ptr(ck_d),
                                            int(max(list(cent_k) + [0])), _ptr(table.seg), _ptr(table.row0), _ptr(table.row1),
                                            table.n_blocks, _ptr(dist), _ptr(arg), _ptr(thr), _ptr(dec), C.c_void_p(ws_ptr),
                                            _stream()), "oodb200_vec_score_tc_f32")
    if thr is not None:
        return dist, arg, dec
    return dist, arg


TC_SCORE_MIN_ROWS = 4096      # below this the FP32 kernel's single launch wins over prep + norms + tensor-core launch


def vec_score_one(x: torch.Tensor, seg_off: Sequence[int], cent: torch.Tensor, cent_unit: Optional[torch.Tensor],
                  cent_row_off: Sequence[int], cent_k: Sequence[int], slot: int, normalize: bool = True,
                  thr: Optional[torch.Tensor] = None, tensor_core: Optional[bool] = None):
    """One metric of K2 on pooled vectors.  Default: the FP32 kernel (csrc/fit_kernels.cu::vec_score_fast_kernel), whose
    per-lane accumulation and reduction tree are those of the fused per-box scorer -- distances that feed a threshold
    (fit) and distances compared with it (decisions) come from ONE arithmetic.  `tensor_core=True` (or
    OODB200_VEC_TC=1 when the argument is None) sends 'l2' / 'cosine' to the tcgen05 kernel (vec_score_tc) when the
    shape fits (D % 32 == 0, 128 <= D <= 2048, <= 64 centroids per segment, >= TC_SCORE_MIN_ROWS rows): the 1e-3 tier of
    BASELINE.json, meant for the K = 64 dense-contraction case (C5).  Same return value as vec_score."""
    import os
    name = {v: k for k, v in METRIC_SLOT.items()}[slot]
    n, dim = int(x.shape[0]), int(x.shape[1])
    if tensor_core is None:
        tensor_core = os.environ.get("OODB200_VEC_TC", "0") == "1"
    if (tensor_core and name in ("l2", "cosine") and n >= TC_SCORE_MIN_ROWS and dim % 32 == 0 and 128 <= dim <= 2048
            and max(list(cent_k) + [0]) <= 64):
        xs = normalize_rows(x) if normalize else x
        return vec_score_tc(xs, seg_off, cent_unit if name == "cosine" else cent, cent_row_off, cent_k, name, thr=thr)
    return vec_score(x, seg_off, cent, cent_unit, cent_row_off, cent_k, 1 << slot, normalize=normalize, thr=thr)


def dist_indness(dist: torch.Tensor, slot: torch.Tensor, thr: torch.Tensor, dmin: torch.Tensor, dmax: torch.Tensor,
                 clip: bool = True) -> torch.Tensor:
    """Intended `DistanceMethod.compute_indness` (ood_utils.py:1599-1604) for n distances; slot [n] int32 indexes the
    float64 thr / dmin / dmax tables."""
    lib = _lib.load()
    out = torch.empty(dist.shape, dtype=torch.float32, device=dist.device)
    _lib.check(lib.oodb200_dist_indness_f32(_ptr(dist.contiguous()), _ptr(slot.contiguous()), int(dist.numel()),
                                            _ptr(thr), _ptr(dmin), _ptr(dmax), int(bool(clip)), _ptr(out), _stream()),
               "oodb200_dist_indness_f32")
    return out


def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """sklearn normalize(axis=1) of float32 rows on the device (ood_utils.py:2404-2409)."""
    lib = _lib.load()
    if x.dtype != torch.float32 or not x.is_cuda:
        raise TypeError("normalize_rows: float32 CUDA tensor expected")
    x = x.contiguous()
    out = torch.empty_like(x)
    _lib.check(lib.oodb200_normalize_rows_f32(_ptr(x), int(x.stride(0)), int(x.shape[1]), int(x.shape[0]), _ptr(out),
                                              int(out.stride(0)), _stream()), "oodb200_normalize_rows_f32")
    return out


# ------------------------------------------------------------------------------------------------ fit-data collection
def match_boxes(pred_xyxy: Sequence, pred_cls: Sequence, gt_xyxy: Sequence, gt_cls: Sequence, iou_threshold: float,
                compat: bool = True, device=None):
    """Per image: IoU x same-class score matrix, scipy-compatible assignment and the valid predictions
    (/root/reference/ood_utils.py:233-292) in ONE launch for the batch.  Inputs are per-image sequences ([P_i, 4], [P_i],
    [G_i, 4], [G_i]; torch tensors on any device or arrays).  -> list over images of (valid_preds list, score [P, G] float32
    numpy, (row_ind, col_ind) int64 numpy), what the reference stores on every Results object."""
    lib = _lib.load()
    device = device or default_device()
    n_img = len(pred_xyxy)

    def flat(seq, dtype, width=None):
        ts = [t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t)) for t in seq]
        shape = (-1, width) if width else (-1,)
        if not ts:
            return torch.zeros((0, width) if width else (0,), dtype=dtype, device=device)
        if all(not t.is_cuda for t in ts):
            return h2d(torch.cat([t.reshape(shape).to(dtype) for t in ts]), device, dtype)
        return torch.cat([t.to(device).reshape(shape) for t in ts]).to(dtype).contiguous()

    pc = [int(len(b)) for b in pred_xyxy]
    gc = [int(len(b)) for b in gt_xyxy]
    p_start = np.zeros(n_img + 1, np.int32)
    g_start = np.zeros(n_img + 1, np.int32)
    s_off = np.zeros(n_img + 1, np.int64)
    np.cumsum(pc, out=p_start[1:])
    np.cumsum(gc, out=g_start[1:])
    np.cumsum([a * b for a, b in zip(pc, gc)], out=s_off[1:])
    n, tot = int(p_start[-1]), int(s_off[-1])
    pb, pcl = flat(pred_xyxy, torch.float32, 4), flat(pred_cls, torch.int32)
    gb, gcl = flat(gt_xyxy, torch.float32, 4), flat(gt_cls, torch.int32)
    score = torch.zeros(max(tot, 1), dtype=torch.float32, device=device)
    rows = torch.full((max(n, 1),), -1, dtype=torch.int32, device=device)
    cols = torch.full((max(n, 1),), -1, dtype=torch.int32, device=device)
    valid = torch.zeros(max(n, 1), dtype=torch.uint8, device=device)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    ps_d, gs_d, so_d = h2d(p_start, device), h2d(g_start, device), h2d(s_off, device)
    _lib.check(lib.oodb200_match_boxes_f32(_ptr(pb), _ptr(pcl), _ptr(ps_d), _ptr(gb), _ptr(gcl), _ptr(gs_d), _ptr(so_d), n_img,
                                           float(iou_threshold), int(bool(compat)), _ptr(score), _ptr(rows), _ptr(cols),
                                           _ptr(valid), _ptr(status), _stream()), "oodb200_match_boxes_f32")
    if int(status.item()):
        raise ValueError("match_boxes: an image has more than 1024 boxes, or a box with zero area gives a non-finite IoU "
                         "(scipy's linear_sum_assignment rejects such matrices too)")
    score_h, rows_h, cols_h, valid_h = score.cpu().numpy(), rows.cpu().numpy(), cols.cpu().numpy(), valid.cpu().numpy()
    out = []
    for i in range(n_img):
        a, b = int(p_start[i]), int(p_start[i + 1])
        m = min(pc[i], gc[i])
        out.append((np.nonzero(valid_h[a:b])[0].tolist(), score_h[s_off[i]:s_off[i + 1]].reshape(pc[i], gc[i]),
                    (rows_h[a:a + m].astype(np.int64), cols_h[a:a + m].astype(np.int64))))
    return out


# ------------------------------------------------------------------------------------------------ k-search scores (K7)
def pair_cluster_sums(x: torch.Tensor, labels: torch.Tensor, kc: int, metric: str) -> torch.Tensor:
    """sums[i][c] = sum of dist(x_i, x_j) over the rows j with labels[j] == c (float64 [n, kc]); x float32 [n, D] rows of
    one segment (unit-norm rows for 'cosine'), labels int32 in [0, kc)."""
    lib = _lib.load()
    if x.dtype != torch.float32 or not x.is_cuda or labels.dtype != torch.int32 or not labels.is_cuda:
        raise TypeError("pair_cluster_sums: float32 rows and int32 labels on the device expected")
    x = x.contiguous()
    n, d = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((n, int(kc)), dtype=torch.float64, device=x.device)
    _lib.check(lib.oodb200_pair_cluster_sums_f32(_ptr(x), n, d, int(x.stride(0)), _ptr(labels.contiguous()), int(kc),
                                                 METRIC_SLOT[metric], _ptr(out), _stream()), "oodb200_pair_cluster_sums_f32")
    return out


class PairDistances:
    """The pair distances of ONE set of rows, shared by every labeling the k-search scores (cluster_utils.py:203-302 calls
    `silhouette_score` once per candidate k on the same rows).  When the n x n float32 matrix fits the budget
    (OODB200_PAIR_MATRIX_GB, default 8 GB: n <= 46 340) it is computed once (K7 in store mode) and every labeling costs one
    pass over it; otherwise every labeling recomputes the distances (K7 in fold mode, nothing stored)."""

    def __init__(self, x: torch.Tensor, metric: str, max_bytes: Optional[int] = None):
        if x.dtype != torch.float32 or not x.is_cuda:
            raise TypeError("PairDistances: float32 rows on the device expected")
        self.metric = metric
        self.xs = (normalize_rows(x) if metric == "cosine" else x).contiguous()
        self.n = int(x.shape[0])
        if max_bytes is None:
            max_bytes = int(float(os.environ.get("OODB200_PAIR_MATRIX_GB", "8")) * (1 << 30))
        self.matrix = None
        if 0 < 4 * self.n * self.n <= max_bytes:
            lib = _lib.load()
            self.matrix = torch.empty((self.n, self.n), dtype=torch.float32, device=x.device)
            _lib.check(lib.oodb200_pair_dist_matrix_f32(_ptr(self.xs), self.n, int(self.xs.shape[1]), int(self.xs.stride(0)),
                                                        METRIC_SLOT[metric], _ptr(self.matrix), self.n, _stream()),
                       "oodb200_pair_dist_matrix_f32")

    def cluster_sums(self, labels: torch.Tensor, kc: int) -> torch.Tensor:
        """sums[i][c] like pair_cluster_sums; labels int32 in [0, kc) on the device."""
        if self.matrix is None:
            return pair_cluster_sums(self.xs, labels, kc, self.metric)
        lib = _lib.load()
        order = torch.argsort(labels, stable=True).to(torch.int32)
        off = torch.zeros(int(kc) + 1, dtype=torch.int64, device=labels.device)
        off[1:] = torch.cumsum(torch.bincount(labels.long(), minlength=int(kc)), 0)
        out = torch.empty((self.n, int(kc)), dtype=torch.float64, device=labels.device)
        _lib.check(lib.oodb200_matrix_cluster_sums_f32(_ptr(self.matrix), self.n, self.n, _ptr(order), _ptr(off), int(kc),
                                                       _ptr(out), _stream()), "oodb200_matrix_cluster_sums_f32")
        return out


def silhouette_score(x: torch.Tensor, labels: torch.Tensor, metric: str, pairs: Optional[PairDistances] = None) -> float:
    """sklearn.metrics.silhouette_score(X, labels, metric='l1' | 'l2' | 'cosine') (cluster_utils.py:277): mean over the
    samples of (b - a) / max(a, b), a = mean distance to the other members of the own cluster, b = smallest mean distance
    to another cluster; members of one-sample clusters count 0.  The pair distances are summed per cluster on the GPU
    (K7); `pairs` = the distances of these rows kept from an earlier call (PairDistances) so that a search over k computes
    them once."""
    uniq, enc = torch.unique(labels, return_inverse=True)
    kc, n = int(uniq.numel()), int(labels.numel())
    if not 1 < kc < n:
        raise ValueError(f"Number of labels is {kc}. Valid values are 2 to n_samples - 1 (inclusive)")
    enc32 = enc.to(torch.int32)
    if pairs is None:
        pairs = PairDistances(x, metric, max_bytes=0)          # one labeling: fold, nothing stored
    sums = pairs.cluster_sums(enc32, kc)
    freq = torch.bincount(enc, minlength=kc).to(torch.float64)
    rows = torch.arange(n, device=x.device)
    intra = sums[rows, enc] / (freq - 1.0)[enc]
    sums[rows, enc] = float("inf")
    inter = (sums / freq).min(dim=1).values
    sil = torch.nan_to_num((inter - intra) / torch.maximum(intra, inter), nan=0.0)
    return float(sil.mean())


def calinski_harabasz_score(x: torch.Tensor, labels: torch.Tensor) -> float:
    """sklearn.metrics.calinski_harabasz_score (cluster_utils.py:280): between- over within-cluster dispersion."""
    uniq, enc = torch.unique(labels, return_inverse=True)
    k, n = int(uniq.numel()), int(labels.numel())
    x64 = x.to(torch.float64)
    freq = torch.bincount(enc, minlength=k).to(torch.float64)
    sums = torch.zeros((k, x.shape[1]), dtype=torch.float64, device=x.device).index_add_(0, enc, x64)
    means = sums / freq[:, None]
    extra = float((freq * ((means - x64.mean(0)) ** 2).sum(1)).sum())
    intra = float(((x64 - means[enc]) ** 2).sum())
    return 1.0 if intra == 0.0 else extra * (n - k) / (intra * (k - 1.0))
