"""Exact `np.percentile(scores, q, method='lower')` per segment on the GPU (K5).

Replaces /root/reference/ood_utils.py:613,626 (`generate_thresholds`).  numpy returns the order statistic with index
floor((n-1) * (q / dtype(100))) of the sorted scores, the quantile being rounded to the DATA dtype
(numpy/lib/_function_base_impl.py:4277 and :141-144).  The host computes that index with numpy's own expression; the
device finds the element with a three-pass radix select (11+11+10 bits of an order-preserving key) whose per-pass
histograms are summed across ranks with one all-reduce each, so the result is exact for any sharding.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib

PASSES = ((21, 11), (10, 11), (0, 10))          # (shift, bits)


def lower_index(n: int, q: float, dtype=np.float32) -> int:
    """Index numpy uses for percentile(q, method='lower') on `n` values of `dtype` (q is a python float)."""
    quant = np.true_divide(float(q), np.dtype(dtype).type(100))
    return int(np.floor((int(n) - 1) * quant).astype(np.intp))


def _key_to_float(key: np.ndarray) -> np.ndarray:
    key = key.astype(np.uint32)
    bits = np.where(key & np.uint32(0x80000000), key & np.uint32(0x7FFFFFFF), ~key)
    return bits.astype(np.uint32).view(np.float32)


class _CudaHist:
    """One radix pass through the C ABI (oodb200_radix_hist_u32)."""

    def radix_hist(self, scores, seg_off, prefix, shift, bits, hist, minmax):
        lib = _lib.load()
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        off_d = torch.tensor(list(seg_off), dtype=torch.int64, device=scores.device)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.oodb200_radix_hist_u32(ptr(scores), ptr(off_d), len(seg_off) - 1, int(scores.numel()), ptr(prefix),
                                              shift, bits, ptr(hist), ptr(minmax), stream), "oodb200_radix_hist_u32")


def segment_select(scores: torch.Tensor, seg_off: Sequence[int], ranks: Sequence[Optional[int]], group=None, backend=None):
    """scores [n_local] float32 on the device, seg_off local row offsets [n_seg+1], ranks[g] = GLOBAL order-statistic
    index wanted in segment g (None: skip).  -> (values[float or None], min[float], max[float]) per segment;
    min/max are None for empty segments.  `backend` (tests only): object with `radix_hist`, so that the multi-rank
    exchange can be exercised on CPU under gloo."""
    import torch.distributed as dist
    backend = backend or _CudaHist()
    dev = scores.device
    n_seg = len(seg_off) - 1
    distributed = group is not None                    # explicit: a local selection inside an initialised job stays local
    scores = scores.contiguous()
    want = np.array([-1 if r is None else int(r) for r in ranks], dtype=np.int64)
    prefix = np.zeros(n_seg, dtype=np.uint32)
    remaining = want.copy()
    minmax_i = torch.stack([torch.full((n_seg,), -1, dtype=torch.int32, device=dev),        # 0xFFFFFFFF
                            torch.zeros(n_seg, dtype=torch.int32, device=dev)], dim=1).contiguous()
    for p, (shift, bits) in enumerate(PASSES):
        hist = torch.zeros((n_seg, 1 << bits), dtype=torch.int32, device=dev)
        pfx = torch.from_numpy(prefix.view(np.int32).copy()).to(dev)
        backend.radix_hist(scores, list(seg_off), pfx, shift, bits, hist, minmax_i if p == 0 else None)
        if distributed:
            dist.all_reduce(hist, group=group)                              # one exchange per radix pass
        h = hist.cpu().numpy().astype(np.int64)
        for g in range(n_seg):
            if want[g] < 0:
                continue
            cum = np.cumsum(h[g])
            b = int(np.searchsorted(cum, remaining[g], side="right"))
            if b >= (1 << bits):
                raise RuntimeError(f"radix select: rank {want[g]} beyond the {int(cum[-1])} scores of segment {g}")
            remaining[g] -= (cum[b - 1] if b else 0)
            prefix[g] = (np.uint32(prefix[g]) << np.uint32(bits)) | np.uint32(b)
    mmk = minmax_i.to(torch.int64)
    if distributed:
        lo, hi = (mmk[:, 0] & 0xFFFFFFFF).contiguous(), (mmk[:, 1] & 0xFFFFFFFF).contiguous()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        mmk = torch.stack([lo, hi], dim=1)
    mmk = (mmk.cpu().numpy() & 0xFFFFFFFF).astype(np.uint32)
    vals = _key_to_float(prefix)
    out_v: List[Optional[float]] = [float(vals[g]) if want[g] >= 0 else None for g in range(n_seg)]
    nonempty = mmk[:, 0] <= mmk[:, 1]
    mn = _key_to_float(mmk[:, 0])
    mx = _key_to_float(mmk[:, 1])
    return out_v, [float(mn[g]) if nonempty[g] else None for g in range(n_seg)], \
        [float(mx[g]) if nonempty[g] else None for g in range(n_seg)]
