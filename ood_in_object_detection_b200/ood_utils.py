"""OoD-method class surface of the reference, backed by the sm_100a kernels of liboodb200.so.

Drop-in for the names `ood_evaluation.py` imports from the reference's `ood_utils.py`
(/root/reference/ood_evaluation.py:19-21): same class names, constructor arguments, method names, return types
(python lists / numpy arrays / python floats) and error behaviour (asserts for configuration, ValueError /
NotImplementedError for unsupported modes, logger warnings for data conditions).  What differs is the inside:
the per-image / per-stride / per-box python loops of the reference are ONE batched pass through the CUDA kernels
(ops.py -> include/oodb200.h).  There is no CPU fallback; every scoring call needs a CUDA device.

Reference quirks are reproduced when `reference_compat` is True (the default), because the
bit-exact-decisions bar is defined against the reference as shipped:
  Q1  the class used for the centroid / threshold lookup is the class of the box with the same IN-STRIDE index and
      decisions come out stride-major per image (ood_utils.py:2152-2154)
  Q2  DistanceMethod INDness is always -1 (ood_utils.py:1598, :1609-1612)
  Q4  falsy thresholds ([] / 0 / 0.0) mean "no threshold -> OoD" (ood_utils.py:2173)
  Q8  fit-side matching indexes the score matrix with the position in the assignment list (ood_utils.py:288-291)
Set `method.reference_compat = False` for the intended semantics (class of the box itself, box order, piecewise
linear distance INDness).

Out of the hot path and not rebuilt here: enhanced unknown localisation (EUL), the SDR
reducers' training (ivis / umap are CPU/TF libraries), plotting, metric computation (`compute_metrics` is a hook).
"""
from __future__ import annotations

import inspect
import time
from abc import ABC, abstractmethod
from datetime import timedelta
from logging import Logger
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from . import cluster_utils as _cluster_utils
from . import kmeans as _kmeans
from . import ops
from . import select as _select
from .constants import (AVAILABLE_CLUSTER_OPTIMIZATION_METRICS, FTMAPS_RELATED_OPTIONS, IND_INFO_CREATION_OPTIONS,
                        INTERNAL_ACTIVATIONS_EXTRACTION_OPTIONS, LOGITS_RELATED_OPTIONS, UNKNOWN_CLASS_INDEX,
                        is_valid_cluster_method, kmeans_k)
from .custom_hyperparams import CUSTOM_HYP

# `compute_metrics(all_preds, all_targets, class_names, known_classes, logger) -> dict` of the evaluation harness
# (/root/reference/ood_utils.py:567 calls the one of its own metrics module).  Metric computation is outside the
# scoring hot path; assign the harness' function here before `iterate_data_to_compute_metrics`.
compute_metrics: Optional[Callable] = None


# ------------------------------------------------------------------------------------------------ helpers
def _np(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def _rows_2d(a) -> Union[np.ndarray, torch.Tensor]:
    """[N, C, 1, 1] / [N, C] / [C, 1, 1]-list -> [N, D] view (numpy or torch, no copy when possible)."""
    if isinstance(a, (list, tuple)):
        a = np.stack([_np(v) for v in a], axis=0) if len(a) else np.empty((0, 0), np.float32)
    if a.ndim == 1 and a.shape[0] == 0:                    # format_internal_activations stores np.empty(0) for "no rows"
        return a.reshape(0, 0)
    return a.reshape(a.shape[0], -1) if a.ndim != 2 else a


def _to_device_f32(a, device) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def _split_lists(flat: np.ndarray, counts: Sequence[int], cast) -> List[list]:
    """flat per-box values -> per-image python lists of `cast` (int / float) values, like the reference returns."""
    flat = np.asarray(flat)
    if cast is int:
        vals = flat.astype(np.int64, copy=False).tolist()
    elif cast is float:
        vals = flat.astype(np.float64, copy=False).tolist()
    else:
        vals = [cast(v) for v in flat]
    out, pos = [], 0
    for m in counts:
        out.append(vals[pos:pos + m])
        pos += m
    return out


def _img_hw(res) -> Tuple[int, int]:
    """(H, W) of the network input: `res.orig_img.shape[1:3]` (ood_utils.py:2061; orig_img is the whole batch)."""
    shp = tuple(res.orig_img.shape)
    return int(shp[1]), int(shp[2])


def extract_roi_aligned_features_from_correct_stride(ftmaps: List[Tensor], boxes: List[Tensor], strides: List[Tensor],
                                                     img_shape, device=None, extract_all_strides: bool = False):
    """K1 behind the reference helper's signature (/root/reference/ultralytics/models/yolo/detect/predict.py:13-90).

    ftmaps: 3 batched tensors [N, C_s, H_s, W_s]; boxes: N tensors [M_i, 4] xyxy in input pixels; strides: N tensors
    [M_i] in {0,1,2}.  Returns out[img][stride] = [idx_in_img (int16 [m]), feats [m, C_s, 1, 1]] (device tensors)."""
    device = torch.device(device) if device is not None and str(device) != "cpu" else ops.default_device()
    n_img = len(boxes)
    cls0 = [torch.zeros(len(b), dtype=torch.int32) for b in boxes]
    out = [[[None, None] for _ in range(3)] for _ in range(n_img)]
    passes = range(3) if extract_all_strides else (None,)
    counts = [int(len(b)) for b in boxes]
    start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    img_of = np.repeat(np.arange(n_img), counts)
    for forced in passes:
        st = strides if forced is None else [torch.full((len(b),), forced, dtype=torch.int32) for b in boxes]
        batch = ops.make_batch(list(ftmaps), boxes, st, cls0, int(img_shape[1]), device)
        pooled = ops.roi_pool(batch)
        chw = batch.map_chw.reshape(3, 3)
        st_host = batch.stride_idx.cpu().numpy().astype(np.int64) if forced is None else np.full(int(start[-1]), forced, np.int64)  # ONE read
        for s in (range(3) if forced is None else (forced,)):
            sel = np.nonzero(st_host == s)[0]                      # box rows of this stride, image-major
            per_img = np.bincount(img_of[sel], minlength=n_img) if len(sel) else np.zeros(n_img, np.int64)
            cut = np.concatenate([[0], np.cumsum(per_img)]).astype(np.int64)
            C = int(chw[s, 0])
            if len(sel):                                           # one gather per stride for the whole batch; per image: views
                rows = pooled.index_select(0, torch.from_numpy(sel).to(device))[:, :C].reshape(len(sel), C, 1, 1)
                idx_all = torch.from_numpy((sel - start[img_of[sel]]).astype(np.int16)).to(device)
            for i in range(n_img):
                a, b = int(cut[i]), int(cut[i + 1])
                if b > a:
                    out[i][s] = [idx_all[a:b], rows[a:b]]
                else:
                    out[i][s] = [torch.zeros(0, dtype=torch.int16, device=device), torch.empty(0, device=device)]
    return out


# ------------------------------------------------------------------------------------------------ base class
class OODMethod(ABC):
    """Base of every OoD method (contract: /root/reference/ood_utils.py:44-130).  1 = InD, 0 = OoD."""

    reference_compat: bool = True      # reproduce SURVEY.md quirks Q1 / Q2 (see module docstring)

    def __init__(self, name: str, is_distance_method: bool, per_class: bool, per_stride: bool,
                 iou_threshold_for_matching: float, min_conf_threshold_train: float, min_conf_threshold_test: float,
                 which_internal_activations: str, enhanced_unk_localization: bool = False,
                 saliency_map_computation_function: Callable = None,
                 thresholds_out_of_saliency_map_function: Callable = None, **kwargs):
        self.name = name
        self.is_distance_method = is_distance_method
        self.per_class = per_class
        self.per_stride = per_stride
        self.iou_threshold_for_matching = iou_threshold_for_matching
        self.min_conf_threshold_train = min_conf_threshold_train
        self.min_conf_threshold_test = min_conf_threshold_test
        self.thresholds = None
        self.which_internal_activations = self.validate_internal_activations_option(which_internal_activations)
        self.enhanced_unk_localization = enhanced_unk_localization
        if enhanced_unk_localization:
            self.compute_saliency_map_one_stride = self.validate_saliency_map_computation_function(saliency_map_computation_function)
            self.compute_thresholds_out_of_saliency_map = self.validate_thresholds_out_of_saliency_map_function(thresholds_out_of_saliency_map_function)
        self.use_values_before_sigmoid = False

    # -- validators (ood_utils.py:98-128)
    @staticmethod
    def validate_internal_activations_option(selected_option: str):
        assert selected_option in INTERNAL_ACTIVATIONS_EXTRACTION_OPTIONS, f"Invalid option selected ({selected_option}) for " \
            f"internal activations extraction. Options are: {INTERNAL_ACTIVATIONS_EXTRACTION_OPTIONS}"
        return selected_option

    @staticmethod
    def validate_saliency_map_computation_function(passed_function: Callable) -> Callable:
        assert callable(passed_function), "The passed function is not a callable"
        params = inspect.signature(passed_function).parameters
        assert len(list(params.keys())) == 1, "The passed function must accept only one argument"
        assert params[list(params.keys())[0]].annotation == np.ndarray, "The passed function must accept a Tensor as input"
        assert passed_function(np.random.rand(5, 40, 40)).shape == (40, 40), "The passed function must convert (C, H, W) to (H, W)"
        return passed_function

    @staticmethod
    def validate_thresholds_out_of_saliency_map_function(passed_function: Callable) -> Callable:
        assert callable(passed_function), "The passed function is not a callable"
        params = inspect.signature(passed_function).parameters
        assert params[list(params.keys())[0]].annotation == np.ndarray, \
            "The passed function first argument must be the saliency map and accept a Tensor or np.ndarray as input"
        assert isinstance(passed_function(np.random.rand(80, 80)), list), "The passed function must return a list with the thresholds"
        return passed_function

    # -- abstract surface (ood_utils.py:131-194)
    @abstractmethod
    def extract_internal_activations(self, results, all_activations, targets): ...

    @abstractmethod
    def format_internal_activations(self, all_activations): ...

    @abstractmethod
    def compute_ood_decision_on_results(self, results, logger) -> List[List[int]]: ...

    @abstractmethod
    def compute_scores(self, activations, *args, **kwargs) -> np.ndarray: ...

    @abstractmethod
    def activations_transformation(self, activations, **kwargs): ...

    @abstractmethod
    def compute_distance(self, centroids, features): ...

    # -- data plumbing around the detector (host logic, same contracts as ood_utils.py:196-347)
    @staticmethod
    def log_every_n_batches(n: int, logger, idx_of_batch: int, number_of_batches: int):
        if idx_of_batch % n == 0:
            logger.info(f"{(idx_of_batch / number_of_batches) * 100:02.1f}%: Procesing batch {idx_of_batch} of {number_of_batches}")

    @staticmethod
    def create_targets_dict(data: Dict) -> Dict[str, List[Tensor]]:
        """targets = {'bboxes': per-image [n, 4] xyxy in absolute pixels, 'cls': per-image [n]} from the flat
        ultralytics batch dict (relative cxcywh boxes + batch_idx), ood_utils.py:201-231."""
        bboxes, cls = [], []
        for img_idx in range(len(data['im_file'])):
            idx = torch.where(data['batch_idx'] == img_idx)
            b = data['bboxes'][idx]
            cx, cy, w, h = b.unbind(-1)
            xyxy = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)
            scale = torch.as_tensor(tuple(data['resized_shape'][img_idx]) + tuple(data['resized_shape'][img_idx]))
            bboxes.append(xyxy * scale)
            cls.append(data['cls'][idx].view(-1))
        return dict(bboxes=bboxes, cls=cls)

    @staticmethod
    def match_predicted_boxes_to_targets(results, targets, iou_threshold: float, compat: bool = True):
        """`res.valid_preds` = predictions matched (Hungarian on IoU x same-class mask) to a ground-truth box with
        IoU above the threshold (ood_utils.py:233-292).  The whole batch is one launch (csrc/matching.cu): score matrices,
        scipy's assignment algorithm with its tie rules, and the walk over the assignment; `res.assignment_score_matrix`
        and `res.assignment` are filled like the reference does.

        compat=True keeps the reference's indexing (Q8): it walks `enumerate(assignment[1])` and tests
        `score[i, col]` with i = POSITION in the assignment, not the assigned row `assignment[0][i]`
        (ood_utils.py:288-291) -- identical when every prediction is assigned (P <= G), different when P > G.
        compat=False tests the assigned (row, col) pairs."""
        if not len(results):
            return
        out = ops.match_boxes([res.boxes.xyxy.detach() for res in results], [res.boxes.cls.detach() for res in results],
                              [targets['bboxes'][i].detach() for i in range(len(results))],
                              [targets['cls'][i].detach() for i in range(len(results))], iou_threshold, compat=compat)
        for res, (valid, score, assignment) in zip(results, out):
            res.assignment_score_matrix = torch.from_numpy(score)
            res.assignment = assignment
            res.valid_preds = valid

    def prepare_data_for_model(self, data, device):
        if isinstance(data, dict):
            return data['img'].to(device), self.create_targets_dict(data)
        imgs, targets = data
        return imgs, targets

    def _empty_activation_lists(self, n_classes: int):
        if not self.per_class:
            raise NotImplementedError("Not implemented yet")
        if self.per_stride:
            return [[[] for _ in range(3)] for _ in range(n_classes)]
        return [[] for _ in range(n_classes)]

    def iterate_data_to_extract_ind_activations(self, data_loader, model, device, logger: Logger):
        """Run the detector over the in-distribution split and collect the activations of the valid predictions
        (ood_utils.py:294-336).  `model.predict(imgs, ...)` must return Results-like objects."""
        logger.warning(f"Using a confidence threshold of {self.min_conf_threshold_train} for training")
        all_internal_activations = self._empty_activation_lists(len(model.names))
        number_of_batches = len(data_loader)
        for idx_of_batch, data in enumerate(data_loader):
            self.log_every_n_batches(50, logger, idx_of_batch, number_of_batches)
            imgs, targets = self.prepare_data_for_model(data, device)
            imgs = imgs.float() / 255
            results = model.predict(imgs, save=False, verbose=False, conf=self.min_conf_threshold_train, device=device)
            self.match_predicted_boxes_to_targets(results, targets, self.iou_threshold_for_matching, compat=self.reference_compat)
            self.extract_internal_activations(results, all_internal_activations, targets)
        self.format_internal_activations(all_internal_activations)
        return all_internal_activations

    def _decide_for_metrics(self, model, imgs, device, logger):
        results = model.predict(imgs, save=False, verbose=False, conf=self.min_conf_threshold_test, device=device)
        return results, self.compute_ood_decision_on_results(results, logger)

    def iterate_data_to_compute_metrics(self, model, device, dataloader, logger: Logger, known_classes: List[int]) -> Dict[str, float]:
        """Detector -> OoD decisions -> predictions with OoD boxes relabelled as class 80 -> `compute_metrics`
        (ood_utils.py:428-580).  The scoring of every batch is one fused GPU pass."""
        if compute_metrics is None:
            raise NotImplementedError("assign ood_utils.compute_metrics (the evaluation harness' metric function) first; "
                                      "metric computation is outside the GPU hot path (INTEGRATION.md)")
        if self.enhanced_unk_localization:
            raise NotImplementedError("enhanced unknown localisation: only the proposal ranking (rank_unknown_proposals, SURVEY.md §8f rank 4) is rebuilt; the saliency / region-proposal stage around it is CPU image processing outside the hot path")
        logger.warning(f"Using a confidence threshold of {self.min_conf_threshold_test} for tests")
        assert hasattr(dataloader.dataset, "number_of_classes"), \
            "The dataset does not have the attribute number_of_classes to know the number of classes known in the dataset"
        class_names = list(dataloader.dataset.data['names'].values())[:dataloader.dataset.number_of_classes]
        class_names.append('unknown')
        known = torch.tensor(known_classes, dtype=torch.float32)
        unk = torch.tensor(UNKNOWN_CLASS_INDEX, dtype=torch.float32)
        all_preds, all_targets, processed = [], [], 0
        number_of_batches = len(dataloader)
        for idx_of_batch, data in enumerate(dataloader):
            if idx_of_batch % 50 == 0 or idx_of_batch == number_of_batches - 1:
                logger.info(f"{(idx_of_batch / number_of_batches) * 100:02.1f}%: Procesing batch {idx_of_batch + 1} of {number_of_batches}")
            imgs, targets = self.prepare_data_for_model(data, device)
            imgs = imgs.float() / 255
            results, ood_decision = self._decide_for_metrics(model, imgs, device, logger)
            for img_idx, res in enumerate(results):
                dec = torch.tensor(ood_decision[img_idx], dtype=torch.float32)
                cls = torch.where(dec == 0, unk, res.boxes.cls.cpu())
                all_preds.append({'img_idx': processed + img_idx, 'img_name': Path(data['im_file'][img_idx]).stem,
                                  'bboxes': res.boxes.xyxy.cpu(), 'cls': cls, 'conf': res.boxes.conf.cpu()})
                tcls = targets['cls'][img_idx]
                all_targets.append({'img_idx': processed + img_idx, 'img_name': Path(data['im_file'][img_idx]).stem,
                                    'bboxes': targets['bboxes'][img_idx],
                                    'cls': torch.where(torch.isin(tcls, known), tcls, unk)})
            processed += len(imgs)
        results_dict = compute_metrics(all_preds, all_targets, class_names, known_classes, logger)
        n_known = sum(int(torch.sum(t['cls'] != UNKNOWN_CLASS_INDEX)) for t in all_targets)
        n_unk = sum(int(torch.sum(t['cls'] == UNKNOWN_CLASS_INDEX)) for t in all_targets)
        logger.info(f"Number of target known boxes: {n_known}")
        logger.info(f"Number of target unknown boxes: {n_unk}")
        return results_dict

    # -- thresholds (ood_utils.py:583-637): exact 'lower' percentile per (class[, stride]) on the GPU (K5)
    def generate_thresholds(self, ind_scores: list, tpr: float, logger: Logger, group=None):
        """`float(np.percentile(scores, q, method='lower'))` per class (and stride) with q = 100*tpr for distance
        methods and (1-tpr)*100 for logits methods; segments with <= MIN_NUMBER_OF_SAMPLES_FOR_THR scores get no
        threshold ([] per stride / 0 per class).  All segments are selected in the same three radix passes.
        `group`: torch.distributed process group when the scores are sharded across ranks (each rank passes its
        local scores; the segment sizes are then summed over ranks)."""
        used_tpr = 100 * tpr if self.is_distance_method else (1 - tpr) * 100
        min_n, good_n = CUSTOM_HYP.MIN_NUMBER_OF_SAMPLES_FOR_THR, CUSTOM_HYP.GOOD_NUM_SAMPLES
        if not self.per_class:
            raise NotImplementedError("Not implemented yet")
        if self.per_stride:
            thresholds = [[[] for _ in range(3)] for _ in range(len(ind_scores))]
            keys = [(c, s) for c in range(len(ind_scores)) for s in range(len(ind_scores[c]))]
            segs = [ind_scores[c][s] for c, s in keys]
        else:
            thresholds = [0 for _ in range(len(ind_scores))]
            keys = [(c, None) for c in range(len(ind_scores))]
            segs = list(ind_scores)
        dev = ops.default_device()
        sizes = [int(len(v)) for v in segs]
        total = list(sizes)
        if group is not None:
            t = torch.tensor(sizes, dtype=torch.int64, device=dev)
            torch.distributed.all_reduce(t, group=group)
            total = [int(v) for v in t.cpu()]
        chunks, ranks, dtypes = [], [], []
        for v, n_loc, n in zip(segs, sizes, total):
            dt = np.float64 if (not isinstance(v, torch.Tensor) and np.asarray(v).dtype == np.float64) else np.float32
            dtypes.append(dt)
            ranks.append(_select.lower_index(n, used_tpr, dt) if n > min_n else None)
            if n_loc:
                chunks.append(_to_device_f32(v, dev).reshape(-1))
        if not any(r is not None for r in ranks):
            vals = [None] * len(segs)
        else:
            flat = torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.float32, device=dev)
            off = np.concatenate([[0], np.cumsum(sizes)]).tolist()
            vals, _, _ = _select.segment_select(flat, off, ranks, group=group)
        for (c, s), n, v in zip(keys, total, vals):
            where = f"Class {c:03}, Stride {s}" if s is not None else f"Class {c}"
            if v is not None:
                if s is None:
                    thresholds[c] = float(v)
                else:
                    thresholds[c][s] = float(v)
                if n < good_n:
                    logger.warning(f"{where}: has {n} samples. The threshold may not be accurate")
            elif c < 20:
                logger.warning(f"{where} -> Has less than {min_n} samples. No threshold is generated")
        return thresholds

    def compute_extra_possible_unkwnown_bboxes_and_decision(self, *args, **kwargs):
        raise NotImplementedError("enhanced unknown localisation: only the proposal ranking (rank_unknown_proposals, SURVEY.md §8f rank 4) is rebuilt; the saliency / region-proposal stage around it is CPU image processing outside the hot path")


# ------------------------------------------------------------------------------------------------ logits family
class LogitsMethod(OODMethod):
    """Methods scored from the [NC] raw class logits of every box (ood_utils.py:1183-1364) -- kernel K3."""

    _slot: int = -1              # OODB200_LOGIT_* of the subclass

    def __init__(self, name: str, per_class: bool, per_stride: bool, iou_threshold_for_matching: float,
                 min_conf_threshold_train: float, min_conf_threshold_test: float, use_values_before_sigmoid: bool, **kwargs):
        super().__init__(name, False, per_class, per_stride, iou_threshold_for_matching, min_conf_threshold_train,
                         min_conf_threshold_test, 'logits', False)
        self.cluster_method = 'None'
        self.use_values_before_sigmoid = use_values_before_sigmoid
        self.min_score = None
        self.max_score = None

    def _method_mask(self) -> int:
        return 1 << self._slot

    # temperatures of the K3 launch (Energy / ODIN override)
    def _temperatures(self) -> Tuple[float, float]:
        return 1.0, 1000.0

    def _table(self, values, nc: int, device) -> Optional[Tensor]:
        """per-class python list -> float64 [5, nc] device table with this method's row filled."""
        if values is None:
            return None
        row = np.zeros(nc, dtype=np.float64)
        for c in range(min(nc, len(values))):
            v = values[c]
            row[c] = float(v) if not (isinstance(v, (list, tuple)) and len(v) == 0) else 0.0
        tab = np.zeros((ops.N_LOGIT, nc), dtype=np.float64)
        tab[self._slot] = row
        return ops.h2d(tab, device)

    def _gather(self, results):
        dev = ops.default_device()
        counts = [int(len(res.boxes.cls)) for res in results]
        if sum(counts) == 0:
            return dev, counts, None, None
        rows = [(res.extra_item, res.boxes.cls) for res, m in zip(results, counts) if m]
        if all(isinstance(z, torch.Tensor) and not z.is_cuda and not c.is_cuda for z, c in rows):   # host inputs: one copy each
            logits = ops.h2d(torch.cat([z.reshape(len(c), -1) for z, c in rows]), dev, torch.float32)
            cls = ops.h2d(torch.cat([c.reshape(-1) for _, c in rows]), dev, torch.int32)
        else:
            logits = torch.cat([_to_device_f32(z, dev).reshape(len(c), -1) for z, c in rows])
            cls = torch.cat([c.to(dev) for _, c in rows]).to(torch.int32)
        return dev, counts, logits, cls

    def _launch(self, logits: Tensor, cls: Tensor, with_tables: bool):
        nc = int(logits.shape[1])
        dev = logits.device
        te, to = self._temperatures()
        thr = self._table(self.thresholds, nc, dev) if with_tables else None
        smin = self._table(self.min_score, nc, dev) if with_tables and self.min_score is not None else None
        smax = self._table(self.max_score, nc, dev) if with_tables and self.max_score is not None else None
        out = ops.logit_score(logits, cls, self._method_mask(), t_energy=te, t_odin=to, thr=thr, smin=smin, smax=smax,
                              clip=CUSTOM_HYP.fusion.CLIP_FUSION_SCORES)
        return out

    def _scores_device(self, logits: Tensor, cls: Tensor) -> Tensor:
        return self._launch(logits, cls, False).scores[self._slot]

    def compute_ood_decision_on_results(self, results, logger: Logger) -> List[List[int]]:
        """0 if score < thresholds[cls] else 1, box order (ood_utils.py:1195-1208); one launch for the batch."""
        dev, counts, logits, cls = self._gather(results)
        if logits is None:
            return [[] for _ in counts]
        out = self._launch(logits, cls, True)
        self._post_launch_checks(out)
        return _split_lists(out.decision[self._slot].cpu().numpy(), counts, int)

    def compute_INDness_scores_on_results(self, results, logger: Logger) -> List[List[float]]:
        """Piecewise-linear INDness in [-1, 1] through (min_score, -1), (thr, 0), (max_score, +1)
        (ood_utils.py:1210-1257)."""
        if not CUSTOM_HYP.fusion.LOGITS_USE_PIECEWISE_FUNCTION:
            raise NotImplementedError("Not implemented yet")
        dev, counts, logits, cls = self._gather(results)
        if logits is None:
            return [[] for _ in counts]
        out = self._launch(logits, cls, True)
        self._post_launch_checks(out)
        return _split_lists(out.indness[self._slot].cpu().numpy(), counts, float)

    def compute_indness(self, score: float, cls_idx: int) -> float:
        """Scalar form of the INDness map (ood_utils.py:1224-1257); the batched path evaluates it inside K3."""
        if not CUSTOM_HYP.fusion.LOGITS_USE_PIECEWISE_FUNCTION:
            raise NotImplementedError("Not implemented yet")
        t = self.thresholds[cls_idx]
        if score > t:
            a, b = 1 / (self.max_score[cls_idx] - t), -t / (self.max_score[cls_idx] - t)
        elif score < t:
            a, b = -1 / (self.min_score[cls_idx] - t), t / (self.min_score[cls_idx] - t)
        else:
            a, b = 0, 0
        indness = a * score + b
        return max(-1, min(indness, 1)) if CUSTOM_HYP.fusion.CLIP_FUSION_SCORES else indness

    def _post_launch_checks(self, out) -> None:
        pass

    def compute_scores(self, logits, cls_idx) -> np.ndarray:
        """Scores of n boxes of class `cls_idx` (int, or an [n] vector): logits [n, NC] or [NC]."""
        dev = ops.default_device()
        z = _to_device_f32(logits, dev)
        if z.ndim == 1:
            z = z[None, :]
        if isinstance(cls_idx, (int, np.integer)):
            cls = torch.full((z.shape[0],), int(cls_idx), dtype=torch.int32, device=dev)
        else:
            cls = torch.as_tensor(np.asarray(_np(cls_idx)), dtype=torch.int32).to(dev).reshape(-1)
        out = self._launch(z, cls, False)
        self._post_launch_checks(out)
        return out.scores[self._slot].cpu().numpy()

    def extract_internal_activations(self, results, all_activations: list, targets):
        """Append the logit rows of the valid predictions to all_activations[cls] (ood_utils.py:1284-1297)."""
        for res in results:
            if not len(res.valid_preds):
                continue
            idx = torch.as_tensor(list(res.valid_preds), dtype=torch.long)
            cls = res.boxes.cls.detach().cpu()[idx].to(torch.int64).numpy()
            rows = res.extra_item.detach()[idx.to(res.extra_item.device)].cpu()
            for c in np.unique(cls):
                all_activations[int(c)].append(rows[torch.from_numpy(cls == c)])

    def format_internal_activations(self, all_activations: list):
        """all_activations[cls] -> one [N_cls, NC] tensor (ood_utils.py:1300-1309)."""
        for c in range(len(all_activations)):
            chunks = [v if v.ndim == 2 else v[None, :] for v in all_activations[c]] if isinstance(all_activations[c], list) else None
            if chunks is None:
                continue
            all_activations[c] = torch.cat(chunks, dim=0) if len(chunks) else torch.tensor([])

    def compute_scores_from_activations(self, activations: list, logger: Logger):
        """scores[cls] = this method's score of every InD logit row of the class, all classes in ONE launch;
        also records min_score / max_score (ood_utils.py:1311-1347)."""
        if not self.per_class:
            raise NotImplementedError("Not implemented yet")
        dev = ops.default_device()
        sizes = [int(len(a)) for a in activations]
        scores = [np.array([], dtype=np.float32) for _ in activations]
        if sum(sizes):
            z = torch.cat([_to_device_f32(a, dev).reshape(n, -1) for a, n in zip(activations, sizes) if n])
            cls = torch.repeat_interleave(torch.arange(len(sizes), dtype=torch.int32, device=dev),
                                          torch.tensor(sizes, device=dev))
            out = self._launch(z, cls, False)
            self._post_launch_checks(out)
            flat = out.scores[self._slot].cpu().numpy()
            pos = 0
            for c, n in enumerate(sizes):
                if n:
                    scores[c] = flat[pos:pos + n].copy()
                    pos += n
        self.obtain_min_max_distances(scores)
        return scores

    def obtain_min_max_distances(self, scores):
        if self.per_class:
            self.min_score = [float(np.min(s)) if len(s) > 0 else 0.0 for s in scores]
            self.max_score = [float(np.max(s)) if len(s) > 0 else 0.0 for s in scores]

    def activations_transformation(self, activations, **kwargs):
        return activations

    def compute_distance(self, centroids, features):
        raise NotImplementedError("This method is not needed for methods using logits")


class NoMethod(LogitsMethod):
    """Everything is InD (ood_utils.py:1366-1385)."""

    def __init__(self, **kwargs):
        super().__init__('No OoD method', **kwargs)

    def compute_scores(self, logits, cls_idx) -> np.ndarray:
        n = 1 if len(logits.shape) == 1 else int(logits.shape[0])
        return np.ones(n)

    def compute_scores_from_activations(self, activations, logger):
        scores = [np.ones(len(a)) if len(a) > 0 else np.array([], dtype=np.float32) for a in activations]
        self.obtain_min_max_distances(scores)
        return scores

    def compute_ood_decision_on_results(self, results, logger) -> List[List[int]]:
        return [[1] * int(len(res.boxes.cls)) for res in results]


class MSP(LogitsMethod):
    """softmax(logits)[cls] (ood_utils.py:1388-1397)."""
    _slot = ops.LOGIT_SLOT["MSP"]

    def __init__(self, **kwargs):
        super().__init__('MSP', **kwargs)


class Energy(LogitsMethod):
    """T * logsumexp(logits / T) (ood_utils.py:1400-1412)."""
    _slot = ops.LOGIT_SLOT["Energy"]

    def __init__(self, temper: float, **kwargs):
        super().__init__('Energy', **kwargs)
        self.temper = temper

    def _temperatures(self):
        return float(self.temper), 1000.0


class ODIN(LogitsMethod):
    """softmax(logits / T)[cls] (ood_utils.py:1415-1427)."""
    _slot = ops.LOGIT_SLOT["ODIN"]

    def __init__(self, temper: float, **kwargs):
        super().__init__('ODIN', **kwargs)
        self.temper = temper

    def _temperatures(self):
        return 1.0, float(self.temper)


class Sigmoid(LogitsMethod):
    """sigmoid(logits)[cls]; asserts that cls is the arg-max logit (ood_utils.py:1430-1443; its `name` is 'MSP'
    in the reference too)."""
    _slot = ops.LOGIT_SLOT["Sigmoid"]

    def __init__(self, **kwargs):
        super().__init__('MSP', **kwargs)

    def _method_mask(self) -> int:
        # the detector already applied the sigmoid: K3's Sigmoid slot then returns the input value itself
        return (1 << self._slot) | (0 if self.use_values_before_sigmoid else ops.LOGIT_FLAG_POST_SIGMOID)

    def _post_launch_checks(self, out) -> None:
        assert int(out.sigmoid_mismatch.item()) == 0, "The max logit is not the one of the predicted class"


class MaxLogit(LogitsMethod):
    """max_j logits[j].  NOT in the reference (BASELINE.json names it, the closest reference method is
    `Sigmoid`); defined here as `logits.max(1)`, parity unpinned."""
    _slot = ops.LOGIT_SLOT["MaxLogit"]

    def __init__(self, **kwargs):
        super().__init__('MaxLogit', **kwargs)


# ------------------------------------------------------------------------------------------------ distance family
class DistanceMethod(OODMethod):
    """Feature-map methods: RoI-pooled vector of every box vs the centroids of its (class, stride)
    (ood_utils.py:1447-2410) -- kernels K1+K2 (scoring), K4 (k-means), K2 standalone + K5 (fit scores, thresholds)."""

    metric: str
    normalize_activations: bool = True     # vanilla FMap methods L2-normalise the pooled vector (ood_utils.py:2409)
    activations_on_device: bool = False    # keep collected InD activations as CUDA tensors (large fits)
    fit_reduce: str = "allreduce"          # sharded fit: "ordered" = rank-count-invariant reductions (identical bits for 1/2/4/8 ranks)

    def __init__(self, name: str, per_class: bool, per_stride: bool, cluster_method: str, metric: str,
                 cluster_optimization_metric: str, agg_method: str, ind_info_creation_option: str,
                 which_internal_activations: str, **kwargs):
        which_internal_activations = self.validate_correct_which_internal_activations_distance_methods(which_internal_activations)
        self._clusters = None
        self._packed = None
        super().__init__(name, True, per_class, per_stride, which_internal_activations=which_internal_activations, **kwargs)
        self.metric = metric
        self.cluster_method = self.check_cluster_method_selected(cluster_method)
        self.cluster_optimization_metric = self.check_cluster_optimization_metric_selected(cluster_optimization_metric)
        self.agg_method = self.select_agg_method(agg_method)
        self._agg_name = agg_method
        self.ind_info_creation_option = self.validate_correct_ind_info_creation_option(ind_info_creation_option)
        self.min_dist = None
        self.max_dist = None

    # fitted state: callers assign `.clusters` / `.thresholds` directly (also from disk, ood_evaluation.py:471-476,
    # :529, :574), so the packed device tables are rebuilt lazily whenever these attributes change
    @property
    def clusters(self):
        return self._clusters

    @clusters.setter
    def clusters(self, value):
        self._clusters = value
        self._packed = None

    def invalidate_device_tables(self) -> None:
        """Call after mutating `.clusters` in place (re-assignment is detected automatically)."""
        self._packed = None

    # -- validators (ood_utils.py:1473-1493)
    def validate_correct_which_internal_activations_distance_methods(self, which_internal_activations: str) -> str:
        assert which_internal_activations in FTMAPS_RELATED_OPTIONS, \
            f"which_internal_activations must be one of {FTMAPS_RELATED_OPTIONS}, but got {which_internal_activations}"
        return which_internal_activations

    def validate_correct_ind_info_creation_option(self, ind_info_creation_option: str) -> str:
        assert ind_info_creation_option in IND_INFO_CREATION_OPTIONS, \
            f"ind_info_creation_option must be one of {IND_INFO_CREATION_OPTIONS}, but got {ind_info_creation_option}"
        return ind_info_creation_option

    def select_agg_method(self, agg_method: str) -> Callable:
        assert agg_method in ['mean', 'median'], f"agg_method must be one of ['mean', 'median'], but got {agg_method}"
        return np.mean if agg_method == 'mean' else np.median

    def check_cluster_method_selected(self, cluster_method: str) -> str:
        assert is_valid_cluster_method(cluster_method), f"cluster_method must be one of the available clustering methods, but got {cluster_method}"
        return cluster_method

    def check_cluster_optimization_metric_selected(self, cluster_optimization_metric: str) -> str:
        assert cluster_optimization_metric in AVAILABLE_CLUSTER_OPTIMIZATION_METRICS, \
            f"cluster_method must be one of {AVAILABLE_CLUSTER_OPTIMIZATION_METRICS}, but got {cluster_optimization_metric}"
        return cluster_optimization_metric

    # -- device tables
    @property
    def _metric_slot(self) -> int:
        return ops.METRIC_SLOT[self.metric]

    def _device_table(self, dims: Sequence[int], device) -> ops.CentroidTable:
        """Packed centroids (cached until `.clusters` is re-assigned) + thresholds (re-read every call: 3*NC floats,
        so in-place edits of `.thresholds` are honoured)."""
        if self._clusters is None:
            raise RuntimeError(f"{self.name}: `.clusters` is not set (fit or load the clusters first)")
        key = (tuple(int(d) for d in dims), str(device))
        if self._packed is None or self._packed[0] != key:
            table = ops.pack_centroids(self._clusters, {}, list(dims), device)
            self._packed = (key, table, None)
        _, table, thr_host = self._packed
        thr = ops.pack_thresholds({self._metric_slot: self.thresholds}, table.nc) if self.thresholds is not None \
            else np.full((3, 3 * table.nc), np.nan)
        if thr_host is None or not np.array_equal(thr, thr_host, equal_nan=True):
            table.thr = torch.from_numpy(thr).to(device)
            self._packed = (key, table, thr)
        return table

    # -- enhanced unknown localisation: ranking of the unknown proposals (ood_utils.py:1031-1084)
    def rank_unknown_proposals(self, feature_map, proposals, selected_stride: int, operation: Optional[str] = None):
        """Rank of every unknown-object proposal of one image: RoIAlign (1x1, spatial_scale 1, aligned=False) of the
        proposals on the (padded) feature map of the selected stride, the transformed vector's distance to the nearest
        centroid of EVERY known class that has clusters on that stride, folded over the classes with
        CUSTOM_HYP.unk.rank.RANK_BOXES_OPERATION ('mean' / 'max' / 'sum' / 'min' (x100, or as is with the index of the
        closest class when USE_OOD_THR_TO_REMOVE_PROPS) / 'geometric_mean' / 'entropy').
        feature_map: [C, H, W] tensor; proposals: [P, 4] xyxy in feature-map cells.  Pooling (K1) and the
        [classes, P] distance matrix (K2, one launch, one segment per class) run on the GPU; the fold over <= NC values
        per proposal is the reference's own numpy / scipy expression.
        -> ranks [P] (and, for 'min' with USE_OOD_THR_TO_REMOVE_PROPS, the index of the closest class among the classes
        with clusters, as `(ranks, idx_of_closest_cluster)`)."""
        if self._clusters is None:
            raise RuntimeError(f"{self.name}: `.clusters` is not set (fit or load the clusters first)")
        op = operation or CUSTOM_HYP.unk.rank.RANK_BOXES_OPERATION
        dev = ops.default_device()
        fm = feature_map if isinstance(feature_map, torch.Tensor) else torch.as_tensor(np.asarray(feature_map))
        assert fm.dim() == 3, "feature_map must be [C, H, W]"
        boxes = torch.as_tensor(_np(proposals), dtype=torch.float32).reshape(-1, 4)
        n_prop, dim = int(boxes.shape[0]), int(fm.shape[0])
        if n_prop == 0:
            return np.zeros(0, dtype=np.float32)
        s = int(selected_stride)
        dummy = torch.zeros((1, 1, 1), dtype=torch.float32, device=dev)
        maps = [dummy, dummy, dummy]
        maps[s] = fm
        # spatial_scale = 1: the "image" is the feature map itself
        batch = ops.make_batch([maps], [boxes], [torch.full((n_prop,), float(s))], [torch.zeros(n_prop)], int(fm.shape[2]), dev)
        pooled = ops.roi_pool(batch)[:, :dim].contiguous()
        classes = [c for c, per_cls in enumerate(self._clusters) if s < len(per_cls) and len(per_cls[s]) > 0]
        if not classes:
            raise ValueError(f"no class has clusters on stride {s}")
        # one segment per class over the same vectors; the transformation is applied per class (ood_utils.py:1047-1052)
        xr = self._transform_device(pooled.repeat(len(classes), 1), classes, [n_prop] * len(classes), s).contiguous()
        d_eff = int(xr.shape[1])
        cents = [np.ascontiguousarray(_np(self._clusters[c][s]), dtype=np.float32).reshape(-1, d_eff) for c in classes]
        ks = [int(c.shape[0]) for c in cents]
        cent = np.concatenate(cents)
        row_off = np.concatenate([[0], np.cumsum(ks)])[:-1].tolist()
        cent_d = ops.h2d(cent, dev)
        unit_d = ops.h2d(ops._unit_rows(cent), dev) if self.metric == 'cosine' else None
        seg_off = [i * n_prop for i in range(len(classes) + 1)]
        dist, _ = ops.vec_score(xr, seg_off, cent_d, unit_d, row_off, ks, 1 << self._metric_slot, normalize=False)
        d = dist[self._metric_slot].reshape(len(classes), n_prop).cpu().numpy()
        return fold_proposal_distances(d, op, CUSTOM_HYP.unk.rank.USE_OOD_THR_TO_REMOVE_PROPS)

    # -- scoring primitives with the reference's numpy-in / numpy-out contracts
    def compute_scores(self, activations, cluster) -> np.ndarray:
        return self.compute_distance(cluster, activations)

    def activations_transformation(self, activations, **kwargs):
        """`sklearn.preprocessing.normalize(activations.reshape(N, -1), axis=1)` on the GPU (ood_utils.py:2404-2409).
        numpy in -> numpy out; CUDA tensor in -> CUDA tensor out."""
        x = _rows_2d(activations)
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return ops.normalize_rows(x.to(torch.float32))
        return ops.normalize_rows(_to_device_f32(x, ops.default_device())).cpu().numpy()

    def compute_distance(self, cluster, activations):
        """min over the K centroids of the pairwise distance (ood_utils.py:2422-2430): cluster [K, D],
        activations [n, D] (already transformed) -> [n] float32.  (The reference returns float64 for 'l1' because
        scipy's cdist does; the values agree to float32 rounding.)"""
        dev = ops.default_device()
        x = _to_device_f32(_rows_2d(activations), dev)
        c = np.ascontiguousarray(_np(cluster), dtype=np.float32).reshape(-1, x.shape[1])
        cent = torch.from_numpy(c).to(dev)
        unit = torch.from_numpy(ops._unit_rows(c)).to(dev) if self.metric == 'cosine' else None
        dist, _ = ops.vec_score(x, [0, x.shape[0]], cent, unit, [0], [c.shape[0]], 1 << self._metric_slot, normalize=False)
        out = dist[self._metric_slot]
        return out if (isinstance(activations, torch.Tensor) and activations.is_cuda) else out.cpu().numpy()

    # -- decisions
    def _fused_scores(self, results):
        """One fused K1+K2 launch over every box of `results` ('ftmaps_and_strides')."""
        dev = ops.default_device()
        hw = _img_hw(results[0])
        assert all(_img_hw(r) == hw for r in results), "all images of a batch must share the network input shape"
        maps = [list(res.extra_item[0]) for res in results]
        strides = [res.extra_item[1] for res in results]
        batch = ops.make_batch(maps, [res.boxes.xyxy for res in results], strides, [res.boxes.cls for res in results],
                               hw[1], dev)
        dims = [int(c) for c in batch.map_chw.reshape(3, 3)[:, 0]]
        table = self._device_table(dims, dev)
        out = ops.fmap_score(batch, table, 1 << self._metric_slot, normalize=self.normalize_activations,
                             compat_q1=self.reference_compat, want_plan=True)
        return batch, table, out

    def _vectors_per_stride(self, results):
        """Per-box vectors of the modes that do not pool here: 'roi_aligned_ftmaps' (pre-pooled per-stride lists) and
        'ftmaps_and_strides_exact_pos' (the anchor cell of every box, ood_utils.py:2068-2117).
        -> per image: list of 3 (box_idx int64 host array, vectors [m, C_s] device tensor)."""
        dev = ops.default_device()
        out = []
        for res in results:
            per_stride = []
            if self.which_internal_activations == 'roi_aligned_ftmaps':
                for idx, feats in res.extra_item:
                    idx_h = _np(idx).astype(np.int64).reshape(-1)
                    v = _to_device_f32(feats, dev).reshape(len(idx_h), -1) if len(idx_h) else None
                    per_stride.append((idx_h, v))
            else:
                ftmaps, pos = res.extra_item
                flat = [_to_device_f32(ft, dev).reshape(ft.shape[0], -1) for ft in ftmaps]
                w = int(tuple(res.orig_img.shape)[-1])
                edges = np.cumsum([0] + [(w // s) ** 2 for s in (8, 16, 32)])
                pos_h = _np(pos).astype(np.int64).reshape(-1)
                if len(pos_h) and pos_h.max() >= edges[-1]:
                    raise ValueError(f"stride position {int(pos_h.max())} cannot be greater than stride index max {int(edges[-1])}")
                sidx = np.searchsorted(edges, pos_h, side='right') - 1
                for s in range(3):
                    sel = np.nonzero(sidx == s)[0]
                    v = flat[s].index_select(1, torch.from_numpy(pos_h[sel] - edges[s]).to(dev)).t().contiguous() if len(sel) else None
                    per_stride.append((sel, v))
            out.append(per_stride)
        return out

    embeds_vectors: bool = False           # SDR methods: `_transform_device` maps the pooled vector to another space

    def _pooled_per_stride(self, results):
        """'ftmaps_and_strides' for the methods that cannot use the fused pass (SDR: a learned reducer sits between pooling
        and scoring): pool every box on its stride (K1) and hand the vectors out in the per-image / per-stride form of
        `_vectors_per_stride` (box indices in box order per stride, like predict.py:78-88)."""
        dev = ops.default_device()
        hw = _img_hw(results[0])
        assert all(_img_hw(r) == hw for r in results), "all images of a batch must share the network input shape"
        batch = ops.make_batch([list(res.extra_item[0]) for res in results], [res.boxes.xyxy for res in results],
                               [res.extra_item[1] for res in results], [res.boxes.cls for res in results], hw[1], dev)
        dims = [int(c) for c in batch.map_chw.reshape(3, 3)[:, 0]]
        pooled = ops.roi_pool(batch) if batch.n else None
        st_h = batch.stride_idx.cpu().numpy().astype(np.int64) if batch.n else np.zeros(0, np.int64)
        start = np.concatenate([[0], np.cumsum(batch.counts)]).astype(np.int64)
        out = []
        for i in range(len(results)):
            st = st_h[start[i]:start[i + 1]]
            per_stride = []
            for s in range(3):
                idx = np.nonzero(st == s)[0]
                v = pooled.index_select(0, torch.from_numpy(start[i] + idx).to(dev))[:, :dims[s]].contiguous() if len(idx) else None
                per_stride.append((idx, v))
            out.append(per_stride)
        return out

    def _vector_scores(self, results, want_decision: bool = True, per_img=None):
        """Distances / decisions for pre-extracted vectors, K2 standalone, one launch per stride.
        Returns (counts, dist [n] numpy, decision [n] numpy, cls_used [n], stride [n]) in the reference's output order."""
        dev = ops.default_device()
        per_img = self._vectors_per_stride(results) if per_img is None else per_img
        counts = [int(len(res.boxes.cls)) for res in results]
        n = sum(sum(len(ix) for ix, _ in ps) for ps in per_img)
        dist = np.zeros(n, np.float32)
        dec = np.zeros(n, np.uint8)
        cls_used = np.zeros(n, np.int64)
        stride_of = np.zeros(n, np.int64)
        nc = len(self._clusters)
        m = self._metric_slot
        meta = [[] for _ in range(3)]       # per stride: (output position, class used) of every row, in source order
        srcs = [[] for _ in range(3)]       # per stride: vector tensors, concatenated in the same order
        pos0 = 0
        for res, ps in zip(results, per_img):
            cls_h = _np(res.boxes.cls).astype(np.int64).reshape(-1)
            k = 0
            for s, (idx_h, v) in enumerate(ps):
                if len(idx_h) == 0:
                    continue
                srcs[s].append(v)
                for j, b in enumerate(idx_h):
                    if self.reference_compat:                    # Q1: class by in-stride index, stride-major output
                        meta[s].append((pos0 + k, int(cls_h[j])))
                    else:
                        meta[s].append((pos0 + int(b), int(cls_h[b])))
                    k += 1
            pos0 += k
        for s in range(3):
            if not meta[s]:
                continue
            out_pos = np.array([t[0] for t in meta[s]], np.int64)
            cl_src = np.array([t[1] for t in meta[s]], np.int64)
            order = np.argsort(cl_src, kind="stable")            # rows grouped by class = K2 segments
            x = torch.cat(srcs[s]).index_select(0, torch.from_numpy(order).to(dev))
            cl = cl_src[order]
            seg_off = np.searchsorted(cl, np.arange(nc + 1)).tolist()
            if self.embeds_vectors:
                # the learned embedding of every (class used, stride) segment that has clusters; other rows score 1000
                live = [c for c in range(nc) if seg_off[c + 1] > seg_off[c] and len(self._clusters[c][s]) > 0]
                emb = None
                if live:
                    xin = torch.cat([x[seg_off[c]:seg_off[c + 1]] for c in live])
                    part = self._transform_device(xin, live, [seg_off[c + 1] - seg_off[c] for c in live], s)
                    emb = torch.zeros((x.shape[0], part.shape[1]), dtype=torch.float32, device=dev)
                    pos = 0
                    for c in live:
                        m_c = seg_off[c + 1] - seg_off[c]
                        emb[seg_off[c]:seg_off[c + 1]] = part[pos:pos + m_c]
                        pos += m_c
                x = emb if emb is not None else torch.zeros((x.shape[0], 4), dtype=torch.float32, device=dev)
            dim = int(x.shape[1])
            cent_rows, row_off, kk = [], [], []
            for c in range(nc):
                a = self._cluster_rows(c, s, dim)
                row_off.append(sum(kk))
                kk.append(a.shape[0])
                cent_rows.append(a)
            cent_np = np.concatenate(cent_rows) if sum(kk) else np.zeros((1, dim), np.float32)
            cent = torch.from_numpy(cent_np).to(dev)
            unit = torch.from_numpy(ops._unit_rows(cent_np)).to(dev) if self.metric == 'cosine' else None
            thr = np.full((3, nc), np.nan)
            if self.thresholds is not None:
                thr[m] = ops.pack_thresholds({m: self.thresholds}, nc)[m].reshape(3, nc)[s]
            d, _, de = ops.vec_score_one(x, seg_off, cent, unit, row_off, kk, m, normalize=self.normalize_activations,
                                         thr=torch.from_numpy(thr).to(dev))
            o = out_pos[order]
            dist[o], dec[o], cls_used[o], stride_of[o] = d[m].cpu().numpy(), de[m].cpu().numpy(), cl, s
        return counts, dist, dec, cls_used, stride_of

    def _cluster_rows(self, c: int, s: int, dim: int) -> np.ndarray:
        a = self._clusters[c][s] if s < len(self._clusters[c]) else []
        a = np.asarray(a, dtype=np.float32)
        return a.reshape(-1, dim) if a.size else np.zeros((0, dim), np.float32)

    def _warn_missing_clusters(self, cls_used: np.ndarray, stride_of: np.ndarray, logger) -> None:
        if logger is None or self._clusters is None:
            return
        missing = {}
        for c, s in zip(cls_used.tolist(), stride_of.tolist()):
            if 0 <= c < len(self._clusters) and 0 <= s < 3 and len(self._clusters[c][s]) == 0:
                missing[(c, s)] = missing.get((c, s), 0) + 1
        for (c, s), k in sorted(missing.items()):
            logger.warning(f'{k} boxes are viewed as OOD: there is no cluster for class {c} and stride {s}')

    def score_results(self, results):
        """Batched scoring with everything the kernels produce.  Returns a dict of numpy arrays in the output order
        of `compute_ood_decision_on_results` (flat over images): dist, argmin, decision, cls_used, stride, and
        `counts` (boxes per image)."""
        if len(results) == 0:
            return dict(counts=[], dist=np.zeros(0, np.float32), argmin=np.zeros(0, np.int32), decision=np.zeros(0, np.uint8),
                        cls_used=np.zeros(0, np.int64), stride=np.zeros(0, np.int64))
        if self.which_internal_activations == 'ftmaps_and_strides' and self.embeds_vectors:
            counts, dist, dec, cls_used, stride_of = self._vector_scores(results, per_img=self._pooled_per_stride(results))
            return dict(counts=counts, dist=dist, argmin=None, decision=dec, cls_used=cls_used, stride=stride_of)
        if self.which_internal_activations == 'ftmaps_and_strides':
            batch, table, out = self._fused_scores(results)
            m = self._metric_slot
            pos = out.out_index.cpu().numpy().astype(np.int64)
            cls_used = np.empty(batch.n, np.int64)
            stride_of = np.empty(batch.n, np.int64)
            cls_used[pos] = out.cls_used.cpu().numpy()
            stride_of[pos] = batch.stride_idx.cpu().numpy()
            return dict(counts=batch.counts, dist=out.dist[m].cpu().numpy(), argmin=out.argmin[m].cpu().numpy(),
                        decision=out.decision[m].cpu().numpy(), cls_used=cls_used, stride=stride_of)
        if self.which_internal_activations in ('roi_aligned_ftmaps', 'ftmaps_and_strides_exact_pos'):
            counts, dist, dec, cls_used, stride_of = self._vector_scores(results)
            return dict(counts=counts, dist=dist, argmin=None, decision=dec, cls_used=cls_used, stride=stride_of)
        raise ValueError(f"The method {self.which_internal_activations} is invalid implemented yet")

    def compute_ood_decision_on_results(self, results, logger) -> List[List[int]]:
        """1 (InD) iff a threshold exists for (class, stride) and distance < threshold; a missing cluster scores
        1000 (ood_utils.py:2038-2180).  The whole batch is one fused pool + normalise + distance-min + threshold
        launch; the output is per image in the reference's order (stride-major under reference_compat, Q1)."""
        r = self.score_results(results)
        self._warn_missing_clusters(r["cls_used"], r["stride"], logger)
        return _split_lists(r["decision"], r["counts"], int)

    def compute_INDness_scores_on_results(self, results, logger) -> List[List[float]]:
        """INDness in [-1, 1] of every box (ood_utils.py:1498-1620).  reference_compat: the reference's per-stride
        code path always returns -1 (Q2).  Otherwise: piecewise linear through (min_dist, +1), (thr, 0), (max_dist, -1)."""
        if self.which_internal_activations not in ('ftmaps_and_strides', 'roi_aligned_ftmaps'):
            raise ValueError(f"The method {self.which_internal_activations} is invalid implemented yet")
        if CUSTOM_HYP.fusion.DISTANCE_USE_FROM_ZERO_TO_THR or not CUSTOM_HYP.fusion.DISTANCE_USE_IN_DISTRIBUTION_TO_DEFINE_LIMITS:
            raise NotImplementedError("only DISTANCE_USE_IN_DISTRIBUTION_TO_DEFINE_LIMITS is supported")
        if not (self.per_class and self.per_stride):
            raise NotImplementedError("Not implemented yet")
        if self.reference_compat:
            return [[-1] * int(len(res.boxes.cls)) for res in results]
        r = self.score_results(results)
        nc = len(self._clusters)
        dev = ops.default_device()
        m = self._metric_slot
        thr = ops.pack_thresholds({m: self.thresholds}, nc)[m]                       # [3*nc], index s*nc + c
        tab = lambda v: np.array([[float(v[c][s]) if not isinstance(v[c][s], list) else 0.0 for c in range(nc)]
                                  for s in range(3)], np.float64).reshape(-1)
        slot = torch.from_numpy((r["stride"] * nc + r["cls_used"]).astype(np.int32)).to(dev)
        ind = ops.dist_indness(torch.from_numpy(r["dist"]).to(dev), slot, torch.from_numpy(thr).to(dev),
                               torch.from_numpy(tab(self.min_dist)).to(dev), torch.from_numpy(tab(self.max_dist)).to(dev),
                               clip=CUSTOM_HYP.fusion.CLIP_FUSION_SCORES)
        return _split_lists(ind.cpu().numpy(), r["counts"], float)

    def compute_indness(self, score: float, cls_idx: int, stride_idx: int) -> float:
        """Scalar INDness.  reference_compat reproduces Q2 (always -1 for per-class-per-stride thresholds)."""
        if self.reference_compat or not self.thresholds[cls_idx][stride_idx]:
            return -1
        t = self.thresholds[cls_idx][stride_idx]
        if score > t:
            a, b = -1 / (self.max_dist[cls_idx][stride_idx] - t), t / (self.max_dist[cls_idx][stride_idx] - t)
        elif score < t:
            a, b = 1 / (self.min_dist[cls_idx][stride_idx] - t), -t / (self.min_dist[cls_idx][stride_idx] - t)
        else:
            a, b = 0, 0
        indness = a * score + b
        return max(-1, min(indness, 1)) if CUSTOM_HYP.fusion.CLIP_FUSION_SCORES else indness

    # -- fit: activation collection (ood_utils.py:1655-1836)
    def _append_rows(self, all_activations, rows: Tensor, cls: np.ndarray, st: np.ndarray, dims) -> None:
        """rows [m, Cmax] device; scatter into all_activations[cls][stride] as [k, C_s, 1, 1] chunks."""
        host = None if self.activations_on_device else rows.cpu().numpy()
        for c in np.unique(cls):
            for s in np.unique(st[cls == c]):
                sel = np.nonzero((cls == c) & (st == s))[0]
                d = int(dims[int(s)])
                if self.activations_on_device:
                    chunk = rows.index_select(0, torch.from_numpy(sel).to(rows.device))[:, :d].reshape(len(sel), d, 1, 1)
                else:
                    chunk = host[sel][:, :d].reshape(len(sel), d, 1, 1)
                all_activations[int(c)][int(s)].append(chunk)

    def extract_internal_activations(self, results, all_activations, targets):
        """Pool (K1) every predicted box of the batch on its own stride and keep the vectors of the valid
        predictions, grouped by (predicted class, stride) -- 'valid_preds_one_stride' (ood_utils.py:1715-1780)."""
        if self.ind_info_creation_option in ('all_targets_one_stride', 'all_targets_all_strides'):
            raise NotImplementedError("the all_targets_* options do not run in the reference either (SURVEY.md Q6)")
        if self.ind_info_creation_option != 'valid_preds_one_stride':
            raise NotImplementedError("Not implemented yet")
        dev = ops.default_device()
        if self.which_internal_activations == 'ftmaps_and_strides':
            hw = _img_hw(results[0])
            maps = [list(res.extra_item[0]) for res in results]
            strides = [res.extra_item[1] for res in results]
            batch = ops.make_batch(maps, [res.boxes.xyxy for res in results], strides, [res.boxes.cls for res in results],
                                   hw[1], dev)
            if batch.n == 0:
                return
            pooled = ops.roi_pool(batch)
            start = np.concatenate([[0], np.cumsum(batch.counts)])
            keep = np.concatenate([start[i] + np.asarray(sorted(res.valid_preds), dtype=np.int64)
                                   for i, res in enumerate(results)] + [np.zeros(0, np.int64)])
            if not len(keep):
                return
            rows = pooled.index_select(0, torch.from_numpy(keep).to(dev))
            cls = batch.cls.cpu().numpy()[keep].astype(np.int64)
            st = batch.stride_idx.cpu().numpy()[keep].astype(np.int64)
            ok = (st >= 0) & (st <= 2)
            self._append_rows(all_activations, rows[torch.from_numpy(ok).to(dev)] if not ok.all() else rows, cls[ok], st[ok],
                              batch.map_chw.reshape(3, 3)[:, 0])
        elif self.which_internal_activations in ('roi_aligned_ftmaps', 'ftmaps_and_strides_exact_pos'):
            per_img = self._vectors_per_stride(results)
            for res, ps in zip(results, per_img):
                cls_h = _np(res.boxes.cls).astype(np.int64).reshape(-1)
                valid = set(int(v) for v in res.valid_preds)
                for s, (idx_h, v) in enumerate(ps):
                    sel = np.array([j for j, b in enumerate(idx_h) if int(b) in valid], dtype=np.int64)
                    if len(sel):
                        rows = v.index_select(0, torch.from_numpy(sel).to(dev))
                        self._append_rows(all_activations, rows, cls_h[idx_h[sel]], np.full(len(sel), s), [rows.shape[1]] * 3)
        else:
            raise NotImplementedError("The method to extract internal activations is not implemented yet")

    def format_internal_activations(self, all_activations):
        """all_activations[cls][stride]: list of chunks -> one [N, C, 1, 1] array (np.empty(0) when there is none),
        ood_utils.py:1838-1874.  Accepts the reference's per-box [C, 1, 1] entries as well."""
        for c, per_cls in enumerate(all_activations):
            for s, chunks in enumerate(per_cls):
                if not isinstance(chunks, list):
                    continue
                if len(chunks) == 0:
                    all_activations[c][s] = np.empty(0)
                    continue
                fix = [v[None] if v.ndim == 3 else v for v in chunks]
                if isinstance(fix[0], torch.Tensor):
                    all_activations[c][s] = torch.cat(fix, dim=0)
                else:
                    all_activations[c][s] = np.concatenate(fix, axis=0)

    # -- fit: segment packing shared by clusters / scores
    def _stride_segments(self, tensors, s: int, min_len: int, device, group=None):
        """Classes whose stride-s segment has more than `min_len` rows -> (classes, sizes, x [sum, D] device).
        Sharded fit (`group`): every class is listed on every rank, also with 0 local rows (a class without samples on
        a stride, or a rank whose shard of it is empty); the row length is agreed on across the ranks so that a rank
        without any row still takes part in the collectives with a [0, D] matrix."""
        classes, sizes, parts = [], [], []
        for c, per_cls in enumerate(tensors):
            a = per_cls[s] if s < len(per_cls) else []
            if len(a) > min_len:
                classes.append(c)
                sizes.append(int(len(a)))
                if len(a):
                    parts.append(_to_device_f32(_rows_2d(a), device))
        dim = int(parts[0].shape[1]) if parts else 0
        if group is not None:
            t = torch.tensor([dim], dtype=torch.int64, device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=group)
            dim = int(t.item())
        if parts:
            x = torch.cat(parts)
        else:
            x = torch.zeros((0, dim), dtype=torch.float32, device=device) if (classes and dim) else None
        return classes, sizes, x

    def _transform_device(self, x: Tensor, classes, sizes, stride_idx: int) -> Tensor:
        """Device form of `activations_transformation` for a stack of segments (overridden by the SDR methods)."""
        return ops.normalize_rows(x) if self.normalize_activations else x

    # -- fit: clusters (ood_utils.py:2263-2371)
    def generate_clusters(self, ind_tensors, logger: Logger, group=None):
        """clusters[cls][stride] = [K, C_s] float32 centroids (np.empty(0) when there are <= MIN_SAMPLES vectors).
        'one': mean of the normalised vectors; 'KMeans_<k>': k-means labels (sklearn-compatible seeding + Lloyd, K4,
        all classes of a stride in the same launches) then per-label member means; 'KMeans': the number of clusters of
        every segment searched over 2..14 by silhouette / Calinski-Harabasz score (K7); 'all': every vector.
        `group`: process group when ind_tensors holds this rank's row shard of every segment (k-means only)."""
        t1 = time.perf_counter()
        if not (self.per_class and self.per_stride):
            raise NotImplementedError("Not implemented yet")
        median = self._agg_name == 'median'                # np.median over the members (ood_utils.py:1481-1483, :2306, :2365)
        if median and group is not None:
            raise NotImplementedError("agg_method='median' needs every member row of a cluster on one rank: fit it unsharded")
        method = self.cluster_method
        k = kmeans_k(method)
        if method not in ('one', 'all', 'KMeans') and k is None:
            raise NotImplementedError(f"cluster_method '{method}' is a CPU-library clusterer outside the GPU hot path")
        if k is not None and k < 2:
            raise ValueError("The number of clusters must be greater than 1")
        dev = ops.default_device()
        clusters = [[np.empty(0) for _ in range(3)] for _ in range(len(ind_tensors))]
        min_samples = CUSTOM_HYP.clusters.MIN_SAMPLES
        for s in range(3):
            classes, sizes, x = self._stride_segments(ind_tensors, s, min_samples if group is None else -1, dev, group)
            if group is not None:
                classes, sizes, x = self._drop_small_global(classes, sizes, x, min_samples, group)
            if not classes or x is None:
                continue
            x = self._transform_device(x, classes, sizes, s)
            if method == 'KMeans' and group is not None:
                # the searched fit needs all pair distances of a segment on one GPU: the segments are dealt out to the ranks
                self._searched_kmeans_sharded(x, classes, sizes, s, clusters, group, logger, median)
                continue
            if method == 'all':
                host = x.cpu().numpy()
                off = np.concatenate([[0], np.cumsum(sizes)])
                for i, c in enumerate(classes):
                    clusters[c][s] = host[off[i]:off[i + 1]].copy()
                continue
            gsizes = self._global_sizes(sizes, dev, group)
            # fit_reduce="ordered": the rank-count-invariant reduction, ALSO in a single process (same bits for 1/2/4/8 ranks)
            ordered = dict(reduce="ordered", global_sizes=gsizes, rot=classes) if self.fit_reduce == "ordered" else {}
            agg = (lambda xx, ss, ll, kk, gg: _kmeans.member_medians(xx, ss, ll, kk)) if median else \
                (lambda xx, ss, ll, kk, gg: _kmeans.member_means(xx, ss, ll, kk, group=gg, **(ordered if gg is group else {})))
            if method == 'one':
                means, counts = agg(x, sizes, None, 1, group)
            elif method == 'KMeans':
                # number of clusters searched per segment (cluster_utils.py:75-80, :203-356): labels of the best k
                off = np.concatenate([[0], np.cumsum(sizes)])
                labels = [_cluster_utils.search_number_of_clusters(x[off[i]:off[i + 1]], self.metric,
                                                                   self.cluster_optimization_metric, logger)[0]
                          for i in range(len(classes))]
                kmax = max(CUSTOM_HYP.clusters.RANGE_OF_CLUSTERS)
                means, counts = agg(x, sizes, torch.cat(labels), kmax, None)
            else:
                world, rank = (torch.distributed.get_world_size(group), torch.distributed.get_rank(group)) if group is not None else (1, 0)
                res = self._kmeans_fit(x, sizes, gsizes, k, world, rank, group, classes)
                means, counts = agg(x, sizes, res.labels, k, group)
            means, counts = means.cpu().numpy(), counts.cpu().numpy()
            for i, c in enumerate(classes):
                present = counts[i] > 0                      # `sorted(set(labels))`: empty clusters have no centroid
                clusters[c][s] = means[i][present].copy()
                if sizes[i] < 50:
                    logger.warning(f'WARNING: Class {c:03}, Stride {s} -> Only {sizes[i]} samples')
        for c in range(min(20, len(ind_tensors))):
            for s in range(3):
                if len(clusters[c][s]) == 0:
                    logger.warning(f'SKIPPING Class {c:03}, Stride {s} -> NO SAMPLES')
        x_ = str(timedelta(seconds=time.perf_counter() - t1)).split(':')
        logger.info(f'Clusters generated in {x_[0]} Hours, {x_[1]} Minutes {x_[2]} Seconds')
        return clusters

    def _searched_kmeans_sharded(self, x, classes, sizes, s, clusters, group, logger, median):
        """cluster_method='KMeans' with row-sharded input: the rows of every segment are all-gathered (rank-major = the
        row order of `kmeans.shard_rows`), segment i is searched and aggregated on rank i % world exactly like the
        single-process fit, and the small per-segment centroid arrays are exchanged."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = x.device
        t = torch.tensor(sizes, dtype=torch.int64, device=dev)
        all_sizes = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(all_sizes, t, group=group)
        all_sizes = [[int(v) for v in a.cpu()] for a in all_sizes]
        max_local = max(sum(a) for a in all_sizes)
        pad = torch.zeros((max_local, x.shape[1]), dtype=torch.float32, device=dev)
        pad[:x.shape[0]] = x
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        mine = {}
        kmax = max(CUSTOM_HYP.clusters.RANGE_OF_CLUSTERS)
        for i, c in enumerate(classes):
            if i % world != rank:
                continue
            parts = []
            for r in _kmeans.ranks_in_row_order(c, world):   # rows of class c in row order (rotation key = class index)
                o = sum(all_sizes[r][:i])
                parts.append(bufs[r][o:o + all_sizes[r][i]])
            xi = torch.cat(parts).contiguous()
            lab = _cluster_utils.search_number_of_clusters(xi, self.metric, self.cluster_optimization_metric, logger)[0]
            fn = _kmeans.member_medians if median else _kmeans.member_means
            means, counts = fn(xi, [int(xi.shape[0])], lab, kmax)
            means, counts = means.cpu().numpy(), counts.cpu().numpy()
            mine[c] = means[0][counts[0] > 0].copy()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        for part in gathered:
            for c, arr in part.items():
                clusters[c][s] = arr

    @staticmethod
    def _global_sizes(sizes, dev, group):
        if group is None:
            return list(sizes)
        t = torch.tensor(sizes, dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(t, group=group)
        return [int(v) for v in t.cpu()]

    def _drop_small_global(self, classes, sizes, x, min_samples, group):
        """Sharded fit: a segment is kept when its GLOBAL row count exceeds MIN_SAMPLES."""
        dev = ops.default_device()
        g = self._global_sizes(sizes, dev, group)
        keep = [i for i, n in enumerate(g) if n > min_samples]
        if len(keep) == len(classes):
            return classes, sizes, x
        off = np.concatenate([[0], np.cumsum(sizes)])
        parts = [x[off[i]:off[i + 1]] for i in keep]
        return [classes[i] for i in keep], [sizes[i] for i in keep], (torch.cat(parts) if parts else None)

    def _kmeans_fit(self, x, sizes, gsizes, k, world, rank, group, classes=None):
        """Sharded fit: rank r must hold the rows `kmeans.shard_rows([n_c], world, r, rot=[c])` of class c (per stride):
        the rotation key of a segment is its CLASS index, so the split does not depend on which classes are dropped."""
        if world == 1:
            return _kmeans.kmeans_fit_predict_single(x, sizes, k, random_state=10, reduce=self.fit_reduce)
        return _kmeans.kmeans_fit_sharded(x, sizes, gsizes, k, world, rank, group, random_state=10, rot=classes,
                                          reduce=self.fit_reduce)

    def generate_one_cluster_per_class_and_stride(self, ind_tensors, clusters_per_class_and_stride, logger):
        """In-place form kept for API compatibility (ood_utils.py:2297-2314)."""
        saved, self.cluster_method = self.cluster_method, 'one'
        try:
            out = self.generate_clusters(ind_tensors, logger)
        finally:
            self.cluster_method = saved
        for c in range(len(out)):
            clusters_per_class_and_stride[c] = out[c]

    def generate_multiple_cluster_per_class_per_stride(self, ind_tensors, clusters_per_class_and_stride, logger):
        out = self.generate_clusters(ind_tensors, logger)
        for c in range(len(out)):
            clusters_per_class_and_stride[c] = out[c]

    # -- fit: scores of the InD activations (ood_utils.py:1877-1915, :2000-2036)
    def compute_scores_from_activations(self, activations, logger: Logger):
        """scores[cls][stride] = distance of every InD vector to its nearest centroid of (cls, stride); every class of
        a stride is scored in one K2 launch.  Records min_dist / max_dist."""
        if not (self.per_class and self.per_stride):
            raise NotImplementedError("Not implemented yet")
        dev = ops.default_device()
        nc = len(activations)
        scores = [[[] for _ in range(3)] for _ in range(nc)]
        m = self._metric_slot
        for s in range(3):
            for c in range(nc):
                a = activations[c][s] if s < len(activations[c]) else []
                if len(a) == 0:
                    if c < 20:
                        logger.warning(f'SKIPPING Class {c:03}, Stride {s} -> NO SAMPLES')
                    scores[c][s] = np.empty(0)
                elif len(a) < 50:
                    logger.warning(f'WARNING: Class {c:03}, Stride {s} -> Only {len(a)} samples')
            classes, sizes, x = self._stride_segments(activations, s, 0, dev)
            keep = [i for i, c in enumerate(classes) if len(self._clusters[c][s]) > 0]
            if not keep:
                continue
            if len(keep) != len(classes):
                off = np.concatenate([[0], np.cumsum(sizes)])
                x = torch.cat([x[off[i]:off[i + 1]] for i in keep])
                classes, sizes = [classes[i] for i in keep], [sizes[i] for i in keep]
            x = self._transform_device(x, classes, sizes, s)
            dim = int(x.shape[1])
            cents = [self._cluster_rows(c, s, dim) for c in classes]
            kk = [a.shape[0] for a in cents]
            row_off = np.concatenate([[0], np.cumsum(kk)])[:-1].tolist()
            cent_np = np.concatenate(cents)
            cent = torch.from_numpy(cent_np).to(dev)
            unit = torch.from_numpy(ops._unit_rows(cent_np)).to(dev) if self.metric == 'cosine' else None
            seg_off = np.concatenate([[0], np.cumsum(sizes)]).tolist()
            d, _ = ops.vec_score_one(x, seg_off, cent, unit, row_off, kk, m, normalize=False)
            d = d[m] if self.activations_on_device else d[m].cpu().numpy()
            for i, c in enumerate(classes):
                scores[c][s] = d[seg_off[i]:seg_off[i + 1]]
        self.obtain_min_max_distances(scores)
        return scores

    def obtain_min_max_distances(self, scores):
        if not (self.per_class and self.per_stride):
            raise NotImplementedError("Not implemented yet")
        mn = lambda v: float(v.min()) if len(v) > 0 else 0
        mx = lambda v: float(v.max()) if len(v) > 0 else 0
        self.min_dist = [[mn(v) for v in per_cls] for per_cls in scores]
        self.max_dist = [[mx(v) for v in per_cls] for per_cls in scores]

    def compute_scores_clusters_per_class_and_stride(self, activations, scores, logger):
        out = self.compute_scores_from_activations(activations, logger)
        for c in range(len(out)):
            scores[c] = out[c]

    def compute_scores_one_class_one_stride(self, clusters_one_cls_one_stride, ind_activations_one_cls_one_stride):
        scores = []
        if len(ind_activations_one_cls_one_stride) > 0:
            if len(clusters_one_cls_one_stride) > 0:
                scores = self.compute_distance(clusters_one_cls_one_stride, ind_activations_one_cls_one_stride)
            else:
                raise ValueError("The clusters must have at least one sample")
        return scores

    def compute_scores_from_activations_for_unk_proposals(self, activations, logger):
        raise NotImplementedError("enhanced unknown localisation: only the proposal ranking (rank_unknown_proposals, SURVEY.md §8f rank 4) is rebuilt; the saliency / region-proposal stage around it is CPU image processing outside the hot path")

    def generate_unk_prop_thr(self, scores, tpr) -> None:
        raise NotImplementedError("enhanced unknown localisation: only the proposal ranking (rank_unknown_proposals, SURVEY.md §8f rank 4) is rebuilt; the saliency / region-proposal stage around it is CPU image processing outside the hot path")


class _PairwiseDistanceClustersPerClassPerStride(DistanceMethod):
    def __init__(self, name: str, metric: str, **kwargs):
        AVAILABLE_PAIRWISE_METRICS = ['cosine', 'l1', 'l2', 'manhattan', 'euclidean']
        super().__init__(name=name, metric=metric, per_class=True, per_stride=True, **kwargs)
        assert self.per_class and self.per_stride, "This method is only compatible with per_class and per_stride"
        assert self.metric in AVAILABLE_PAIRWISE_METRICS, f"The metric must be one of {AVAILABLE_PAIRWISE_METRICS}. Current value: {self.metric}"


class L1DistanceOneClusterPerStride(_PairwiseDistanceClustersPerClassPerStride):
    def __init__(self, **kwargs):
        super().__init__('L1DistancePerStride', 'l1', **kwargs)


class L2DistanceOneClusterPerStride(_PairwiseDistanceClustersPerClassPerStride):
    def __init__(self, **kwargs):
        super().__init__('L2DistancePerStride', 'l2', **kwargs)


class CosineDistanceOneClusterPerStride(_PairwiseDistanceClustersPerClassPerStride):
    def __init__(self, **kwargs):
        super().__init__('CosineDistancePerStride', 'cosine', **kwargs)


class _DimensionalityReductionMethod(_PairwiseDistanceClustersPerClassPerStride):
    """SDR methods (ood_utils.py:2433-2571): the pooled vector is embedded by a trained reducer and scored WITHOUT L2
    normalisation of the embedding.  Training the reducer (ivis / umap: CPU / TensorFlow libraries) is outside the hot
    path: supply fitted reducers through `set_reducers` -- either one per stride (`reducers[stride]`, what the
    reference trains, :2486-2492, :2526-2536) or one per (class, stride) (`reducers[cls][stride]`); a reducer is a
    callable or an object with `.transform(ndarray [n, C]) -> ndarray [n, d]`.  Pooling (K1), the normalisation the
    ivis methods apply BEFORE the embedding (:2542-2548) and the scoring of the embedded vectors (K2) run on the GPU;
    the reducer itself runs wherever its library runs."""

    normalize_activations = False          # the embedded vectors are scored as they are
    normalize_before_reduce = False        # ivis: sklearn normalize() of the pooled vector before the embedding
    embeds_vectors = True

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.reducers = None
        self.is_dimensionality_reduction_trained = False

    def set_reducers(self, reducers) -> None:
        self.reducers = reducers
        self.is_dimensionality_reduction_trained = reducers is not None

    def train_dimensionality_reduction_module(self, activations, logger):
        raise NotImplementedError("training the SDR reducer (ivis / umap) is outside the GPU hot path; fit it with the "
                                  "reference tooling and pass it through set_reducers()")

    def generate_clusters(self, ind_tensors, logger: Logger, group=None):
        if not self.is_dimensionality_reduction_trained:           # ood_utils.py:2450-2456
            self.train_dimensionality_reduction_module(ind_tensors, logger)
            self.is_dimensionality_reduction_trained = True
        return super().generate_clusters(ind_tensors, logger, group=group)

    def _reducer(self, cls_idx: int, stride_idx: int):
        r = self.reducers
        if r is None:
            raise RuntimeError(f"{self.name}: no reducers set (see set_reducers)")
        per_stride = len(r) == 3 and not isinstance(r[0], (list, tuple))
        return r[stride_idx] if per_stride else r[cls_idx][stride_idx]

    def _reduce(self, a: np.ndarray, cls_idx: int, stride_idx: int) -> np.ndarray:
        r = self._reducer(cls_idx, stride_idx)
        out = r.transform(a) if hasattr(r, "transform") else r(a)
        return np.ascontiguousarray(out, dtype=np.float32)

    def activations_transformation(self, activations, cls_idx: int = None, stride_idx: int = None, **kwargs):
        a = _rows_2d(activations)
        if self.normalize_before_reduce:
            a = ops.normalize_rows(_to_device_f32(a, ops.default_device()))
        return self._reduce(_np(a).astype(np.float32), cls_idx, stride_idx)

    def _transform_device(self, x, classes, sizes, stride_idx):
        if self.normalize_before_reduce:
            x = ops.normalize_rows(x)
        host = x.cpu().numpy()
        off = np.concatenate([[0], np.cumsum(sizes)])
        parts = [self._reduce(host[off[i]:off[i + 1]], c, stride_idx) for i, c in enumerate(classes)]
        return torch.from_numpy(np.concatenate(parts)).to(x.device)

    def _fused_scores(self, results):
        raise NotImplementedError("SDR methods embed the pooled vector between pooling and scoring: score_results() pools "
                                  "(K1), embeds and scores (K2) instead of the fused pass")

    def score_vectors(self, vectors: np.ndarray, cls_idx: int, stride_idx: int) -> np.ndarray:
        """Distances of already-pooled vectors of one (class, stride): embed, then score on the GPU."""
        return self.compute_distance(self._clusters[cls_idx][stride_idx], self.activations_transformation(
            vectors, cls_idx=cls_idx, stride_idx=stride_idx))


class UmapMethod(_DimensionalityReductionMethod):
    """name / metric as in the reference (ood_utils.py:2476-2478)."""
    def __init__(self, **kwargs):
        kwargs.pop('metric', None)
        super().__init__(name='CosineDistancePerStride', metric='cosine', **kwargs)


class _IvisMethodPairwiseDistance(_DimensionalityReductionMethod):
    normalize_before_reduce = True         # ood_utils.py:2542-2548

    def __init__(self, metric, name, **kwargs):
        super().__init__(name=name, metric=metric, **kwargs)


class IvisMethodCosine(_IvisMethodPairwiseDistance):
    def __init__(self, **kwargs):
        super().__init__('cosine', 'IvisCosineDistancePerStride', **kwargs)


class IvisMethodL1(_IvisMethodPairwiseDistance):
    def __init__(self, **kwargs):
        super().__init__('manhattan', 'IvisL1DistancePerStride', **kwargs)


class IvisMethodL2(_IvisMethodPairwiseDistance):
    def __init__(self, **kwargs):
        super().__init__('euclidean', 'IvisL2DistancePerStride', **kwargs)


class ActivationsExtractor(DistanceMethod):
    """Collects pooled activations only (create_dataset_of_activations.py:15, ood_utils.py:2599-2617)."""

    def __init__(self, **kwargs):
        super().__init__(name='ActivationsExtractor', per_class=True, per_stride=True, metric='l2', **kwargs)

    def compute_distance(self, cluster, activations):
        raise NotImplementedError("Not implemented yet")

    def activations_transformation(self, activations, **kwargs):
        raise NotImplementedError("Not implemented yet")


# ------------------------------------------------------------------------------------------------ fusion
def _flatten_lists(d: List[list], dtype) -> Tuple[np.ndarray, List[int]]:
    counts = [len(v) for v in d]
    flat = np.fromiter((x for v in d for x in v), dtype=dtype, count=sum(counts))
    return flat, counts


class _FusionBase(OODMethod):
    def extract_internal_activations(self, results, all_activations, targets):
        pass

    def format_internal_activations(self, all_activations):
        pass

    def compute_scores(self, activations, *args, **kwargs):
        pass

    def activations_transformation(self, activations, **kwargs):
        raise NotImplementedError("This method is not going to be called directly")

    def compute_distance(self, centroids, features):
        raise NotImplementedError("This method is not going to be called directly")

    def _methods(self) -> list:
        raise NotImplementedError

    def iterate_data_to_extract_ind_activations(self, data_loader, model, device, logger: Logger):
        """One detector pass per sub-method, each with its own extra output (the reference's version raises a
        TypeError because of a wrong keyword; this one does what it intends)."""
        out = []
        for m in self._methods():
            configure_extra_output_of_the_model(model, m)
            out.append(m.iterate_data_to_extract_ind_activations(data_loader, model, device, logger))
        return tuple(out)

    def generate_thresholds(self, ind_scores: list, tpr: float, logger: Logger, group=None):
        return tuple(m.generate_thresholds(sc, tpr, logger, group=group) for m, sc in zip(self._methods(), ind_scores))

    def compute_scores_from_activations(self, activations, logger: Logger):
        return tuple(m.compute_scores_from_activations(a, logger) for m, a in zip(self._methods(), activations))

    def _decide_for_metrics(self, model, imgs, device, logger):
        """Detector + decision once per sub-method (each needs its own extra output), then the fusion rule
        (ood_utils.py:2966-3003)."""
        decisions, results = [], None
        for m in self._methods():
            configure_extra_output_of_the_model(model, m)
            res = model.predict(imgs, save=False, verbose=False, conf=self.min_conf_threshold_test, device=device)
            if self.fusion_strategy == 'score':
                decisions.append(m.compute_INDness_scores_on_results(res, logger))
            else:
                decisions.append(m.compute_ood_decision_on_results(res, logger))
            if results is not None:
                for a, b in zip(results, res):
                    assert torch.allclose(a.boxes.xyxy, b.boxes.xyxy) and torch.allclose(a.boxes.cls, b.boxes.cls) \
                        and torch.allclose(a.boxes.conf, b.boxes.conf), "Results are not the same for the fused methods"
            else:
                results = res
        return results, self.fuse_ood_decisions(*decisions)


class FusionMethod(_FusionBase):
    """Two methods fused position-wise with 'and' (max), 'or' (min) or 'score' (sum of INDness > 0),
    ood_utils.py:2760-2940 -- kernel K6."""

    def __init__(self, method1, method2, fusion_strategy: str, fusion_method_name: str, cluster_method: str, **kwargs):
        self.method1, self.method2 = method1, method2
        self.fusion_strategy = fusion_strategy
        is_distance_method = bool(method1.is_distance_method or method2.is_distance_method)
        super().__init__(name=fusion_method_name, per_class=True, per_stride=True, is_distance_method=is_distance_method,
                         which_internal_activations="none", **kwargs)
        self.cluster_method = cluster_method if is_distance_method else 'None'

    def _methods(self):
        return [self.method1, self.method2]

    @property
    def clusters(self):
        if self.method1.is_distance_method and self.method2.is_distance_method:
            return self.method1.clusters, self.method2.clusters
        if self.method1.is_distance_method:
            return self.method1.clusters
        if self.method2.is_distance_method:
            return self.method2.clusters
        raise ValueError("This should not be called if none of the methods is a distance method")

    @clusters.setter
    def clusters(self, clusters):
        if self.method1.is_distance_method and self.method2.is_distance_method:
            self.method1.clusters, self.method2.clusters = clusters[0], clusters[1]
        elif self.method1.is_distance_method:
            self.method1.clusters = clusters
        elif self.method2.is_distance_method:
            self.method2.clusters = clusters
        else:
            raise ValueError("At least one of the methods must be a distance method to set the clusters")

    @property
    def thresholds(self):
        return self.method1.thresholds, self.method2.thresholds

    @thresholds.setter
    def thresholds(self, thresholds):
        if thresholds is not None:
            if len(thresholds) != 2:
                raise ValueError("The thresholds must be a tuple with two elements, one for the logits and one for the distance")
            self.method1.thresholds, self.method2.thresholds = thresholds[0], thresholds[1]
        else:
            self.method1.thresholds = None
            self.method2.thresholds = None

    def generate_clusters(self, ind_tensors, logger: Logger):
        if self.method1.is_distance_method and self.method2.is_distance_method:
            return [self.method1.generate_clusters(ind_tensors[0], logger), self.method2.generate_clusters(ind_tensors[1], logger)]
        if self.method1.is_distance_method:
            return self.method1.generate_clusters(ind_tensors[0], logger)
        if self.method2.is_distance_method:
            return self.method2.generate_clusters(ind_tensors[1], logger)
        raise ValueError("Both methods must be distance methods to generate the clusters")

    def compute_ood_decision_on_results(self, results, logger, results2=None) -> List[List[int]]:
        """Extension (the reference leaves this a no-op and fuses inside its metrics loop): decide with both methods
        and fuse.  `results` must carry method1's extra item, `results2` (default: the same list) method2's."""
        results2 = results if results2 is None else results2
        if self.fusion_strategy == 'score':
            d1 = self.method1.compute_INDness_scores_on_results(results, logger)
            d2 = self.method2.compute_INDness_scores_on_results(results2, logger)
        else:
            d1 = self.method1.compute_ood_decision_on_results(results, logger)
            d2 = self.method2.compute_ood_decision_on_results(results2, logger)
        return self.fuse_ood_decisions(d1, d2)

    def fuse_ood_decisions(self, ood_decision1: List[list], ood_decision2: List[list]) -> List[List[int]]:
        """Position-wise fusion (ood_utils.py:2906-2940); 1 = InD."""
        if self.fusion_strategy not in ("and", "or", "score"):
            raise NotImplementedError("Not implemented yet")
        for a, b in zip(ood_decision1, ood_decision2):
            assert len(a) == len(b), "The number of bboxes is different"
        dev = ops.default_device()
        if self.fusion_strategy == "score":
            s1, counts = _flatten_lists(ood_decision1, np.float32)
            s2, _ = _flatten_lists(ood_decision2, np.float32)
            out = ops.fuse_scores(torch.from_numpy(s1).to(dev), torch.from_numpy(s2).to(dev))
        else:
            d1, counts = _flatten_lists(ood_decision1, np.uint8)
            d2, _ = _flatten_lists(ood_decision2, np.uint8)
            out = ops.fuse_decisions(torch.from_numpy(d1).to(dev), torch.from_numpy(d2).to(dev), self.fusion_strategy)
        return _split_lists(out.cpu().numpy(), counts, int)


class TripleFusionMethod(_FusionBase):
    """Majority vote of three methods (ood_utils.py:3092-3301)."""

    def __init__(self, method1, method2, method3, cluster_method: str, **kwargs):
        self.method1, self.method2, self.method3 = method1, method2, method3
        self.fusion_strategy = 'majority_voting'
        is_distance_method = bool(method1.is_distance_method or method2.is_distance_method or method3.is_distance_method)
        super().__init__(name=f'fusion-{method1.name}-{method2.name}_{method3.name}', per_class=True, per_stride=True,
                         is_distance_method=is_distance_method, which_internal_activations="none", **kwargs)
        self.cluster_method = cluster_method if is_distance_method else 'None'

    def _methods(self):
        return [self.method1, self.method2, self.method3]

    @property
    def clusters(self):
        d = [m.clusters for m in self._methods() if m.is_distance_method]
        if not d:
            raise ValueError("This should not be called if none of the methods is a distance method")
        return tuple(d) if len(d) > 1 else d[0]

    @clusters.setter
    def clusters(self, clusters):
        d = [m for m in self._methods() if m.is_distance_method]
        if not d:
            raise ValueError("At least one of the methods must be a distance method to set the clusters")
        if len(d) == 1:
            d[0].clusters = clusters
        else:
            for m, c in zip(d, clusters):
                m.clusters = c

    @property
    def thresholds(self):
        return self.method1.thresholds, self.method2.thresholds, self.method3.thresholds

    @thresholds.setter
    def thresholds(self, thresholds):
        if thresholds is not None:
            if len(thresholds) != 3:
                raise ValueError("The thresholds must be a tuple with three elements, one per method")
            for m, t in zip(self._methods(), thresholds):
                m.thresholds = t
        else:
            for m in self._methods():
                m.thresholds = None

    def generate_clusters(self, ind_tensors, logger: Logger):
        """Clusters of the distance sub-methods (ood_utils.py:3246-3271: with one distance method its activations are
        taken from its own position, with several from consecutive positions)."""
        d = [(i, m) for i, m in enumerate(self._methods()) if m.is_distance_method]
        if not d:
            raise ValueError("Both methods must be distance methods to generate the clusters")
        if len(d) == 1:
            i, m = d[0]
            return m.generate_clusters(ind_tensors[i], logger)
        return [m.generate_clusters(ind_tensors[j], logger) for j, (_, m) in enumerate(d)]

    def compute_ood_decision_on_results(self, results, logger, results2=None, results3=None) -> List[List[int]]:
        rs = [results, results if results2 is None else results2, results if results3 is None else results3]
        return self.fuse_ood_decisions(*[m.compute_ood_decision_on_results(r, logger) for m, r in zip(self._methods(), rs)])

    def fuse_ood_decisions(self, ood_decision1, ood_decision2, ood_decision3) -> List[List[int]]:
        if self.fusion_strategy != 'majority_voting':
            raise ValueError("Only valid majority_voting fusion strategy")
        for a, b, c in zip(ood_decision1, ood_decision2, ood_decision3):
            assert len(a) == len(b) == len(c), "The number of bboxes is different"
        dev = ops.default_device()
        d1, counts = _flatten_lists(ood_decision1, np.uint8)
        d2, _ = _flatten_lists(ood_decision2, np.uint8)
        d3, _ = _flatten_lists(ood_decision3, np.uint8)
        t = lambda a: torch.from_numpy(a).to(dev)
        out = ops.fuse_decisions(t(d1), t(d2), 'majority_voting', t(d3))
        return _split_lists(out.cpu().numpy(), counts, int)


def fold_proposal_distances(distances_per_proposal: np.ndarray, operation: str, use_ood_thr_to_remove_props: bool = False):
    """[classes, P] distances -> [P] ranks, the reference's expressions (ood_utils.py:1057-1084)."""
    d = np.asarray(distances_per_proposal)
    if operation == 'mean':
        return d.mean(axis=0)
    if operation == 'max':
        return d.max(axis=0)
    if operation == 'sum':
        return d.sum(axis=0)
    if operation == 'min':
        if use_ood_thr_to_remove_props:
            return d.min(axis=0), np.argsort(d, axis=0)[0]
        return d.min(axis=0) * 100                       # "to compensate the low values"
    if operation == 'geometric_mean':
        from scipy.stats import gmean
        return gmean(d, axis=0)
    if operation == 'entropy':
        from scipy.stats import entropy
        return entropy(d / d.sum(axis=0), axis=0)
    raise NotImplementedError("This operation is not implemented yet")


# ------------------------------------------------------------------------------------------------ fused multi-method pass
def compute_ood_decisions_fused(methods: Sequence[OODMethod], results, logger, logits_results=None) -> Dict[str, List[List[int]]]:
    """Decisions of several methods on the same detections with the feature maps uploaded / gathered ONCE.

    Extension of the reference surface (which scores one method per detector pass, ood_evaluation.py:183-278):
    every `DistanceMethod` in `methods` that shares its `.clusters` object with the others is scored in the same fused
    launch (metric mask), the logit methods in one logit launch.  `results` carry `(ftmaps, strides)` extra items,
    `logits_results` (default: `results`) the raw class logits.  Returns {method.name or 'name#i': decisions}, each
    exactly what `method.compute_ood_decision_on_results` returns."""
    out: Dict[str, List[List[int]]] = {}
    dist_m = [m for m in methods if isinstance(m, DistanceMethod)]
    logit_m = [m for m in methods if isinstance(m, LogitsMethod)]
    key = lambda m, i: m.name if m.name not in out else f"{m.name}#{i}"
    # Order of work: (1) queue the upload of the feature maps, (2) flatten the detections / logits and launch every kernel
    # while that copy runs, (3) read all decisions back, (4) build the python lists.  Nothing waits on the device before (3).
    pending = []                                       # (result key, device tensor row, counts)
    share = []
    if dist_m:
        lead = dist_m[0]
        share = [m for m in dist_m if m.clusters is lead.clusters and m.which_internal_activations == 'ftmaps_and_strides'
                 and m.reference_compat == lead.reference_compat and m.normalize_activations == lead.normalize_activations]
        if len(share) > 1 and len(results):
            dev = ops.default_device()
            staged = ops.stage_maps([list(r.extra_item[0]) for r in results], len(results), dev)
            hw = _img_hw(results[0])
            batch = ops.make_batch(staged, [r.boxes.xyxy for r in results], [r.extra_item[1] for r in results],
                                   [r.boxes.cls for r in results], hw[1], dev)
            dims = [int(c) for c in batch.map_chw.reshape(3, 3)[:, 0]]
            table = ops.pack_centroids_cached(lead.clusters, {m._metric_slot: m.thresholds for m in share}, dims, dev)
            mask = 0
            for m in share:
                mask |= 1 << m._metric_slot
            res = ops.fmap_score(batch, table, mask, normalize=lead.normalize_activations, compat_q1=lead.reference_compat)
            for i, m in enumerate(share):
                k = key(m, i)
                out[k] = None                          # keeps the reference's order of keys
                pending.append((k, res.decision, m._metric_slot, batch.counts))
        else:
            share = []
    lres = results if logits_results is None else logits_results
    # logit methods that read the same kind of logits: ONE upload, ONE K3 launch (method mask), one read-back
    slots = [m._slot for m in logit_m]
    fusable = len(logit_m) > 1 and len(set(slots)) == len(slots) and all(s >= 0 for s in slots) \
        and len({m.use_values_before_sigmoid for m in logit_m}) == 1
    lres_out = None
    if fusable:
        lead_l = logit_m[0]
        dev, counts, logits, cls = lead_l._gather(lres)
        if logits is not None:
            nc = int(logits.shape[1])
            rows = lambda attr: [(m._slot, getattr(m, attr)) for m in logit_m if getattr(m, attr) is not None]
            thr = _logit_table(rows('thresholds'), nc, dev)
            with_ind = all(m.min_score is not None and m.max_score is not None for m in logit_m)
            smin = _logit_table(rows('min_score'), nc, dev) if with_ind else None
            smax = _logit_table(rows('max_score'), nc, dev) if with_ind else None
            te, to, mask = 1.0, 1000.0, 0
            for m in logit_m:
                mte, mto = m._temperatures()
                te = mte if m._slot == ops.LOGIT_SLOT['Energy'] else te
                to = mto if m._slot == ops.LOGIT_SLOT['ODIN'] else to
                mask |= m._method_mask()           # carries Sigmoid's post-sigmoid flag (use_values_before_sigmoid=False)
            lres_out = ops.logit_score(logits, cls, mask, t_energy=te, t_odin=to, thr=thr, smin=smin, smax=smax,
                                       clip=CUSTOM_HYP.fusion.CLIP_FUSION_SCORES)
    # (3) + (4)
    host = {}
    for k, dec, row, counts_k in pending:
        if id(dec) not in host:
            host[id(dec)] = dec.cpu().numpy()
        out[k] = _split_lists(host[id(dec)][row], counts_k, int)
    for i, m in enumerate(dist_m):
        if m not in share:
            out[key(m, i)] = m.compute_ood_decision_on_results(results, logger)
    if fusable:
        if lres_out is None:
            for i, m in enumerate(logit_m):
                out[key(m, i)] = [[] for _ in counts]
        else:
            for m in logit_m:
                m._post_launch_checks(lres_out)
            dec = lres_out.decision.cpu().numpy()
            for i, m in enumerate(logit_m):
                out[key(m, i)] = _split_lists(dec[m._slot], counts, int)
    else:
        for i, m in enumerate(logit_m):
            out[key(m, i)] = m.compute_ood_decision_on_results(lres, logger)
    return out


def _logit_table(rows, nc: int, device) -> Tensor:
    """[(slot, per-class python list)] -> float64 [N_LOGIT, nc] device table (LogitsMethod._table for several methods)."""
    tab = np.zeros((ops.N_LOGIT, nc), dtype=np.float64)
    for slot, values in rows:
        for c in range(min(nc, len(values))):
            v = values[c]
            tab[slot, c] = float(v) if not (isinstance(v, (list, tuple)) and len(v) == 0) else 0.0
    return ops.h2d(tab, device)


# ------------------------------------------------------------------------------------------------ detector hook-up
def configure_extra_output_of_the_model(model, ood_method):
    """Tell the (reference-patched ultralytics) detector which extra item to attach to its Results
    (ood_utils.py:3523-3543).  Pure attribute plumbing on `model.model`."""
    model.model.model[-1].output_values_before_sigmoid = False
    if ood_method.which_internal_activations in FTMAPS_RELATED_OPTIONS:
        model.model.which_layers_to_extract = "convolutional_layers"
    elif ood_method.which_internal_activations in LOGITS_RELATED_OPTIONS:
        model.model.which_layers_to_extract = "logits"
        if ood_method.use_values_before_sigmoid:
            model.model.model[-1].output_values_before_sigmoid = True
    elif ood_method.which_internal_activations == "none":
        model.model.which_layers_to_extract = "none"
    else:
        raise ValueError(f"The option {ood_method.which_internal_activations} is not valid.")
    model.model.extraction_mode = ood_method.which_internal_activations
    if "yolov10" in str(getattr(model, "ckpt_path", "")):
        model.model.model[23].validating = False
