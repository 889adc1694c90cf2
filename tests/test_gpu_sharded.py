"""SURVEY.md section 4 test plan (iii) on hardware: the N-rank fit equals the single-GPU fit.

Two ranks run the REAL CUDA kernels on row shards of every segment -- over NCCL when the box has two GPUs, else both on
cuda:0 over gloo (gloo moves CUDA tensors for all-reduce / broadcast only, so the worker stages all-gathers through the
host; the product code is unchanged) -- and the parent compares with the single-process fit on the same data:
  * k-means labels: bit-exact for both reductions; with reduce="ordered" centres, member means, fit scores and the exact
    percentile thresholds are bit-identical to the single-GPU run;
  * the class surface: generate_clusters(group=...) with empty (class, stride) cells, a class without samples on a
    stride and ragged shards, then generate_thresholds(group=...), equal the single-process result.
"""
from __future__ import annotations

import logging
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ood_in_object_detection_b200 import synth

pytestmark = pytest.mark.gpu
LOG = logging.getLogger("test")
LOG.setLevel(logging.ERROR)

SIZES = [70000, 17000, 5, 0, 33000]        # several super-blocks, a tiny segment, an empty one (>= 65 536 rows: device seeding)
DIM, K = 128, 4                             # D = 128: the tcgen05 Lloyd step takes it
NC = 5
CLS_SIZES = {0: [9000, 0, 5000], 1: [0, 0, 0], 2: [4100, 30, 0], 3: [12, 7000, 2], 4: [8200, 4097, 4096]}   # [class][stride]
CLS_DIMS = (128, 160, 192)
DIST_KW = dict(agg_method="mean", cluster_method="KMeans_4", cluster_optimization_metric="silhouette",
               ind_info_creation_option="valid_preds_one_stride", which_internal_activations="ftmaps_and_strides",
               iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _segments():
    return [synth.blob_vectors(10 + i, n, DIM, K, 7.0)[0] if n else np.zeros((0, DIM), np.float32) for i, n in enumerate(SIZES)]


def _class_activations():
    """[class][stride] -> [n, C_s, 1, 1] float32 (np.empty(0) when there is none), like format_internal_activations."""
    out = []
    for c in range(NC):
        row = []
        for s in range(3):
            n = CLS_SIZES[c][s]
            row.append(np.abs(synth.blob_vectors(100 + 10 * c + s, n, CLS_DIMS[s], 4, 7.0, unit_norm=False)[0]).reshape(n, -1, 1, 1)
                       if n else np.empty(0))
        out.append(row)
    return out


def _patch_gloo_all_gather():
    real = dist.all_gather

    def staged(tensor_list, tensor, group=None, async_op=False):
        if not tensor.is_cuda:
            return real(tensor_list, tensor, group=group, async_op=async_op)
        host = [torch.empty(t.shape, dtype=t.dtype) for t in tensor_list]
        real(host, tensor.cpu(), group=group)
        for d, h in zip(tensor_list, host):
            d.copy_(h)
    dist.all_gather = staged


def _fit_class_surface(group, world, rank):
    """fit (clusters -> scores -> thresholds) of the three distance methods on this rank's rows."""
    from ood_in_object_detection_b200 import kmeans, ood_utils
    acts = _class_activations()
    if group is not None:
        for c in range(NC):
            for s in range(3):
                n = CLS_SIZES[c][s]
                if n:
                    a, cnt = kmeans.shard_rows([n], world, rank, rot=[c])[0]
                    acts[c][s] = acts[c][s][a:a + cnt] if cnt else np.empty(0)
    out = {}
    for tag, cls in (("l1", ood_utils.L1DistanceOneClusterPerStride), ("cos", ood_utils.CosineDistanceOneClusterPerStride)):
        m = cls(**DIST_KW)
        m.fit_reduce = "ordered"
        clusters = m.generate_clusters(acts, LOG, group=group)
        m.clusters = clusters
        scores = m.compute_scores_from_activations(acts, LOG)
        thr = m.generate_thresholds(scores, 0.95, LOG, group=group)
        for c in range(NC):
            for s in range(3):
                out[f"{tag}_cl_{c}_{s}"] = np.asarray(clusters[c][s], np.float32)
                out[f"{tag}_thr_{c}_{s}"] = np.array(thr[c][s] if thr[c][s] != [] else np.nan, np.float64)
    return out


def _worker(rank: int, world: int, port: int, outdir: str, backend: str):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank if backend == "nccl" else 0)
    dev = torch.device("cuda", torch.cuda.current_device())
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        _patch_gloo_all_gather()
    try:
        from ood_in_object_detection_b200 import kmeans, ops, select
        segs = _segments()
        shard = kmeans.shard_rows(SIZES, world, rank)
        local = [s[a:a + n] for s, (a, n) in zip(segs, shard)]
        lsizes = [len(v) for v in local]
        x = torch.from_numpy(np.concatenate(local).reshape(-1, DIM)).to(dev)
        res = {"rows": np.array(x.shape[0])}
        g = dist.group.WORLD
        for mode in ("allreduce", "ordered"):
            r = kmeans.kmeans_fit_sharded(x, lsizes, SIZES, K, world, rank, group=g, reduce=mode)
            res[f"labels_{mode}"] = r.labels.cpu().numpy()
            res[f"centers_{mode}"] = r.centers.cpu().numpy()
            res[f"n_iter_{mode}"] = np.array(r.n_iter)
            assert r.seconds["seeding"] == "device"
        lab = torch.from_numpy(res["labels_ordered"]).to(dev)
        means, counts = kmeans.member_means(x, lsizes, lab, K, group=g, reduce="ordered", global_sizes=SIZES)
        res["means"], res["counts"] = means.cpu().numpy(), counts.cpu().numpy()
        off = np.concatenate([[0], np.cumsum(lsizes)]).tolist()
        d, _ = ops.vec_score_one(x, off, means.reshape(-1, DIM).contiguous(), None, [i * K for i in range(len(SIZES))],
                                 [K] * len(SIZES), ops.METRIC_SLOT["l2"], normalize=False)
        ranks = [select.lower_index(n, 95.0) if n > 10 else None for n in SIZES]
        thr, mn, mx = select.segment_select(d[ops.METRIC_SLOT["l2"]].contiguous(), off, ranks, group=g)
        res["thr"] = np.array([np.nan if v is None else v for v in thr])
        res.update(_fit_class_surface(g, world, rank))
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_rank_run():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d, backend), nprocs=2, join=True)
        yield [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(2)], backend


@pytest.fixture(scope="module")
def single_run():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from ood_in_object_detection_b200 import kmeans, ops, select
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(np.concatenate(_segments())).to(dev)
    out = {}
    for mode in ("allreduce", "ordered"):
        r = kmeans.kmeans_fit_predict_single(x, SIZES, K, reduce=mode)
        out[f"labels_{mode}"], out[f"centers_{mode}"], out[f"n_iter_{mode}"] = r.labels.cpu().numpy(), r.centers.cpu().numpy(), r.n_iter
    lab = torch.from_numpy(out["labels_ordered"]).to(dev)
    means, counts = kmeans.member_means(x, SIZES, lab, K, reduce="ordered")
    out["means"], out["counts"] = means.cpu().numpy(), counts.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(SIZES)]).tolist()
    d, _ = ops.vec_score_one(x, off, means.reshape(-1, DIM).contiguous(), None, [i * K for i in range(len(SIZES))],
                             [K] * len(SIZES), ops.METRIC_SLOT["l2"], normalize=False)
    ranks = [select.lower_index(n, 95.0) if n > 10 else None for n in SIZES]
    thr, _, _ = select.segment_select(d[ops.METRIC_SLOT["l2"]].contiguous(), off, ranks)
    out["thr"] = np.array([np.nan if v is None else v for v in thr])
    out.update(_fit_class_surface(None, 1, 0))
    return out


def _global_labels(run, key):
    from ood_in_object_detection_b200 import kmeans
    parts = []
    for g in range(len(SIZES)):
        for r in kmeans.ranks_in_row_order(g, 2):
            sh = kmeans.shard_rows(SIZES, 2, r)
            loc_off = np.concatenate([[0], np.cumsum([c for _, c in sh])])
            parts.append(run[r][key][loc_off[g]:loc_off[g] + sh[g][1]])
    return np.concatenate(parts)


def test_two_rank_kmeans_equals_single_gpu(two_rank_run, single_run):
    run, backend = two_rank_run
    assert int(run[0]["rows"]) + int(run[1]["rows"]) == sum(SIZES) and min(int(run[0]["rows"]), int(run[1]["rows"])) > 0
    for mode in ("allreduce", "ordered"):
        assert np.array_equal(_global_labels(run, f"labels_{mode}"), single_run[f"labels_{mode}"]), (mode, backend)
        assert np.array_equal(run[0][f"centers_{mode}"], run[1][f"centers_{mode}"])
        assert list(run[0][f"n_iter_{mode}"]) == list(single_run[f"n_iter_{mode}"])
    assert np.array_equal(single_run["labels_allreduce"], single_run["labels_ordered"])
    # the ordered reduction does not depend on the number of ranks: identical bits
    assert np.array_equal(run[0]["centers_ordered"], single_run["centers_ordered"])
    np.testing.assert_allclose(run[0]["centers_allreduce"], single_run["centers_allreduce"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(run[0]["means"], single_run["means"]) and np.array_equal(run[0]["counts"], single_run["counts"])
    assert np.array_equal(run[0]["thr"], single_run["thr"], equal_nan=True)
    assert np.array_equal(run[0]["thr"], run[1]["thr"], equal_nan=True)


def test_two_rank_class_surface_fit_equals_single_gpu(two_rank_run, single_run):
    """generate_clusters(group=...) / generate_thresholds(group=...) with empty (class, stride) cells: same clusters
    and thresholds as the single-process fit (bit-identical under fit_reduce='ordered')."""
    run, _ = two_rank_run
    n_cl = 0
    for tag in ("l1", "cos"):
        for c in range(NC):
            for s in range(3):
                a, b = run[0][f"{tag}_cl_{c}_{s}"], single_run[f"{tag}_cl_{c}_{s}"]
                assert a.shape == b.shape, (tag, c, s)
                assert np.array_equal(a, b), (tag, c, s)
                assert np.array_equal(a, run[1][f"{tag}_cl_{c}_{s}"])
                n_cl += a.size > 0
                t0, t1 = run[0][f"{tag}_thr_{c}_{s}"], single_run[f"{tag}_thr_{c}_{s}"]
                assert np.array_equal(t0, t1, equal_nan=True), (tag, c, s, t0, t1)
    assert n_cl >= 12                                     # the populated cells really were fitted
    assert single_run["l1_cl_1_0"].size == 0 and single_run["l1_cl_3_2"].size == 0     # no samples / <= MIN_SAMPLES rows
