"""world_size-2 `gloo` runs of the sharded fit path on CPU (no GPU needed): row sharding of every segment, the one
all-reduce per Lloyd iteration, the rank-count-invariant ordered reduction, the per-pass histogram exchange of the
exact percentile, and sharded member means.  The device steps are replaced by tests/cpu_backend.py; what is under
test is the host control flow of ood_in_object_detection_b200/{kmeans,select}.py that runs unchanged on NCCL."""
from __future__ import annotations

import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ood_in_object_detection_b200 import kmeans, select, synth
from tests.cpu_backend import NumpyBackend

SIZES = [40000, 17000, 5, 0, 33000]        # ragged segments: several super-blocks, a tiny one, an empty one
DIM, K = 12, 4


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data():
    segs = [synth.blob_vectors(10 + i, n, DIM, K, 7.0)[0] if n else np.zeros((0, DIM), np.float32) for i, n in enumerate(SIZES)]
    return segs


def _scores():
    rng = np.random.default_rng(5)
    out = [rng.normal(0, 1, size=n).astype(np.float32) for n in (70000, 11, 0, 3000)]
    out[0][:5000] = out[0][0]              # ties
    return out


def _worker(rank: int, world: int, port: int, outdir: str):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        be = NumpyBackend()
        segs = _data()
        shard = kmeans.shard_rows(SIZES, world, rank)
        local = [s[a:a + n] for s, (a, n) in zip(segs, shard)]
        x = torch.from_numpy(np.concatenate(local))
        res = {}
        for mode in ("allreduce", "ordered"):
            r = kmeans.kmeans_fit_sharded(x, [len(v) for v in local], SIZES, K, world, rank, group=dist.group.WORLD,
                                          backend=be, reduce=mode)
            res[f"labels_{mode}"] = r.labels.numpy()
            res[f"centers_{mode}"] = r.centers.numpy()
            res[f"n_iter_{mode}"] = np.array(r.n_iter)
        means, counts = kmeans.member_means(x, [len(v) for v in local], torch.from_numpy(res["labels_ordered"]), K,
                                            group=dist.group.WORLD, backend=be)
        res["means"], res["counts"] = means.numpy(), counts.numpy()
        # exact percentile over sharded scores: rank r holds every world-th score of each segment
        sc = _scores()
        mine = [v[rank::world] for v in sc]
        off = np.concatenate([[0], np.cumsum([len(v) for v in mine])]).tolist()
        ranks = [select.lower_index(len(v), 95.0) if len(v) > 5 else None for v in sc]
        vals, mn, mx = select.segment_select(torch.from_numpy(np.concatenate(mine)), off, ranks, group=dist.group.WORLD, backend=be)
        res["sel"] = np.array([np.nan if v is None else v for v in vals])
        res["mn"] = np.array([np.nan if v is None else v for v in mn])
        res["mx"] = np.array([np.nan if v is None else v for v in mx])
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_rank_run():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d), nprocs=2, join=True)
        yield [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(2)]


@pytest.fixture(scope="module")
def single_run():
    be = NumpyBackend()
    x = torch.from_numpy(np.concatenate(_data()))
    return kmeans.kmeans_fit_predict_single(x, SIZES, K, backend=be), kmeans.kmeans_fit_predict_single(x, SIZES, K, backend=be, reduce="ordered")


def test_sharded_kmeans_equals_single_process(two_rank_run, single_run):
    single_run, single_ordered = single_run
    assert np.array_equal(single_ordered.labels.numpy(), single_run.labels.numpy())
    labels1 = single_run.labels.numpy()
    off = np.concatenate([[0], np.cumsum(SIZES)])
    for mode in ("allreduce", "ordered"):
        parts = []
        for g, n in enumerate(SIZES):                       # re-assemble global row order from the two row shards
            for r in kmeans.ranks_in_row_order(g, 2):
                a, cnt = kmeans.shard_rows(SIZES, 2, r)[g]
                loc_off = np.concatenate([[0], np.cumsum([c for _, c in kmeans.shard_rows(SIZES, 2, r)])])
                parts.append(two_rank_run[r][f"labels_{mode}"][loc_off[g]:loc_off[g] + cnt])
        labels2 = np.concatenate(parts)
        assert len(labels2) == off[-1]
        assert np.array_equal(labels2, labels1), mode        # k-means labels: bit-exact across rank counts
        assert np.array_equal(two_rank_run[0][f"centers_{mode}"], two_rank_run[1][f"centers_{mode}"])   # ranks agree
        assert np.array_equal(two_rank_run[0][f"n_iter_{mode}"], np.array(single_run.n_iter))
    # the ordered reduction is invariant to the number of ranks: identical centre bits for world 1 and 2
    assert np.array_equal(two_rank_run[0]["centers_ordered"], single_ordered.centers.numpy())
    np.testing.assert_allclose(two_rank_run[0]["centers_allreduce"], single_run.centers.numpy(), rtol=1e-5, atol=1e-6)


def test_sharded_member_means(two_rank_run, single_run):
    single_run = single_run[0]
    segs = _data()
    lab = single_run.labels.numpy()
    off = np.concatenate([[0], np.cumsum(SIZES)])
    for g, n in enumerate(SIZES):
        for j in range(K):
            m = lab[off[g]:off[g + 1]] == j
            assert two_rank_run[0]["counts"][g, j] == m.sum()
            if m.any():
                np.testing.assert_allclose(two_rank_run[0]["means"][g, j], segs[g][m].mean(0), rtol=1e-5, atol=1e-6)
    assert np.array_equal(two_rank_run[0]["means"], two_rank_run[1]["means"])


def test_sharded_percentile_is_exact(two_rank_run):
    sc = _scores()
    for r in range(2):
        for g, v in enumerate(sc):
            if len(v) > 5:
                assert two_rank_run[r]["sel"][g] == np.percentile(v, 95.0, method="lower"), g
                assert two_rank_run[r]["mn"][g] == v.min() and two_rank_run[r]["mx"][g] == v.max()
            else:
                assert np.isnan(two_rank_run[r]["sel"][g])
        assert np.isnan(two_rank_run[r]["mn"][2])           # empty segment


def test_block_table_is_rank_count_invariant():
    for world in (1, 2, 4, 8):
        owned = [kmeans.shard_rows(SIZES, world, r) for r in range(world)]
        for g, n in enumerate(SIZES):
            pos = 0
            for r in kmeans.ranks_in_row_order(g, world):    # contiguous, ordered (rotating owner), complete cover of every segment
                a, cnt = owned[r][g]
                assert a == pos or cnt == 0
                pos += cnt
            assert pos == n
        tabs = [kmeans.build_blocks(SIZES, world, r, torch.device("cpu"))[0] for r in range(world)]
        assert sum(t.n_super_local for t in tabs) == tabs[0].n_super_global
        assert sum(t.n_blocks for t in tabs) == kmeans.build_blocks(SIZES, 1, 0, torch.device("cpu"))[0].n_blocks


def test_block_size_choice_and_big_block_tables():
    """Large all-reduce-mode fits take 1024-row CTA blocks, everything else (and always the rank-count-invariant "ordered"
    reduction) the fixed 512-row partition; a table of either block size covers the same rows, owns the same 4096-row
    super-blocks per rank, and the numpy stand-in fits to identical labels on both."""
    big = [200_000] * 20
    assert kmeans.auto_block_rows(big, 1) == kmeans.BIG_BLOCK_ROWS and kmeans.auto_block_rows(big, 4) == kmeans.BIG_BLOCK_ROWS
    assert kmeans.auto_block_rows(big, 8) == kmeans.BLOCK_ROWS                       # 3.3 waves of big blocks: not worth it
    assert kmeans.auto_block_rows(big, 1, "ordered") == kmeans.BLOCK_ROWS
    assert kmeans.auto_block_rows(SIZES, 1) == kmeans.BLOCK_ROWS
    dev = torch.device("cpu")
    sizes = [9000, 0, 4096, 4097, 1, 20000]
    for world in (1, 2, 3):
        for r in range(world):
            t_small, sh_small, off_small = kmeans.build_blocks(sizes, world, r, dev)
            t_big, sh_big, off_big = kmeans.build_blocks(sizes, world, r, dev, block_rows=1024)
            assert sh_small == sh_big and list(off_small) == list(off_big)           # ownership does not depend on the block size
            assert t_big.n_super_local == t_small.n_super_local and t_big.n_super_global == t_small.n_super_global
            for t, rows in ((t_small, 512), (t_big, 1024)):
                r0, r1, seg = t.row0.numpy(), t.row1.numpy(), t.seg.numpy()
                assert ((r1 - r0) > 0).all() and ((r1 - r0) <= rows).all()
                assert (r0[1:] == r1[:-1]).all() and (len(r0) == 0 or (r0[0] == 0 and r1[-1] == off_small[-1]))   # contiguous cover
                assert (np.diff(seg) >= 0).all()
                sf = t.super_first.numpy()
                assert sf[0] == 0 and sf[-1] == t.n_blocks and ((r1[sf[1:] - 1] - r0[sf[:-1]]) <= kmeans.SUPER_ROWS).all()
    with pytest.raises(AssertionError):
        kmeans.build_blocks(sizes, 1, 0, dev, block_rows=768)
    be = NumpyBackend()
    x = torch.from_numpy(np.concatenate(_data()))
    tab, shard, off = kmeans.build_blocks(SIZES, 1, 0, dev, block_rows=1024)
    a = kmeans.kmeans_fit(x, SIZES, K, tab, off, shard, backend=be)
    b = kmeans.kmeans_fit_predict_single(x, SIZES, K, backend=be)
    assert np.array_equal(a.labels.numpy(), b.labels.numpy())


def test_device_seeding_control_flow_matches_host_seeding():
    """seeding="device" (scan + search, gather, distance pass, pick as backend steps; uniforms drawn up front) picks
    the same seeds as the host loop written with sklearn's expressions, hence identical labels."""
    be = NumpyBackend()
    x = torch.from_numpy(np.concatenate(_data()))
    a = kmeans.kmeans_fit_predict_single(x, SIZES, K, backend=be, seeding="device")
    b = kmeans.kmeans_fit_predict_single(x, SIZES, K, backend=be, seeding="host")
    assert a.seconds["seeding"] == "device" and b.seconds["seeding"] == "host"
    assert np.array_equal(a.labels.numpy(), b.labels.numpy())
    assert np.allclose(a.centers.numpy(), b.centers.numpy(), rtol=0, atol=1e-6)


SMALL = [3000, 0, 2500]                    # one super-block per live segment, both owned by rank 1 (rotating ownership): rank 0 owns NO row


def _worker_empty_rank(rank: int, world: int, port: int, outdir: str):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        be = NumpyBackend()
        segs = [synth.blob_vectors(40 + i, n, DIM, K, 7.0)[0] if n else np.zeros((0, DIM), np.float32) for i, n in enumerate(SMALL)]
        shard = kmeans.shard_rows(SMALL, world, rank)
        local = [s[a:a + n] for s, (a, n) in zip(segs, shard)]
        x = torch.from_numpy(np.concatenate(local).reshape(-1, DIM))
        r = kmeans.kmeans_fit_sharded(x, [len(v) for v in local], SMALL, K, world, rank, group=dist.group.WORLD, backend=be)
        means, counts = kmeans.member_means(x, [len(v) for v in local], r.labels, K, group=dist.group.WORLD, backend=be)
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), labels=r.labels.numpy(), centers=r.centers.numpy(),
                 means=means.numpy(), counts=counts.numpy(), rows=np.array(x.shape[0]))
    finally:
        dist.destroy_process_group()


def test_a_rank_without_rows_takes_part_in_the_fit():
    """Fewer super-blocks per segment than ranks: rank 0 owns nothing, still joins every collective, and the result equals
    the single-process fit (the case the 8-GPU bench hit with 2 M vectors)."""
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker_empty_rank, args=(2, _free_port(), d), nprocs=2, join=True)
        out = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(2)]
    assert int(out[0]["rows"]) == 0 and int(out[1]["rows"]) == sum(SMALL)
    segs = [synth.blob_vectors(40 + i, n, DIM, K, 7.0)[0] if n else np.zeros((0, DIM), np.float32) for i, n in enumerate(SMALL)]
    single = kmeans.kmeans_fit_predict_single(torch.from_numpy(np.concatenate(segs)), SMALL, K, backend=NumpyBackend())
    assert len(out[0]["labels"]) == 0
    assert np.array_equal(out[1]["labels"], single.labels.numpy())
    assert np.array_equal(out[0]["centers"], out[1]["centers"]) and np.array_equal(out[0]["means"], out[1]["means"])
    np.testing.assert_allclose(out[1]["centers"], single.centers.numpy(), rtol=1e-5, atol=1e-6)
