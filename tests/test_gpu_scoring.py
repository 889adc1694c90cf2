"""GPU parity tests: the CUDA kernels (through the C ABI) against the golden vectors frozen from the
reference and against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): decisions / assignments bit-exact; distances and scores within
1e-5 relative in float32.  A decision is exempt only when the reference distance sits within 1e-5
relative of its threshold (SURVEY.md §7 "threshold ties"); such boxes are counted and must be rare.
"""
import numpy as np
import pytest
import torch

from tests.helpers import scoring_case, split, unpack_nested

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    from ood_in_object_detection_b200 import ops as _ops
    _ops.default_device()
    return _ops


def _batch_from_images(ops, images, img_w):
    return ops.make_batch([[torch.from_numpy(m) for m in im["maps"]] for im in images],
                          [torch.from_numpy(im["boxes"]) for im in images],
                          [torch.from_numpy(im["strides"]) for im in images],
                          [torch.from_numpy(im["cls"]) for im in images], img_w)


def test_roi_pool_edges(ops, golden):
    """K1 against the reference extractor on border / degenerate / full-image boxes, both stride modes."""
    g = golden("golden_roi_edges.npz")
    n = g["n_boxes"]
    img = int(g["img"])
    maps = [torch.from_numpy(g[f"map{s}"]) for s in range(3)]
    boxes, strides = split(g["boxes"], n), split(g["strides"], n)
    cls = [np.zeros(len(b), np.float32) for b in boxes]
    batch = ops.make_batch(maps, boxes, strides, cls, img)
    pooled = ops.roi_pool(batch).cpu().numpy()
    start = np.concatenate([[0], np.cumsum(n)])
    for i in range(3):
        for s in range(3):
            idx = g[f"all0_idx_{i}_{s}"].astype(int)
            exp = g[f"all0_feat_{i}_{s}"]
            got = pooled[start[i] + idx][:, :exp.shape[1]]
            scale = np.abs(exp).max(axis=1, keepdims=True) + 1e-30
            assert np.all(np.abs(got - exp) <= RTOL * scale), (i, s)
    # extract_all_strides=True == every box pooled on each of the three maps
    for s in range(3):
        b2 = ops.make_batch(maps, boxes, [np.full(len(b), s, np.float32) for b in boxes], cls, img)
        p2 = ops.roi_pool(b2).cpu().numpy()
        for i in range(3):
            exp = g[f"all1_feat_{i}_{s}"]
            got = p2[start[i]:start[i + 1]][:, :exp.shape[1]]
            scale = np.abs(exp).max(axis=1, keepdims=True) + 1e-30
            assert np.all(np.abs(got - exp) <= RTOL * scale), (i, s)


@pytest.mark.parametrize("name", ["golden_c1_one.npz", "golden_small_kmeans5.npz"])
def test_fused_scoring_matches_reference(ops, golden, name):
    """K1+K2 (+Q1 plan): distances 1e-5, arg-min / class-used / order / decisions exact, all three metrics in ONE pass."""
    from oracle import decide
    g = golden(name)
    images, _ = scoring_case(g)
    nc, img = int(g["nc"]), int(g["img"])
    batch = _batch_from_images(ops, images, img)
    dims = [int(c) for c in g["channels"]]
    clusters = unpack_nested(g, "l2_clusters", nc)     # centroids do not depend on the metric
    for tag in ("l1", "cos"):
        for c in range(nc):
            for s in range(3):
                assert np.array_equal(g[f"{tag}_clusters_{c}_{s}"], clusters[c][s])
    thr = {ops.METRIC_SLOT[m]: unpack_nested(g, f"{t}_thr", nc, as_threshold=True)
           for t, m in (("l1", "l1"), ("l2", "l2"), ("cos", "cosine"))}
    table = ops.pack_centroids(clusters, thr, dims)
    res = ops.fmap_score(batch, table, 0b111, True, compat_q1=True, want_plan=True)
    torch.cuda.synchronize()
    dist, dec, arg = res.dist.cpu().numpy(), res.decision.cpu().numpy(), res.argmin.cpu().numpy()
    # Q1 bookkeeping: class looked up with the in-stride index, stride-major output order
    pos = res.out_index.cpu().numpy()
    assert np.array_equal(np.sort(pos), np.arange(batch.n))
    cls_out = np.empty(batch.n, np.int32)
    cls_out[pos] = res.cls_used.cpu().numpy()
    cu2, oi2 = ops.q1_plan(batch)                      # the standalone plan entry agrees
    assert np.array_equal(oi2.cpu().numpy(), pos) and np.array_equal(cu2.cpu().numpy(), res.cls_used.cpu().numpy())
    n_exempt = 0
    for tag, metric in (("l1", "l1"), ("l2", "l2"), ("cos", "cosine")):
        m = ops.METRIC_SLOT[metric]
        assert np.array_equal(cls_out, g[f"{tag}_cls_used"])
        ref_d = g[f"{tag}_dist"]
        np.testing.assert_allclose(dist[m], ref_d, rtol=RTOL, atol=(5e-7 if metric == "cosine" else 0))
        thr_flat = np.array([thr[m][c][s] if thr[m][c][s] != [] else np.nan
                             for c, s in zip(g[f"{tag}_cls_used"], g[f"{tag}_stride_of"])], np.float64)
        near = np.abs(ref_d - thr_flat) <= RTOL * np.abs(thr_flat)
        n_exempt += int(near.sum())
        assert np.array_equal(dec[m][~near], g[f"{tag}_decisions"][~near]), tag
        # arg-min against the oracle (the reference only returns the min)
        _, det = decide.distance_decisions(images, clusters, thr[m], metric, return_details=True)
        assert np.array_equal(arg[m], np.array([t[1] for im in det for t in im])), tag
    assert n_exempt <= 2


def test_quirks(ops, golden):
    g = golden("golden_quirks.npz")
    img = int(g["img"])
    im = dict(maps=[g[f"map{s}"][0] for s in range(3)], boxes=g["boxes"], cls=g["cls"], strides=g["strides"])
    batch = _batch_from_images(ops, [im], img)
    clusters = unpack_nested(g, "clusters", 3)
    dims = [c.shape[1] for c in clusters[0]]

    def run(cl, thr, compat=True):
        table = ops.pack_centroids(cl, {1: thr}, dims)
        r = ops.fmap_score(batch, table, 0b010, True, compat_q1=compat)
        return r.decision[1].cpu().numpy().tolist(), r.dist[1].cpu().numpy()

    cases = {"q1": [[1e9] * 3, [1e-9] * 3, [1e-9] * 3], "q4_zero": [[0.0] * 3, [1e9] * 3, [1e9] * 3],
             "q4_empty": [[[]] * 3, [1e9] * 3, [1e9] * 3]}
    for k, thr in cases.items():
        assert run(clusters, thr)[0] == g[f"{k}_decisions"].tolist(), k
    assert run(clusters, cases["q1"], compat=False)[0] == [0, 1, 0]          # intended (box-order) semantics
    miss = [clusters[0], [np.empty(0), clusters[1][1], clusters[1][2]], clusters[2]]
    d, dist = run(miss, [[1e9] * 3] * 3)
    assert d == g["missing_cluster_decisions"].tolist() and dist[0] == 1000.0
    assert run(miss, [[999.0] * 3] * 3)[0] == g["missing_cluster_thr999_decisions"].tolist()


def test_logits_and_fusion(ops, golden):
    g = golden("golden_logits.npz")
    dev = ops.default_device()
    nc = int(g["nc"])
    logits = torch.from_numpy(g["logits"]).to(dev)
    cls = torch.from_numpy(g["cls"]).to(dev).to(torch.int32)
    names = ("MSP", "Energy", "ODIN", "Sigmoid")
    thr = np.zeros((5, nc))
    mn = np.zeros((5, nc))
    mx = np.zeros((5, nc))
    for nme in names:
        k = ops.LOGIT_SLOT[nme]
        thr[k], mn[k], mx[k] = g[f"{nme}_thr"], g[f"{nme}_min"], g[f"{nme}_max"]
    t = lambda a: torch.from_numpy(a).to(dev)
    out = ops.logit_score(logits, cls, 0b11111, 1.0, 1000.0, t(thr), t(mn), t(mx))
    torch.cuda.synchronize()
    assert int(out.sigmoid_mismatch.item()) == 0
    sc, dec, ind = out.scores.cpu().numpy(), out.decision.cpu().numpy(), out.indness.cpu().numpy()
    fitted = g["fitted_mask"]
    for nme in names:
        k = ops.LOGIT_SLOT[nme]
        np.testing.assert_allclose(sc[k], g[f"{nme}_scores"], rtol=RTOL)
        near = np.abs(g[f"{nme}_scores"] - thr[k][g["cls"].astype(int)]) <= RTOL * np.abs(thr[k][g["cls"].astype(int)])
        assert near.sum() <= 2
        assert np.array_equal(dec[k][~near], g[f"{nme}_decisions"][~near]), nme
        tol = 2e-2 if nme == "ODIN" else 1e-4      # ODIN's InD score range is ~1e-4 wide: INDness amplifies f32 rounding
        np.testing.assert_allclose(ind[k][fitted], g[f"{nme}_indness"], atol=tol)
    np.testing.assert_array_equal(sc[4], g["logits"].max(axis=1))             # MaxLogit: defined here, no reference
    # fusion rules
    f = golden("golden_fusion.npz")
    u8 = lambda a: torch.from_numpy(a.astype(np.uint8)).to(dev)
    assert np.array_equal(ops.fuse_decisions(u8(f["d1"]), u8(f["d2"]), "and").cpu().numpy(), f["fuse_and"])
    assert np.array_equal(ops.fuse_decisions(u8(f["d1"]), u8(f["d2"]), "or").cpu().numpy(), f["fuse_or"])
    assert np.array_equal(ops.fuse_decisions(u8(f["d1"]), u8(f["d2"]), "majority_voting", u8(f["d3"])).cpu().numpy(),
                          f["fuse_majority"])
    f32 = lambda a: torch.from_numpy(a.astype(np.float32)).to(dev)
    assert np.array_equal(ops.fuse_scores(f32(f["s1"]), f32(f["s2"])).cpu().numpy(), f["fuse_score"])


@pytest.mark.parametrize("cfg,batch,lam,k", [("C2", 4, 50, 10), ("C5", 1, 120, 16), ("C5", 1, 60, 64)])
def test_fused_scoring_vs_oracle_at_config_shapes(ops, cfg, batch, lam, k):
    """Real YOLOv8s / YOLOv8x map shapes (incl. 160x160 maps and windows > 256 cells), random centroids."""
    from oracle import decide
    from ood_in_object_detection_b200 import synth
    wl = synth.CONFIGS[cfg]
    maps = synth.feature_maps(99, batch, wl.channels, wl.map_hw)
    det = synth.detections(100, batch, wl.img, wl.nc, lam)
    det["boxes"][0][0] = [0, 0, wl.img, wl.img]                               # full-image box on its stride
    for j in range(3):                                                        # whole-image / wide boxes forced onto EVERY stride:
        det["boxes"][0][1 + j] = [3, 5, wl.img - 2, wl.img - 7]               # windows of > 256 16-byte chunks (tiled path)
        det["strides"][0][1 + j] = j
    det["boxes"][0][4] = [0, 10, wl.img, 40]                                  # one very wide, flat window on the finest map
    det["strides"][0][4] = 0
    rng = np.random.default_rng(5)
    clusters = [[np.abs(rng.standard_normal((k, c))).astype(np.float32) / np.sqrt(c) * 1.3 for c in wl.channels]
                for _ in range(wl.nc)]
    images = [dict(maps=[m[i] for m in maps], boxes=det["boxes"][i], cls=det["cls"][i], strides=det["strides"][i],
                   img_hw=(wl.img, wl.img)) for i in range(batch)]
    batch_d = _batch_from_images(ops, images, wl.img)
    thr_by = {}
    oracle_out = {}
    for metric in ("l1", "l2", "cosine"):
        big = [[1e9] * 3 for _ in range(wl.nc)]
        _, det_o = decide.distance_decisions(images, clusters, big, metric, return_details=True)
        d = np.array([t[0] for im in det_o for t in im])
        med = float(np.median(d))
        thr = [[med] * 3 for _ in range(wl.nc)]                               # ~half InD, half OoD
        thr_by[ops.METRIC_SLOT[metric]] = thr
        oracle_out[metric] = (d, np.array([t[1] for im in det_o for t in im]), med)
    table = ops.pack_centroids(clusters, thr_by, list(wl.channels))
    res = ops.fmap_score(batch_d, table, 0b111, True, compat_q1=True)
    torch.cuda.synchronize()
    for metric, (d, a, med) in oracle_out.items():
        m = ops.METRIC_SLOT[metric]
        got = res.dist[m].cpu().numpy()
        np.testing.assert_allclose(got, d, rtol=RTOL, atol=(5e-7 if metric == "cosine" else 0))
        near = np.abs(d - med) <= RTOL * med
        assert np.array_equal(res.decision[m].cpu().numpy()[~near], (d < med).astype(np.uint8)[~near])
        # arg-min may legitimately differ only when two centroids are within rounding of each other
        diff = res.argmin[m].cpu().numpy() != a
        assert diff.sum() <= 1


def test_empty_batch_and_invalid_args(ops):
    from ood_in_object_detection_b200 import synth
    wl = synth.CONFIGS["C1"]
    maps = [torch.from_numpy(m) for m in synth.feature_maps(1, 2, wl.channels, wl.map_hw)]
    empty = [np.zeros((0, 4), np.float32)] * 2
    batch = ops.make_batch(maps, empty, [np.zeros(0, np.float32)] * 2, [np.zeros(0, np.float32)] * 2, wl.img)
    assert ops.roi_pool(batch).shape[0] == 0
    table = ops.pack_centroids([[np.ones((1, c), np.float32) for c in wl.channels]], {1: [[1.0] * 3]}, list(wl.channels))
    assert ops.fmap_score(batch, table, 0b010).dist.shape == (3, 0)
    with pytest.raises(RuntimeError, match="metric_mask"):
        ops.fmap_score(batch, table, 0)


def _channels_last(m: np.ndarray) -> torch.Tensor:
    """[B, C, H, W] array -> tensor with the same values and shape whose memory is [B, H, W, C]."""
    return torch.from_numpy(np.ascontiguousarray(m)).contiguous(memory_format=torch.channels_last)


def test_channels_last_maps_give_the_reference_results(ops, golden):
    """The *_nhwc entry points (maps as torch.channels_last hands them out): pooled vectors within 1e-5 of the reference
    extractor on the edge boxes; fused decisions / arg-min / order identical to the default-layout pass on the golden
    batches, distances within 1e-5; batched tensors, per-image views and host tensors all keep their layout."""
    g = golden("golden_roi_edges.npz")
    n, img = g["n_boxes"], int(g["img"])
    boxes, strides = split(g["boxes"], n), split(g["strides"], n)
    cls = [np.zeros(len(b), np.float32) for b in boxes]
    maps_cl = [_channels_last(g[f"map{s}"]) for s in range(3)]
    batch = ops.make_batch(maps_cl, boxes, strides, cls, img)
    assert batch.nhwc
    pooled = ops.roi_pool(batch).cpu().numpy()
    start = np.concatenate([[0], np.cumsum(n)])
    for i in range(3):
        for s in range(3):
            idx, exp = g[f"all0_idx_{i}_{s}"].astype(int), g[f"all0_feat_{i}_{s}"]
            got = pooled[start[i] + idx][:, :exp.shape[1]]
            scale = np.abs(exp).max(axis=1, keepdims=True) + 1e-30
            assert np.all(np.abs(got - exp) <= RTOL * scale), (i, s)
    for name in ("golden_c1_one.npz", "golden_small_kmeans5.npz"):
        g = golden(name)
        images, maps = scoring_case(g)
        nc, img = int(g["nc"]), int(g["img"])
        dims = [int(c) for c in g["channels"]]
        clusters = unpack_nested(g, "l2_clusters", nc)
        thr = {ops.METRIC_SLOT[m]: unpack_nested(g, f"{t}_thr", nc, as_threshold=True)
               for t, m in (("l1", "l1"), ("l2", "l2"), ("cos", "cosine"))}
        table = ops.pack_centroids(clusters, thr, dims)
        ref = ops.fmap_score(_batch_from_images(ops, images, img), table, 0b111, True, compat_q1=True, want_plan=True)
        args = ([torch.from_numpy(im["boxes"]) for im in images], [torch.from_numpy(im["strides"]) for im in images],
                [torch.from_numpy(im["cls"]) for im in images], img)
        batched = [_channels_last(m) for m in maps]                               # host, batched
        per_image = [[m[i] for m in batched] for i in range(len(images))]         # host, per-image views
        on_device = [m.cuda() for m in batched]                                   # device, batched
        for maps_in in (batched, per_image, on_device):
            b = ops.make_batch(maps_in, *args)
            assert b.nhwc
            res = ops.fmap_score(b, table, 0b111, True, compat_q1=True, want_plan=True)
            assert torch.equal(res.out_index, ref.out_index) and torch.equal(res.cls_used, ref.cls_used)
            d, dr = res.dist.cpu().numpy(), ref.dist.cpu().numpy()
            np.testing.assert_allclose(d, dr, rtol=RTOL, atol=5e-7)
            assert (res.argmin != ref.argmin).sum().item() == 0
            for slot, tag in ((0, "l1"), (1, "l2"), (2, "cos")):
                thr_flat = np.array([thr[slot][c][s] if thr[slot][c][s] != [] else np.nan
                                     for c, s in zip(g[f"{tag}_cls_used"], g[f"{tag}_stride_of"])], np.float64)
                near = np.abs(g[f"{tag}_dist"] - thr_flat) <= RTOL * np.abs(thr_flat)
                assert np.array_equal(res.decision[slot].cpu().numpy()[~near], g[f"{tag}_decisions"][~near]), (name, tag)
