"""GPU parity tests of the fit stage: K2 on vectors, K5 exact percentile, K4 k-means (labels bit-exact vs sklearn
on clustered data, and vs the golden labels produced by the reference's own call site)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from ood_in_object_detection_b200 import ops
    return ops.default_device()


def test_vec_score_matches_oracle(dev):
    from oracle import distance as D
    from ood_in_object_detection_b200 import ops
    rng = np.random.default_rng(0)
    dim, sizes, ks = 72, [50, 0, 300, 7], [3, 2, 10, 1]
    x = np.abs(rng.standard_normal((sum(sizes), dim))).astype(np.float32)
    cents = [(np.abs(rng.standard_normal((k, dim))) / np.sqrt(dim)).astype(np.float32) for k in ks]
    seg_off = np.concatenate([[0], np.cumsum(sizes)])
    cent = np.concatenate(cents)
    unit = np.concatenate([ops._unit_rows(c) for c in cents])
    crow = np.concatenate([[0], np.cumsum(ks)])[:-1]
    d, a = ops.vec_score(torch.from_numpy(x).to(dev), seg_off.tolist(), torch.from_numpy(cent).to(dev),
                         torch.from_numpy(unit).to(dev), crow.tolist(), ks, 0b111, True)
    d, a = d.cpu().numpy(), a.cpu().numpy()
    for m, metric in enumerate(("l1", "l2", "cosine")):
        for g in range(len(sizes)):
            if sizes[g] == 0:
                continue
            xs = D.normalize_rows(x[seg_off[g]:seg_off[g + 1]])
            pw = D.pairwise(cents[g], xs, metric)
            np.testing.assert_allclose(d[m, seg_off[g]:seg_off[g + 1]], pw.min(0), rtol=1e-5, atol=5e-7 if m == 2 else 0)
            assert (a[m, seg_off[g]:seg_off[g + 1]] != pw.argmin(0)).sum() <= 1


def test_exact_percentile_lower(dev, golden):
    from ood_in_object_detection_b200 import select
    g = golden("golden_thresholds.npz")
    rng = np.random.default_rng(1)
    segs = [g[f"v_{n}"] for n in (6, 11, 21, 101, 1001, 4097)]
    segs += [rng.standard_normal(5000).astype(np.float32) * 3, np.zeros(0, np.float32), np.full(17, -2.5, np.float32),
             rng.uniform(0, 1e-30, 300).astype(np.float32)]
    off = np.concatenate([[0], np.cumsum([len(s) for s in segs])])
    scores = torch.from_numpy(np.concatenate(segs)).to(dev)
    for q in (95.0, (1 - 0.95) * 100, 100 * 0.9, (1 - 0.9) * 100, 0.0, 100.0):
        ranks = [select.lower_index(len(s), q) if len(s) else None for s in segs]
        vals, mn, mx = select.segment_select(scores, off.tolist(), ranks)
        for i, s in enumerate(segs):
            if len(s) == 0:
                assert vals[i] is None and mn[i] is None
                continue
            assert vals[i] == float(np.percentile(s, q, method="lower")), (i, q)
            assert mn[i] == float(s.min()) and mx[i] == float(s.max())
    # golden: the reference's generate_thresholds on float32 scores
    for n in (6, 11, 21, 101, 1001, 4097):
        for tpr in (0.9, 0.95, 0.99, 0.8):
            v = g[f"v_{n}"]
            t = torch.from_numpy(v).to(dev)
            r, _, _ = select.segment_select(t, [0, n], [select.lower_index(n, 100 * tpr)])
            assert r[0] == g[f"dist_{n}_{tpr}"][0]
            r, _, _ = select.segment_select(t, [0, n], [select.lower_index(n, (1 - tpr) * 100)])
            assert r[0] == g[f"logit_{n}_{tpr}"][0]


def test_kmeans_labels_match_sklearn_and_reference(dev, golden):
    """Bit-exact labels on clustered data (strict convergence), several ragged segments in one launch."""
    from sklearn.cluster import KMeans
    from threadpoolctl import threadpool_limits
    from ood_in_object_detection_b200 import kmeans, synth
    g = golden("golden_kmeans.npz")
    # (1) the reference's own call site, KMeans_10 / KMeans_5 / n < k
    for tag in "abc":
        x, k = g[f"{tag}_x"], int(g[f"{tag}_k"])
        res = kmeans.kmeans_fit_predict_single(torch.from_numpy(x).to(dev), [len(x)], min(k, len(x)))
        assert np.array_equal(res.labels.cpu().numpy(), g[f"{tag}_labels"]), tag
    # (2) many segments at once, different sizes, vs sklearn per segment
    segs = [synth.blob_vectors(s, n, 48, kk, 7.0)[0] for s, n, kk in ((1, 3000, 8), (2, 700, 8), (3, 5, 3), (4, 20000, 8))]
    sizes = [len(s) for s in segs]
    res = kmeans.kmeans_fit_predict_single(torch.from_numpy(np.concatenate(segs)).to(dev), sizes, 8)
    lab = res.labels.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(sizes)])
    for i, s in enumerate(segs):
        with threadpool_limits(1):
            km = KMeans(n_clusters=min(8, len(s)), random_state=10).fit(s)
        assert np.array_equal(lab[off[i]:off[i + 1]], km.labels_), i
        assert res.n_iter[i] == km.n_iter_
        np.testing.assert_allclose(res.centers[i, :min(8, len(s))].cpu().numpy(), km.cluster_centers_, rtol=1e-4, atol=1e-5)


def test_kmeans_c3_shape_fast_path(dev):
    """D = 576, K = 16 (the TMA-fed register-tiled step kernel, 5 x 128 columns, ragged last chunk) vs sklearn."""
    from sklearn.cluster import KMeans
    from threadpoolctl import threadpool_limits
    from ood_in_object_detection_b200 import kmeans, synth
    segs = [synth.blob_vectors(40 + i, n, 576, 16, 8.0)[0] for i, n in enumerate((2500, 1037))]
    sizes = [len(s) for s in segs]
    x = torch.from_numpy(np.concatenate(segs)).to(dev)
    res = kmeans.kmeans_fit_predict_single(x, sizes, 16)
    lab = res.labels.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(sizes)])
    for i, s in enumerate(segs):
        with threadpool_limits(1):
            km = KMeans(n_clusters=16, random_state=10).fit(s)
        assert np.array_equal(lab[off[i]:off[i + 1]], km.labels_), i
        np.testing.assert_allclose(res.centers[i].cpu().numpy(), km.cluster_centers_, rtol=1e-4, atol=1e-5)
    means, counts = kmeans.member_means(x, sizes, res.labels, 16)         # update == 2 mode of the same kernel
    for i, s in enumerate(segs):
        for j in range(16):
            m = lab[off[i]:off[i + 1]] == j
            assert counts[i, j].item() == m.sum()
            np.testing.assert_allclose(means[i, j].cpu().numpy(), s[m].mean(0), rtol=1e-5, atol=1e-6)


def test_kmeans_realistic_overlap_agreement(dev):
    """Unstructured data: labels are compared through agreement / inertia (sklearn itself is not reproducible across
    thread counts there), not bit for bit."""
    from sklearn.cluster import KMeans
    from threadpoolctl import threadpool_limits
    from ood_in_object_detection_b200 import kmeans
    rng = np.random.default_rng(3)
    x = np.abs(rng.standard_normal((6000, 64))).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    res = kmeans.kmeans_fit_predict_single(torch.from_numpy(x).to(dev), [len(x)], 8)
    with threadpool_limits(1):
        km = KMeans(n_clusters=8, random_state=10).fit(x)
    lab = res.labels.cpu().numpy()
    agree = (lab == km.labels_).mean()
    c = res.centers[0].cpu().numpy()
    inertia = ((x - c[lab]) ** 2).sum()
    assert agree > 0.9 and inertia <= km.inertia_ * 1.01


def test_seed_scan_is_numpy_cumsum_searchsorted(dev):
    """The device scan + search == np.searchsorted(np.cumsum(float32), u * pot) bit for bit, segments split into
    several pieces (ranks), ragged sizes around the 128-value chunk and 4096-value batch boundaries."""
    from ood_in_object_detection_b200 import kmeans
    be = kmeans.CudaBackend(dev)
    rng = np.random.default_rng(11)
    sizes = [1, 127, 128, 129, 4095, 4096, 4097, 50001, 0, 200000]
    n_pieces, n_trials = 3, 4
    segs = [(rng.random(n) ** 4).astype(np.float32) for n in sizes]
    for s in segs[3:5]:
        s[::3] = 0.0                                                  # runs of equal cumulative sums
    # piece layout: piece r of every segment lives in row r of a padded [n_pieces, max_local] buffer
    cuts = [[(n * r) // n_pieces for r in range(n_pieces + 1)] for n in sizes]
    local = [sum(c[r + 1] - c[r] for c in cuts) for r in range(n_pieces)]
    max_local = max(local)
    buf = np.zeros((n_pieces, max_local), np.float32)
    piece_off = np.zeros((len(sizes), n_pieces), np.int64)
    piece_cnt = np.zeros((len(sizes), n_pieces), np.int64)
    fill = [0] * n_pieces
    for g, (s, c) in enumerate(zip(segs, cuts)):
        for r in range(n_pieces):
            cnt = c[r + 1] - c[r]
            buf[r, fill[r]:fill[r] + cnt] = s[c[r]:c[r + 1]]
            piece_off[g, r], piece_cnt[g, r] = r * max_local + fill[r], cnt
            fill[r] += cnt
    uni = rng.random((len(sizes), n_trials))
    uni[1, 0], uni[2, 1] = 0.0, 0.999999999                            # ends of the range
    trials = np.array([4, 4, 3, 4, 4, 2, 4, 4, 4, 4], np.int32)
    pot = np.array([np.float32(s.astype(np.float64).sum()) for s in segs], np.float32)
    pot[7] *= np.float32(1.01)                                         # targets beyond the total -> clipped to n - 1
    on = np.ones(len(sizes), np.int32)
    t = lambda a: torch.from_numpy(a).to(dev)
    cand = torch.full((len(sizes), n_trials), -7, dtype=torch.int64, device=dev)
    be.seed_scan(t(buf), t(piece_off), t(piece_cnt), t(uni), t(pot), t(trials), t(on), max(sizes), cand)
    got = cand.cpu().numpy()
    for g, s in enumerate(segs):
        if not len(s):
            continue
        ids = np.searchsorted(np.cumsum(s), uni[g, :trials[g]] * pot[g])
        np.clip(ids, None, len(s) - 1, out=ids)
        assert np.array_equal(got[g, :trials[g]], ids), (g, got[g], ids)
        assert (got[g, trials[g]:] == ids[0]).all()


def test_seed_scan_boundary_sums_are_the_sequential_float32_sums(dev):
    """The running sums the scan records at every 128-value boundary == np.cumsum(float32) there, bit for bit, on data built to
    break a parallel evaluation of a sequential float chain: exact rounding ties (values on a coarse binary grid), outliers that
    jump several binades at once, long runs of zeros, denormals, an all-zero segment, a constant segment, heavy tails."""
    from ood_in_object_detection_b200 import kmeans
    be = kmeans.CudaBackend(dev)
    rng = np.random.default_rng(23)
    n = 70001
    grid = rng.integers(0, 1024, n).astype(np.float32)                            # integers: ties once ulp(sum) = 2 (sum >= 2^24), 4, ...
    jumps = (rng.random(n) ** 6).astype(np.float32)
    jumps[[5, 1000, 1001, 30000, 65000]] = [3.0e4, 7.0e6, 1.0e-30, 4.0e9, 1.5e3]
    zeros = np.zeros(n, np.float32); zeros[40000:] = rng.random(n - 40000).astype(np.float32) * 1e-3
    den = (rng.random(n) * 1e-41).astype(np.float32); den[60000:] = 1e-20
    const = np.full(n, 0.1, np.float32)
    tails = np.exp(rng.standard_normal(n) * 4).astype(np.float32)
    halves = np.full(n, 513.0, np.float32); halves[::7] = 1025.0                  # odd: EVERY add is a tie once the sum passes 2^24
    big_ties = np.full(200000, 131.0, np.float32); big_ties[::2] = 129.0          # crosses 2^24 after ~129 k odd addends (ties above it)
    segs = [grid, jumps, zeros, den, const, tails, np.zeros(300, np.float32), halves, big_ties,
            (rng.random(200000) ** 4).astype(np.float32)]
    sizes = [len(v) for v in segs]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    # every segment in 3 pieces (ranks) stored out of order: [piece 2 | piece 0 | piece 1], cuts off the 128-value grid
    buf = np.zeros(int(off[-1]), np.float32)
    piece_off, piece_cnt = np.zeros((len(segs), 3), np.int64), np.zeros((len(segs), 3), np.int64)
    for g, v in enumerate(segs):
        c = [0, len(v) // 3 + 5, (2 * len(v)) // 3 + 77, len(v)]
        pos = int(off[g])
        for r in (2, 0, 1):
            cnt = c[r + 1] - c[r]
            buf[pos:pos + cnt] = v[c[r]:c[r + 1]]
            piece_off[g, r], piece_cnt[g, r] = pos, cnt
            pos += cnt
    n_trials = 8
    uni = rng.random((len(segs), n_trials))
    pot = np.array([np.cumsum(v)[-1] for v in segs], np.float32)
    t = lambda a: torch.from_numpy(a).to(dev)
    cand = torch.zeros((len(segs), n_trials), dtype=torch.int64, device=dev)
    be.seed_scan(t(buf), t(piece_off), t(piece_cnt), t(uni), t(pot), t(np.full(len(segs), n_trials, np.int32)),
                 t(np.ones(len(segs), np.int32)), max(sizes), cand)
    sums = be._sbuf.cpu().numpy()
    got = cand.cpu().numpy()
    for g, v in enumerate(segs):
        cum = np.cumsum(v)                                                        # numpy: sequential float32 adds
        ends = np.concatenate([cum[127::128], cum[-1:]]) if len(v) % 128 else cum[127::128]
        assert np.array_equal(sums[g, :len(ends)].view(np.uint32), ends.view(np.uint32)), \
            (g, int(np.flatnonzero(sums[g, :len(ends)] != ends)[0]))
        ids = np.clip(np.searchsorted(cum, uni[g] * pot[g]), None, len(v) - 1)
        assert np.array_equal(got[g], ids), (g, got[g], ids)


def test_seed_sqdist_matches_float64_expansion(dev):
    """Candidate distances: float32(float64 expansion) exactly like sklearn's `_euclidean_distances_upcast`, min
    against `closest`, bit-reproducible potentials; D = 576 (4 rows per warp) and an odd D (one row per warp)."""
    from ood_in_object_detection_b200 import kmeans
    be = kmeans.CudaBackend(dev)
    rng = np.random.default_rng(12)
    for dim, sizes in ((576, [3001, 0, 517]), (50, [999, 130])):
        n = sum(sizes)
        x = (rng.standard_normal((n, dim)) * 0.3).astype(np.float32)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        cand = (rng.standard_normal((len(sizes), 3, dim)) * 0.3).astype(np.float32)
        closest = (rng.random(n) * dim * 0.2).astype(np.float32)
        t = lambda a: torch.from_numpy(a).to(dev)
        out, pots = be.seed_sqdist(t(x), t(off), max(sizes), t(cand), t(closest))
        out2, pots2 = be.seed_sqdist(t(x), t(off), max(sizes), t(cand), t(closest))
        assert torch.equal(pots, pots2) and torch.equal(out, out2)
        out, pots = out.cpu().numpy(), pots.cpu().numpy()
        xs = x.astype(np.float64)
        for g in range(len(sizes)):
            a, b = off[g], off[g + 1]
            if a == b:
                continue
            y = cand[g].astype(np.float64)
            ref = (-2.0 * xs[a:b] @ y.T + (y * y).sum(1)[None] + (xs[a:b] ** 2).sum(1)[:, None])
            ref32 = np.minimum(np.maximum(ref.astype(np.float32), 0), closest[a:b, None])
            # the float64 sums differ from numpy's in the last bits only: after the cast to float32 at most 1 ulp
            np.testing.assert_allclose(out[:, a:b].T, ref32, rtol=2e-7, atol=1e-7)
            assert (out[:, a:b].T != ref32).mean() < 0.01
            np.testing.assert_allclose(pots[g], ref32.astype(np.float64).sum(0), rtol=1e-6)


def test_device_seeding_equals_host_seeding(dev, golden):
    """seeding="device" picks the seeds of the host loop (sklearn's expressions), hence the same labels: golden call
    sites of the reference, ragged multi-segment data, and the C3 shape."""
    from ood_in_object_detection_b200 import kmeans, synth
    g = golden("golden_kmeans.npz")
    for tag in "abc":
        x, k = g[f"{tag}_x"], int(g[f"{tag}_k"])
        res = kmeans.kmeans_fit_predict_single(torch.from_numpy(x).to(dev), [len(x)], min(k, len(x)), seeding="device")
        assert res.seconds["seeding"] == "device"
        assert np.array_equal(res.labels.cpu().numpy(), g[f"{tag}_labels"]), tag
    for dim, kk, spec in ((48, 8, ((1, 3000), (2, 700), (3, 5), (4, 20000), (5, 0))), (576, 16, ((40, 2500), (41, 1037)))):
        segs = [synth.blob_vectors(s, n, dim, kk, 7.5)[0] if n else np.zeros((0, dim), np.float32) for s, n in spec]
        sizes = [len(s) for s in segs]
        x = torch.from_numpy(np.concatenate(segs)).to(dev)
        a = kmeans.kmeans_fit_predict_single(x, sizes, kk, seeding="device")
        b = kmeans.kmeans_fit_predict_single(x, sizes, kk, seeding="host")
        assert torch.equal(a.labels, b.labels)
        assert a.n_iter == b.n_iter
        torch.testing.assert_close(a.centers, b.centers, rtol=0, atol=1e-6)


def test_tcgen05_step_equals_fp32_step(dev):
    """The tensor-core Lloyd step (csrc/kmeans_tc.cu: tcgen05 kind::tf32 with split-float operands, M-step from L2) vs
    the FP32 step kernel on the same centres: identical labels and changed counts, bit-identical block partial sums
    and counts (same row order), for ragged segments, an empty one, short last tiles, K < 16 and the D range."""
    from ood_in_object_detection_b200 import kmeans, synth
    tc, fp = kmeans.CudaBackend(dev), kmeans.CudaBackend(dev)
    tc.tensor_core, fp.tensor_core = True, False
    assert tc.lib.oodb200_kmeans_tc_workspace_bytes(3, 16, 576) > 0
    assert tc.lib.oodb200_kmeans_tc_workspace_bytes(3, 17, 576) == 0 and tc.lib.oodb200_kmeans_tc_workspace_bytes(3, 16, 100) == 0
    n_exact = 0
    for dim, k, spec in ((576, 16, ((1, 5000), (2, 777), (3, 33), (4, 0), (5, 1500))), (128, 5, ((6, 2000), (7, 1029))),
                         (640, 12, ((9, 1300), (10, 64)))):
        segs = [synth.blob_vectors(s, n, dim, k, 3.0)[0] if n else np.zeros((0, dim), np.float32) for s, n in spec]
        sizes = [len(s) for s in segs]
        x = torch.from_numpy(np.concatenate(segs)).to(dev)
        rng = np.random.default_rng(dim)
        cent = np.stack([s[rng.integers(0, len(s), k)] if len(s) else np.zeros((k, dim), np.float32) for s in segs])
        cent = torch.from_numpy(cent + 0.01 * rng.standard_normal(cent.shape).astype(np.float32)).to(dev)
        table, _, _ = kmeans.build_blocks(sizes, 1, 0, dev)
        seg_k = torch.tensor([min(k, n) for n in sizes], dtype=torch.int32, device=dev)
        out = []
        for be in (fp, tc):
            for update in (1, 0):
                labels = torch.full((x.shape[0],), -1, dtype=torch.int32, device=dev)
                chg = torch.zeros(len(sizes), dtype=torch.int32, device=dev)
                ps, pc = be.step(x, k, seg_k, cent, table, None, labels, chg, update)
                torch.cuda.synchronize()
                out.append((labels, chg, None if ps is None else ps.clone(), None if pc is None else pc.clone()))
        (l_fp, c_fp, ps_fp, pc_fp), (l_fp0, _, _, _), (l_tc, c_tc, ps_tc, pc_tc), (l_tc0, _, _, _) = out
        assert torch.equal(l_fp, l_fp0) and torch.equal(l_tc, l_tc0), dim
        differ = (l_fp != l_tc)
        # the two kernels round the cross-term differently (FP32 FMA chain vs split-float tensor-core products): a label
        # may differ only on a float32 tie (checked against float64 below); with identical labels everything is bit-equal
        assert differ.float().mean().item() <= 1e-3, dim
        if not differ.any():
            n_exact += 1
            assert torch.equal(c_fp, c_tc)
            assert torch.equal(pc_fp, pc_tc) and torch.equal(ps_fp, ps_tc), dim
        # labels are the nearest centre in float64 (up to float32 ties)
        off = np.concatenate([[0], np.cumsum(sizes)])
        for g, s in enumerate(segs):
            if not len(s):
                continue
            kg = min(k, len(s))
            d = ((s[:, None, :].astype(np.float64) - cent[g, :kg].cpu().numpy()[None].astype(np.float64)) ** 2).sum(-1)
            lab = l_tc[off[g]:off[g + 1]].cpu().numpy()
            best = d.min(1)
            assert np.all(d[np.arange(len(s)), lab] <= best * (1 + 1e-5) + 1e-7)
            lab_fp = l_fp[off[g]:off[g + 1]].cpu().numpy()
            assert np.all(d[np.arange(len(s)), lab_fp] <= best * (1 + 1e-5) + 1e-7)
    assert n_exact >= 2                                                  # the bit-equality claim was actually exercised


def test_segment_centring_kernels(dev):
    """seg_colsum / seg_center: float64 column sums, x - mean and the sum of squares per segment (KMeans.fit centring and
    `_tolerance`), ragged and empty segments, vector and scalar paths; bit-reproducible."""
    from ood_in_object_detection_b200 import kmeans
    be = kmeans.CudaBackend(dev)
    rng = np.random.default_rng(21)
    for dim, sizes in ((576, [3001, 0, 517, 64]), (50, [999, 1, 130])):
        x = (rng.standard_normal((sum(sizes), dim)) * 0.3 + 0.5).astype(np.float32)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        t = lambda a: torch.from_numpy(a).to(dev)
        sums = be.colsum(t(x), t(off), len(sizes))
        assert torch.equal(sums, be.colsum(t(x), t(off), len(sizes)))
        mean = np.zeros((len(sizes), dim), np.float32)
        for g, n in enumerate(sizes):
            ref = x[off[g]:off[g + 1]].sum(0, dtype=np.float64)
            np.testing.assert_allclose(sums[g].cpu().numpy(), ref, rtol=1e-12, atol=1e-9)
            mean[g] = (ref / max(n, 1)).astype(np.float32)
        out, sq = be.center(t(x), t(off), len(sizes), t(mean))
        for g in range(len(sizes)):
            ref = x[off[g]:off[g + 1]] - mean[g]
            assert np.array_equal(out[off[g]:off[g + 1]].cpu().numpy(), ref)
            np.testing.assert_allclose(sq[g].item(), (ref.astype(np.float64) ** 2).sum(), rtol=1e-12, atol=1e-12)


def test_c3_full_size_fit_properties(dev):
    """BASELINE config C3 at full size (4 M x 576 vectors, 20 classes x K = 16) through size-independent properties:
    every label is the nearest final centre (float64 check on a sample), the centres are the means of their members,
    the counts add up, the fit converged strictly (one more assignment changes nothing), and the tensor-core step and
    the FP32 step agree on the final assignment."""
    from ood_in_object_detection_b200 import kmeans
    if torch.cuda.get_device_properties(dev).total_memory < 60e9:
        pytest.skip("needs ~25 GB of device memory")
    n_seg, per, dim, k = 20, 200_000, 576, 16
    g = torch.Generator(device=dev).manual_seed(77)
    x = torch.empty((n_seg * per, dim), dtype=torch.float32, device=dev)
    for s in range(n_seg):
        centres = torch.randn((k, dim), generator=g, device=dev) * (8.0 / dim ** 0.5) + 1.0
        lab = torch.randint(0, k, (per,), generator=g, device=dev)
        blk = centres[lab] + torch.randn((per, dim), generator=g, device=dev) / dim ** 0.5
        x[s * per:(s + 1) * per] = blk / blk.norm(dim=1, keepdim=True)
        del blk
    sizes = [per] * n_seg
    res = kmeans.kmeans_fit_predict_single(x, sizes, k)
    assert res.seconds["seeding"] == "device"
    lab = res.labels
    assert int(lab.min()) >= 0 and int(lab.max()) < k
    assert all(res.strict), "well-separated mixtures converge strictly"
    cnt = torch.stack([torch.bincount(lab[s * per:(s + 1) * per].long(), minlength=k) for s in range(n_seg)])
    assert torch.equal(cnt.float(), res.counts) and int(cnt.sum()) == n_seg * per
    # nearest centre (float64) on a sample of rows of every segment
    idx = torch.randint(0, per, (2000,), generator=g, device=dev)
    for s in range(n_seg):
        rows = x[s * per + idx].double()
        d = ((rows[:, None, :] - res.centers[s].double()[None]) ** 2).sum(-1)
        mine = d.gather(1, lab[s * per + idx].long()[:, None])[:, 0]
        assert bool((mine <= d.min(1).values * (1 + 1e-6) + 1e-9).all()), s
    # centres are the member means (float64 reference) for two segments
    for s in (0, n_seg - 1):
        seg, sl = x[s * per:(s + 1) * per], lab[s * per:(s + 1) * per].long()
        ref = torch.zeros((k, dim), dtype=torch.float64, device=dev).index_add_(0, sl, seg.double()) / cnt[s][:, None].double()
        torch.testing.assert_close(res.centers[s].double(), ref, rtol=1e-5, atol=1e-6)
    # idempotence of the assignment with the final centres, on both step kernels
    table, _, _ = kmeans.build_blocks(sizes, 1, 0, dev)
    seg_k = torch.full((n_seg,), k, dtype=torch.int32, device=dev)
    xm = torch.stack([x[s * per:(s + 1) * per].mean(0, dtype=torch.float64) for s in range(n_seg)]).float()
    for tc in (True, False):
        be = kmeans.CudaBackend(dev)
        be.tensor_core = tc
        xc, _ = be.center(x, torch.arange(n_seg + 1, device=dev, dtype=torch.int64) * per, n_seg, xm)
        l2 = lab.clone()
        chg = torch.zeros(n_seg, dtype=torch.int32, device=dev)
        be.step(xc, k, seg_k, (res.centers - xm[:, None, :]).contiguous(), table, None, l2, chg, 0)
        assert int(chg.sum()) == 0 and torch.equal(l2, lab), tc
        del xc


def test_tcgen05_vector_scoring_c5_shapes(dev):
    """K2b (tcgen05 cross-term, K = 64 clusters per class: BASELINE config C5) vs the float64 statement of sklearn's
    `euclidean_distances` / `cosine_distances` and vs the FP32 vec_score kernel: distances within the 1e-3 tier the
    north star grants the tensor-core path (measured ~1e-6), arg-min equal except float32 ties, decisions equal away
    from the threshold.  YOLOv8x channel counts (320 / 640), ragged segments, K < 64, a segment without centroids."""
    from ood_in_object_detection_b200 import ops
    rng = np.random.default_rng(64)
    for dim, spec in ((640, ((300, 64), (80, 64), (0, 64), (1500, 37), (513, 0), (129, 1))), (320, ((700, 64), (33, 5)))):
        sizes, ks = [n for n, _ in spec], [k for _, k in spec]
        n = sum(sizes)
        x = np.abs(rng.standard_normal((n, dim))).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        cent = np.abs(rng.standard_normal((sum(ks), dim))).astype(np.float32)
        cent = (0.7 * cent / np.linalg.norm(cent, axis=1, keepdims=True)).astype(np.float32)
        cunit = (cent / np.maximum(np.linalg.norm(cent, axis=1, keepdims=True), 1e-12)).astype(np.float32)
        off = np.concatenate([[0], np.cumsum(sizes)]).tolist()
        crow = np.concatenate([[0], np.cumsum(ks)])[:-1].tolist()
        xd, cd, cud = torch.from_numpy(x).to(dev), torch.from_numpy(cent).to(dev), torch.from_numpy(cunit).to(dev)
        d_fp, a_fp = ops.vec_score(xd, off, cd, cud, crow, ks, 0b110, normalize=False)
        for metric, cmat in (("l2", cd), ("cosine", cud)):
            slot = ops.METRIC_SLOT[metric]
            # thresholds: the median distance of every segment
            ref_d, ref_a = np.full(n, 1000.0), np.full(n, -1)
            for g, (a, b) in enumerate(zip(off[:-1], off[1:])):
                if ks[g] == 0 or a == b:
                    continue
                xs, cs = x[a:b].astype(np.float64), cent[crow[g]:crow[g] + ks[g]].astype(np.float64)
                if metric == "l2":
                    dm = np.sqrt(np.maximum(((xs ** 2).sum(1)[:, None] - 2 * xs @ cs.T + (cs ** 2).sum(1)[None]).astype(np.float32), 0))
                else:
                    cu = cunit[crow[g]:crow[g] + ks[g]].astype(np.float64)
                    dm = np.clip(1.0 - (xs / np.linalg.norm(xs, axis=1, keepdims=True)) @ cu.T, 0, 2)
                ref_d[a:b], ref_a[a:b] = dm.min(1), dm.argmin(1)
            thr = np.full((3, len(sizes)), np.nan)
            for g, (a, b) in enumerate(zip(off[:-1], off[1:])):
                if b > a and ks[g]:
                    thr[slot, g] = np.median(ref_d[a:b])
            dist, arg, dec = ops.vec_score_tc(xd, off, cmat, crow, ks, metric, thr=torch.from_numpy(thr).to(dev))
            torch.cuda.synchronize()
            got, ga = dist[slot].cpu().numpy(), arg[slot].cpu().numpy()
            np.testing.assert_allclose(got, ref_d, rtol=1e-3, atol=1e-6)          # the tier of the tensor-core path
            np.testing.assert_allclose(got, ref_d, rtol=2e-5, atol=2e-6)          # what split-float operands actually give
            np.testing.assert_allclose(got, d_fp[slot].cpu().numpy(), rtol=2e-5, atol=2e-6)
            diff = ga != ref_a
            assert diff.mean() <= 2e-3
            for g, (a, b) in enumerate(zip(off[:-1], off[1:])):
                if ks[g] == 0:
                    assert (got[a:b] == 1000).all() and (ga[a:b] == -1).all() and (dec[slot, a:b] == 0).all()
                elif b > a:
                    near = np.abs(ref_d[a:b] - thr[slot, g]) <= 1e-4 * abs(thr[slot, g])
                    assert np.array_equal(dec[slot, a:b].cpu().numpy()[~near], (ref_d[a:b] < thr[slot, g]).astype(np.uint8)[~near])


def test_pair_cluster_sums_match_the_oracle(dev):
    """K7 against the dense pairwise matrix of the oracle: every (row, cluster) sum, all three metrics, ragged sizes
    (n not a multiple of the 64-row tile, D not a multiple of 4 or 16), one split over column tiles."""
    from oracle import distance as D, silhouette as S
    from ood_in_object_detection_b200 import ops
    rng = np.random.default_rng(5)
    for n, dim, kc in ((1, 8, 1), (70, 24, 3), (333, 50, 7), (900, 37, 14)):
        x = np.abs(rng.standard_normal((n, dim))).astype(np.float32) + 0.05
        lab = rng.integers(0, kc, size=n).astype(np.int32)
        for metric in ("l1", "l2", "cosine"):
            d = S.pairwise_full(x, metric).astype(np.float64)
            want = np.stack([d[:, lab == c].sum(1) for c in range(kc)], 1)
            xs = torch.from_numpy(D.normalize_rows(x) if metric == "cosine" else x).to(dev)
            got = ops.pair_cluster_sums(xs, torch.from_numpy(lab).to(dev), kc, metric).cpu().numpy()
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6 * n if metric == "cosine" else 1e-9)
            # the stored-matrix route of the k-search: the same float32 distances (symmetric, zero diagonal), folded per labeling
            pairs = ops.PairDistances(torch.from_numpy(x).to(dev), metric)
            m = pairs.matrix.cpu().numpy()
            assert m.shape == (n, n) and np.array_equal(m, m.T) and not m.diagonal().any()
            np.testing.assert_allclose(m, d, rtol=2e-6, atol=3e-7 if metric == "cosine" else 1e-9)
            got_m = pairs.cluster_sums(torch.from_numpy(lab).to(dev), kc).cpu().numpy()
            fold = ops.pair_cluster_sums(pairs.xs, torch.from_numpy(lab).to(dev), kc, metric).cpu().numpy()     # same rows, fold route
            np.testing.assert_allclose(got_m, fold, rtol=1e-12, atol=1e-12 * n)
            np.testing.assert_allclose(got_m, want, rtol=1e-5, atol=2e-6 * n if metric == "cosine" else 1e-9)
            lab2 = rng.integers(0, kc, size=n).astype(np.int32)             # a second labeling of the same rows
            want2 = np.stack([d[:, lab2 == c].sum(1) for c in range(kc)], 1)
            np.testing.assert_allclose(pairs.cluster_sums(torch.from_numpy(lab2).to(dev), kc).cpu().numpy(), want2,
                                       rtol=1e-5, atol=2e-6 * n if metric == "cosine" else 1e-9)
            assert ops.PairDistances(torch.from_numpy(x).to(dev), metric, max_bytes=0).matrix is None


def test_silhouette_score_matches_sklearn(dev, golden):
    """ops.silhouette_score / calinski_harabasz_score == sklearn on the reference's final labels (1e-5, stated), also with a
    one-sample cluster (scores 0 for its member) and labels that are not 0..k-1."""
    from sklearn.metrics import calinski_harabasz_score, silhouette_score
    from ood_in_object_detection_b200 import ops
    g = golden("golden_ksearch.npz")
    for tag in "abcd":
        x, lab, metric = g[f"{tag}_x"], g[f"{tag}_labels"].copy(), str(g[f"{tag}_metric"])
        for variant in range(3):
            if variant == 1:
                lab = lab * 3 + 2                                   # sparse label values
            if variant == 2:
                lab[0] = lab.max() + 5                              # a cluster of one sample
            xd, ld = torch.from_numpy(x).to(dev), torch.from_numpy(lab.astype(np.int32)).to(dev)
            for m in ("l1", "l2", "cosine") if variant == 0 else (metric,):
                assert abs(ops.silhouette_score(xd, ld, m) - silhouette_score(x, lab, metric=m)) < 1e-5, (tag, variant, m)
            assert np.isclose(ops.calinski_harabasz_score(xd, ld), calinski_harabasz_score(x, lab), rtol=1e-5)


def test_k_search_matches_the_reference(dev, golden):
    """`cluster_method='KMeans'` (cluster_utils.py:75-80, :160-186, :203-356): per-k scores within 1e-4 of the reference's
    (k-means labels are bit-exact on these separated sets, so only the arithmetic of the score differs), same best k,
    identical final labels; through the mirror of the reference's function and through DistanceMethod.generate_clusters."""
    import logging
    from ood_in_object_detection_b200 import cluster_utils, ood_utils
    log = logging.getLogger("t")
    log.setLevel(logging.CRITICAL)
    g = golden("golden_ksearch.npz")
    for tag in "abcde":
        x, metric, perf = g[f"{tag}_x"], str(g[f"{tag}_metric"]), str(g[f"{tag}_perf"])
        ref_scores, ref_labels = g[f"{tag}_scores"], g[f"{tag}_labels"]
        labels, scores, ks = cluster_utils.search_number_of_clusters(torch.from_numpy(x).to(dev), metric, perf, log)
        assert ks == g[f"{tag}_ks"].tolist()
        assert int(np.argmax(scores)) == int(np.argmax(ref_scores)), (tag, scores, ref_scores.tolist())
        np.testing.assert_allclose(scores, ref_scores, rtol=1e-4, atol=1e-4)
        assert np.array_equal(labels.cpu().numpy(), ref_labels), tag
        lab2 = cluster_utils.find_optimal_number_of_clusters_one_class_one_stride_and_return_labels(x, "KMeans", metric, perf, "", log)
        assert np.array_equal(lab2, ref_labels)
    # the class surface: centroids = member means of the searched labels (ood_utils.py:2359-2366)
    KW = dict(agg_method="mean", cluster_method="KMeans", cluster_optimization_metric="silhouette",
              ind_info_creation_option="valid_preds_one_stride", which_internal_activations="ftmaps_and_strides",
              iou_threshold_for_matching=0.5, min_conf_threshold_train=0.15, min_conf_threshold_test=0.15)
    m = ood_utils.L2DistanceOneClusterPerStride(**KW)
    acts = [[np.empty(0) for _ in range(3)] for _ in range(2)]
    place = {(0, 0): "a", (0, 1): "b", (1, 1): "d", (1, 2): "c"}         # segments of one stride share their dimension
    for (c, s), tag in place.items():
        acts[c][s] = g[f"{tag}_x"][:, :, None, None]
    clusters = m.generate_clusters(acts, log)
    from oracle import distance as D
    xn, lab = D.normalize_rows(g["a_x"]), g["a_labels"]                  # 'a' was searched with l2 / silhouette: same labels
    want = np.stack([xn[lab == j].mean(0) for j in sorted(set(lab.tolist()))])
    np.testing.assert_allclose(clusters[0][0], want, rtol=1e-5, atol=1e-6)
    for (c, s), tag in place.items():
        k_true = {"a": 5, "b": 3, "c": 4, "d": 6}[tag]
        assert clusters[c][s].shape == (k_true, g[f"{tag}_x"].shape[1]), (tag, clusters[c][s].shape)
    assert len(clusters[0][2]) == 0 and len(clusters[1][0]) == 0


def test_k_search_batched_fits_equal_single_fits(dev):
    """The k-search fits every candidate k as one segment of ONE segmented fit over copies of the rows: labels and centres
    equal those of a fit of the rows alone with that k (FP32 kernels at D = 64, the tcgen05 step at D = 128)."""
    from ood_in_object_detection_b200 import kmeans
    rng = np.random.default_rng(21)
    for n, dim in ((2500, 64), (4500, 128)):
        cen = rng.standard_normal((5, dim)).astype(np.float32) * 1.5
        x = torch.from_numpy(np.abs(cen[rng.integers(0, 5, n)] + rng.standard_normal((n, dim)).astype(np.float32))).to(dev)
        ks = list(range(2, 15))
        res = kmeans.kmeans_fit_predict_single(x.repeat(len(ks), 1), [n] * len(ks), max(ks), seg_k=ks, seeding="host")
        for i, k in enumerate(ks):
            one = kmeans.kmeans_fit_predict_single(x, [n], k, seeding="host")
            assert torch.equal(res.labels[i * n:(i + 1) * n], one.labels), (n, dim, k)
            assert torch.equal(res.centers[i, :k], one.centers[0, :k]), (n, dim, k)
    with pytest.raises(ValueError):
        kmeans.kmeans_fit_predict_single(x, [n], 4, seg_k=[5])
