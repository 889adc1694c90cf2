"""numpy stand-in for the device steps of the fit (kmeans.CudaBackend / select's histogram kernel).

TEST INFRASTRUCTURE ONLY: it lets the multi-rank control flow of kmeans.kmeans_fit and select.segment_select
(row sharding, the per-iteration all-reduce, the rank-count-invariant ordered reduction, the per-pass histogram
exchange) run under `gloo` on CPU with world_size 2.  It restates what the CUDA kernels of
ood_in_object_detection_b200/csrc/{kmeans,fit_kernels}.cu compute, with the same block structure; the product never
imports it.
"""
from __future__ import annotations

import numpy as np
import torch


class NumpyBackend:
    def __init__(self, device="cpu"):
        self.device = torch.device(device)

    # squared distances of candidate seeds, float64 expansion cast to float32 (kmeans.cu::sqdist_cand_kernel)
    def sqdist_cand(self, x, seg_off_d, max_seg_rows, cand, closest):
        xs = x.numpy().astype(np.float64)
        off = seg_off_d.numpy()
        n_seg, n_cand = cand.shape[0], cand.shape[1]
        out = np.zeros((n_cand, xs.shape[0]), np.float32)
        pot = np.zeros((n_seg, n_cand), np.float64)
        cl = closest.numpy() if closest is not None else None
        for g in range(n_seg):
            a, b = int(off[g]), int(off[g + 1])
            if b <= a:
                continue
            y = cand[g].numpy().astype(np.float64)
            d = (-2.0 * xs[a:b] @ y.T + (y * y).sum(1)[None, :] + (xs[a:b] * xs[a:b]).sum(1)[:, None]).astype(np.float32)
            d = np.maximum(d, 0)
            if cl is not None:
                d = np.minimum(d, cl[a:b, None])
            out[:, a:b] = d.T
            pot[g] = d.astype(np.float64).sum(0)
        return torch.from_numpy(out), torch.from_numpy(pot)

    # ---- mean-centring stand-ins (kmeans.cu::seg_colsum_kernel / seg_center_kernel) ----
    def colsum(self, x, seg_off_d, n_seg):
        xs, off = x.numpy(), seg_off_d.numpy()
        out = np.zeros((n_seg, xs.shape[1]), np.float64)
        for g in range(n_seg):
            out[g] = xs[off[g]:off[g + 1]].sum(0, dtype=np.float64)
        return torch.from_numpy(out)

    def center(self, x, seg_off_d, n_seg, mean):
        xs, off, m = x.numpy(), seg_off_d.numpy(), mean.numpy()
        out = np.empty_like(xs)
        sq = np.zeros(n_seg, np.float64)
        for g in range(n_seg):
            out[off[g]:off[g + 1]] = xs[off[g]:off[g + 1]] - m[g]
            sq[g] = (out[off[g]:off[g + 1]].astype(np.float64) ** 2).sum()
        return torch.from_numpy(out), torch.from_numpy(sq)

    # ---- device seeding stand-ins (csrc/seed.cu) ----
    supports_device_seeding = True

    def seed_sqdist(self, x, seg_off_d, max_seg_rows, cand, closest):
        out, pot = self.sqdist_cand(x, seg_off_d, max_seg_rows, cand, closest)
        return out, pot

    def seed_scan(self, closest_all, piece_off, piece_cnt, uniform, pot, seg_trials, seg_on, max_seg_rows_global, cand_id):
        flat = closest_all.numpy().reshape(-1)
        po, pc = piece_off.numpy(), piece_cnt.numpy()
        u, pt, tr, on, out = uniform.numpy(), pot.numpy(), seg_trials.numpy(), seg_on.numpy(), cand_id.numpy()
        for g in range(po.shape[0]):
            if not on[g]:
                continue
            cd = np.concatenate([flat[po[g, r]:po[g, r] + pc[g, r]] for r in range(po.shape[1])])
            if not cd.size:
                continue
            vals = u[g, :tr[g]] * pt[g]                                    # float64 * float32 -> float64
            ids = np.searchsorted(np.cumsum(cd.astype(np.float32)), vals)
            np.clip(ids, None, cd.size - 1, out=ids)
            out[g, :tr[g]] = ids
            out[g, tr[g]:] = ids[0]

    def seed_gather(self, x, cand_id, seg_off_d, shard_first):
        xs, ids, off, sf = x.numpy(), cand_id.numpy(), seg_off_d.numpy(), shard_first.numpy()
        vec = np.zeros(ids.shape + (xs.shape[1],), np.float32)
        for g in range(ids.shape[0]):
            for j in range(ids.shape[1]):
                i = ids[g, j] - sf[g]
                if 0 <= i < off[g + 1] - off[g]:
                    vec[g, j] = xs[off[g] + i]
        return torch.from_numpy(vec)

    def seed_pick(self, pots, seg_trials, seg_on, seg_off_d, max_seg_rows, newd, vec, closest, pot, cent, c):
        ps, tr, on, off = pots.numpy(), seg_trials.numpy(), seg_on.numpy(), seg_off_d.numpy()
        nd, cl, pt = newd.numpy(), closest.numpy(), pot.numpy()
        for g in range(ps.shape[0]):
            if not on[g]:
                continue
            best = int(np.argmin(ps[g, :tr[g]].astype(np.float32)))
            cl[off[g]:off[g + 1]] = nd[best, off[g]:off[g + 1]]
            pt[g] = np.float32(ps[g, best])
            cent[g, c] = vec[g, best]

    # Lloyd assignment + per-block partial sums (kmeans.cu::kmeans_step_kernel)
    def step(self, x, k, seg_k, cent, blocks, active, labels, n_changed, update):
        xs, dim = x.numpy(), x.shape[1]
        nb = blocks.n_blocks
        psums = np.zeros((nb, k, dim), np.float32) if update else None
        pcounts = np.zeros((nb, k), np.float32) if update else None
        seg, r0, r1 = blocks.seg.numpy(), blocks.row0.numpy(), blocks.row1.numpy()
        lab, chg = labels.numpy(), n_changed.numpy()
        act = active.numpy() if active is not None else None
        for b in range(nb):
            g = int(seg[b])
            if act is not None and not act[g]:
                continue
            kg = int(seg_k[g])
            c = cent[g, :kg].numpy()
            rows = xs[r0[b]:r1[b]]
            if update == 2:
                new = np.clip(lab[r0[b]:r1[b]], 0, kg - 1)
            else:
                d = (c * c).sum(1)[None, :] - 2.0 * rows @ c.T
                new = d.argmin(1).astype(np.int32)
            chg[g] += int((lab[r0[b]:r1[b]] != new).sum())
            lab[r0[b]:r1[b]] = new
            if update:
                for j in range(kg):
                    m = new == j
                    psums[b, j] = rows[m].sum(0, dtype=np.float32) if m.any() else 0
                    pcounts[b, j] = m.sum()
        t = lambda a: None if a is None else torch.from_numpy(a)
        return t(psums), t(pcounts)

    def reduce(self, part, first, n_groups):
        p, f = part.numpy(), first.numpy()
        out = np.zeros((n_groups,) + p.shape[1:], np.float32)
        for g in range(n_groups):
            for b in range(int(f[g]), int(f[g + 1])):         # fixed order, like kmeans_reduce_kernel
                out[g] += p[b]
        return torch.from_numpy(out)

    def converge(self, n_changed, shift, n_empty, tol_abs, cnts, k, active, state, counts, any_active):
        """kmeans.cu::kmeans_converge_kernel: retire converged segments, count iterations (in place)."""
        chg, sh, ne, tol = n_changed.numpy(), shift.numpy(), n_empty.numpy(), tol_abs.numpy()
        act, st, cn, cs = active.numpy(), state.numpy(), counts.numpy(), cnts.numpy()
        any_ = 0
        for g in range(act.shape[0]):
            if not act[g]:
                continue
            st[0, g] += 1
            st[3, g] += ne[g]
            cn[g] = cs[g]
            if chg[g] == 0:
                st[1, g], act[g] = 1, 0
            elif float(sh[g]) <= tol[g]:
                st[2, g], act[g] = 1, 0
            else:
                any_ = 1
        any_active.numpy()[0] = any_

    def update(self, sums, counts, cent, seg_k, active, out=None):
        s, c, old = sums.numpy(), counts.numpy(), cent.numpy()
        new = old.copy()
        n_seg, k, _ = old.shape
        shift = np.zeros(n_seg, np.float32)
        n_empty = np.zeros(n_seg, np.int32)
        act = active.numpy() if active is not None else np.ones(n_seg, np.int32)
        for g in range(n_seg):
            if not act[g]:
                continue
            for j in range(int(seg_k[g])):
                if c[g, j] > 0:
                    new[g, j] = s[g, j] * np.float32(1.0 / float(c[g, j]))
                else:
                    n_empty[g] += 1
                shift[g] += ((new[g, j] - old[g, j]) ** 2).sum(dtype=np.float32)
        return torch.from_numpy(new), torch.from_numpy(shift), torch.from_numpy(n_empty)

    # one radix-select pass (fit_kernels.cu::radix_hist_kernel)
    def radix_hist(self, scores, seg_off, prefix, shift, bits, hist, minmax):
        bts = scores.numpy().view(np.uint32)
        key = np.where(bts & np.uint32(0x80000000), ~bts, bts | np.uint32(0x80000000)).astype(np.uint32)
        h = hist.numpy()
        pf = prefix.numpy().view(np.uint32)
        for g in range(len(seg_off) - 1):
            kk = key[seg_off[g]:seg_off[g + 1]]
            if not len(kk):
                continue
            if minmax is not None:
                mm = minmax.numpy().view(np.uint32)
                mm[g, 0], mm[g, 1] = min(mm[g, 0], kk.min()), max(mm[g, 1], kk.max())
            if shift + bits < 32:
                kk = kk[(kk >> np.uint32(shift + bits)) == pf[g]]
            h[g] += np.bincount((kk >> np.uint32(shift)) & np.uint32((1 << bits) - 1), minlength=1 << bits).astype(h.dtype)
