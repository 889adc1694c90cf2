"""Segmented k-means fit on the GPU: every (class, stride) problem advances in the same kernel launches.

Replaces `KMeans(n_clusters=k, random_state=10).fit_predict(X)` (/root/reference/cluster_utils.py:62-73),
called once per (class, stride) by /root/reference/ood_utils.py:2345.  What runs where:

  device (liboodb200.so)   distance passes of k-means++ (`oodb200_sqdist_cand_f32`), Lloyd assignment + block
                           partial sums (`oodb200_kmeans_step_f32`), fixed-order reduction, centre update
  host (numpy)             the scalar decisions of sklearn's `_kmeans_plusplus` (RandomState stream, float32
                           cumsum + searchsorted, candidate potentials), written with sklearn's own expressions so
                           that the chosen seeds are the ones sklearn picks; convergence bookkeeping
  torch.distributed        N > 1: rows of every segment are block-sharded across ranks; per Lloyd iteration ONE
                           all-reduce of [n_seg, K, D] sums + [n_seg, K] counts (+ changed-label counts), or, with
                           reduce="ordered", an all-gather of super-block partials summed in a fixed order so
                           that 1/2/4/8-rank runs give identical bits.

All tensor arithmetic on N rows happens in the CUDA kernels; torch is used for memory and collectives.  The
`backend` indirection exists so that the distributed control flow can be exercised on CPU (gloo) in the tests.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

BLOCK_ROWS = 512          # rows per CTA partial (the fixed partition of the rank-count-invariant "ordered" reduction)
BIG_BLOCK_ROWS = 1024     # ... of large fits in all-reduce mode: a CTA's set-up (centroid image, TMEM, barriers) is paid once per
                          # block, and 1024-row blocks make the C3 step 6.6 % faster (3.52 -> 3.29 ms); not below ~6 waves of CTAs
SUPER_ROWS = 4096         # unit of row ownership, whatever the block size
SUPER_BLOCKS = SUPER_ROWS // BLOCK_ROWS          # block partials per super-block (unit of ownership; 4096 rows: C3's 200 000-row segments split
                          # into 49 units, so 8 ranks own 6-7 each)
POLL_LAG = 2              # Lloyd iterations the device may run ahead of the host's convergence poll (1 when an iteration is long:
                          # see kmeans_fit)


def range_index(seg: int, world: int, rank: int) -> int:
    """Which of the `world` contiguous row ranges of segment `seg` the rank owns.  The assignment rotates with the
    segment index: floor(n_super * r / world) gives some range indices one super-block more than others, and without
    the rotation the same ranks would get the larger range of EVERY segment (C3 on 8 ranks: 2 x the rows of their
    neighbours)."""
    return (rank - seg) % world


def ranks_in_row_order(seg: int, world: int) -> List[int]:
    """Ranks owning the consecutive row ranges of segment `seg`, in row order."""
    return [(seg + j) % world for j in range(world)]
DEVICE_SEEDING_MIN_ROWS = 1 << 16   # seeding="auto": below this the host loop costs nothing and tracks sklearn's BLAS


_empty_stub = {}


def _ptr(t):
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


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class CudaBackend:
    """Device steps of the fit, through the C ABI."""

    def __init__(self, device):
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        self.device = device
        # OODB200_KMEANS_TC=0 keeps the Lloyd step on the FP32 kernel (A/B runs); default: tcgen05 where the shape fits
        self.tensor_core = os.environ.get("OODB200_KMEANS_TC", "1") != "0"

    def sqdist_cand(self, x, seg_off_d, max_seg_rows, cand, closest):
        n_seg, n_cand = cand.shape[0], cand.shape[1]
        cand = cand.contiguous()
        n = x.shape[0]
        out = torch.empty((n_cand, n), dtype=torch.float32, device=x.device)
        pot = torch.zeros((n_seg, n_cand), dtype=torch.float64, device=x.device)
        self._lib.check(self.lib.oodb200_sqdist_cand_f32(_ptr(x), x.shape[1], _ptr(seg_off_d), n_seg, int(max_seg_rows),
                                                         _ptr(cand), n_cand, _ptr(closest), _ptr(out), _ptr(pot),
                                                         _stream()), "oodb200_sqdist_cand_f32")
        return out, pot

    # ---- mean-centring (csrc/kmeans.cu) ----
    def colsum(self, x, seg_off_d, n_seg):
        dim = x.shape[1]
        scratch = torch.empty(int(self.lib.oodb200_segment_scratch_doubles(n_seg, dim)), dtype=torch.float64, device=x.device)
        sums = torch.zeros((n_seg, dim), dtype=torch.float64, device=x.device)
        self._lib.check(self.lib.oodb200_segment_colsum_f64(_ptr(x), dim, _ptr(seg_off_d), n_seg, _ptr(scratch), _ptr(sums),
                                                            _stream()), "oodb200_segment_colsum_f64")
        return sums

    def center(self, x, seg_off_d, n_seg, mean):
        dim = x.shape[1]
        scratch = torch.empty(int(self.lib.oodb200_segment_scratch_doubles(n_seg, dim)), dtype=torch.float64, device=x.device)
        out = torch.empty_like(x)
        sq = torch.zeros(n_seg, dtype=torch.float64, device=x.device)
        self._lib.check(self.lib.oodb200_segment_center_f32(_ptr(x), dim, _ptr(seg_off_d), n_seg, _ptr(mean), _ptr(out),
                                                            _ptr(scratch), _ptr(sq), _stream()), "oodb200_segment_center_f32")
        return out, sq

    # ---- device seeding (csrc/seed.cu) ----
    supports_device_seeding = True

    def seed_sqdist(self, x, seg_off_d, max_seg_rows, cand, closest):
        """-> (min(closest, d) [n_cand, n_local], potentials [n_seg, n_cand] float64, bit-reproducible)"""
        n_seg, n_cand = cand.shape[0], cand.shape[1]
        gx = int(self.lib.oodb200_seed_grid(int(max_seg_rows)))
        out = torch.empty((n_cand, x.shape[0]), dtype=torch.float32, device=x.device)
        part = torch.empty((n_seg, gx, 4), dtype=torch.float64, device=x.device)
        pots = torch.empty((n_seg, n_cand), dtype=torch.float64, device=x.device)
        self._lib.check(self.lib.oodb200_seed_sqdist_f32(_ptr(x), x.shape[1], _ptr(seg_off_d), n_seg, int(max_seg_rows),
                                                         _ptr(cand), n_cand, _ptr(closest), _ptr(out), _ptr(part),
                                                         _ptr(pots), _stream()), "oodb200_seed_sqdist_f32")
        return out, pots

    def seed_scan(self, closest_all, piece_off, piece_cnt, uniform, pot, seg_trials, seg_on, max_seg_rows_global, cand_id):
        n_seg, n_pieces = piece_off.shape
        n_trials = uniform.shape[1]
        max_chunks = 32 * ((int(max_seg_rows_global) + 4095) // 4096) + 32
        key = (n_seg, max_chunks, str(closest_all.device))
        if getattr(self, "_skey", None) != key:
            self._sbuf = torch.empty((n_seg, max_chunks), dtype=torch.float32, device=closest_all.device)
            self._skey = key
        self._lib.check(self.lib.oodb200_seed_scan_f32(_ptr(closest_all), _ptr(piece_off), _ptr(piece_cnt), n_seg, n_pieces,
                                                       _ptr(uniform), _ptr(pot), _ptr(seg_trials), _ptr(seg_on), n_trials,
                                                       _ptr(self._sbuf), max_chunks, _ptr(cand_id), _stream()),
                        "oodb200_seed_scan_f32")

    def seed_gather(self, x, cand_id, seg_off_d, shard_first):
        n_seg, n_cand = cand_id.shape
        vec = torch.empty((n_seg, n_cand, x.shape[1]), dtype=torch.float32, device=x.device)
        self._lib.check(self.lib.oodb200_seed_gather_f32(_ptr(x), x.shape[1], _ptr(cand_id), n_seg, n_cand, _ptr(seg_off_d),
                                                         _ptr(shard_first), _ptr(vec), _stream()), "oodb200_seed_gather_f32")
        return vec

    def seed_pick(self, pots, seg_trials, seg_on, seg_off_d, max_seg_rows, newd, vec, closest, pot, cent, c):
        n_seg, n_cand = pots.shape
        k, dim = cent.shape[1], cent.shape[2]
        out = C.c_void_p(cent.data_ptr() + 4 * c * dim)
        self._lib.check(self.lib.oodb200_seed_pick_f32(_ptr(pots), _ptr(seg_trials), _ptr(seg_on), n_seg, n_cand,
                                                       _ptr(seg_off_d), int(max_seg_rows), _ptr(newd), _ptr(vec), dim,
                                                       _ptr(closest), _ptr(pot), out, k * dim, None, _stream()),
                        "oodb200_seed_pick_f32")

    def _partials(self, n_blocks, k, dim, device):
        """Block-partial buffers, reused across the Lloyd iterations of a fit (144 MB at C3/2: not worth re-allocating)."""
        key = (n_blocks, k, dim, str(device))
        if getattr(self, "_pkey", None) != key:
            self._pbuf = (torch.empty((n_blocks, k, dim), dtype=torch.float32, device=device),
                          torch.empty((n_blocks, k), dtype=torch.float32, device=device))
            self._pkey = key
        return self._pbuf

    def step(self, x, k, seg_k, cent, blocks, active, labels, n_changed, update):
        n_blocks = blocks.n_blocks
        psums, pcounts = self._partials(n_blocks, k, x.shape[1], x.device) if update else (None, None)
        n_seg = int(seg_k.shape[0])
        ws_bytes = int(self.lib.oodb200_kmeans_tc_workspace_bytes(n_seg, k, x.shape[1])) if self.tensor_core else 0
        if ws_bytes and update in (0, 1, True, False) and n_blocks and x.data_ptr() % 16 == 0:
            # tcgen05 path (csrc/kmeans_tc.cu): cross-term on the tensor pipe, rows read once by TMA
            key = (ws_bytes, str(x.device))
            if getattr(self, "_wkey", None) != key:
                self._wbuf = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
                self._wkey = key
            self._lib.check(self.lib.oodb200_kmeans_step_tc_f32(
                _ptr(x), int(x.shape[0]), x.shape[1], n_seg, k, _ptr(seg_k), _ptr(cent), _ptr(blocks.seg), _ptr(blocks.row0),
                _ptr(blocks.row1), n_blocks, _ptr(active), _ptr(labels), _ptr(psums), _ptr(pcounts), _ptr(n_changed),
                int(update), _ptr(self._wbuf), _stream()), "oodb200_kmeans_step_tc_f32")
            return psums, pcounts
        self._lib.check(self.lib.oodb200_kmeans_step_f32(
            _ptr(x), x.shape[1], int(seg_k.shape[0]), k, _ptr(seg_k), _ptr(cent), _ptr(blocks.seg), _ptr(blocks.row0),
            _ptr(blocks.row1), n_blocks, _ptr(active), _ptr(labels), _ptr(psums), _ptr(pcounts), _ptr(n_changed),
            int(update), _stream()), "oodb200_kmeans_step_f32")
        return psums, pcounts

    def reduce(self, part, first, n_groups):
        elems = int(np.prod(part.shape[1:]))
        out = torch.empty((n_groups,) + tuple(part.shape[1:]), dtype=torch.float32, device=part.device)
        self._lib.check(self.lib.oodb200_kmeans_reduce_f32(_ptr(part), _ptr(first), n_groups, elems, _ptr(out), _stream()),
                        "oodb200_kmeans_reduce_f32")
        return out

    def reduce_into(self, part, first, n_groups, out):
        """kmeans_reduce into a caller-owned contiguous buffer (a slice of the all-reduce buffer: no concatenation)."""
        elems = int(np.prod(part.shape[1:]))
        self._lib.check(self.lib.oodb200_kmeans_reduce_f32(_ptr(part), _ptr(first), n_groups, elems, _ptr(out), _stream()),
                        "oodb200_kmeans_reduce_f32")
        return out

    def reduce_step(self, psums, pcounts, first, n_groups, out_s, out_c, n_changed, chg_f):
        """Both reductions of a Lloyd iteration + the changed-label counters (to float32, then cleared) in one launch."""
        self._lib.check(self.lib.oodb200_kmeans_reduce_step_f32(
            _ptr(psums), _ptr(pcounts), _ptr(first), n_groups, int(np.prod(psums.shape[1:])), int(np.prod(pcounts.shape[1:])),
            _ptr(out_s), _ptr(out_c), _ptr(n_changed), _ptr(chg_f), _stream()), "oodb200_kmeans_reduce_step_f32")

    def update(self, sums, counts, cent, seg_k, active, out=None):
        n_seg, k, dim = cent.shape
        if out is None:
            out = (torch.empty_like(cent), torch.empty(n_seg, dtype=torch.float32, device=cent.device),
                   torch.empty(n_seg, dtype=torch.int32, device=cent.device))
        new, shift, n_empty = out
        self._lib.check(self.lib.oodb200_kmeans_update_f32(_ptr(sums), _ptr(counts), _ptr(cent), _ptr(seg_k), _ptr(active),
                                                           n_seg, k, dim, _ptr(new), _ptr(shift), _ptr(n_empty), _stream()),
                        "oodb200_kmeans_update_f32")
        return new, shift, n_empty

    def update_peers(self, peers, slot, n_sum, n_cnt, cent, seg_k, active, out, cnts_out, chg_out):
        """Centre update with the iteration's all-reduce fused in: the partial sums / counts / changed-label counts of every
        rank are read from the peers' symmetric buffers over NVLink and added in rank order
        (csrc/kmeans.cu::kmeans_update_kernel, n_peers > 0)."""
        n_seg, k, dim = cent.shape
        new, shift, n_empty = out
        peers["epoch"] = (peers["epoch"] + 1) & 0xFFFFFFFF                    # same sequence on every rank: the barrier's ticket
        self._lib.check(self.lib.oodb200_kmeans_update_peers_f32(
            C.c_void_p(peers["ptrs"][slot]), peers["world"], int(n_sum), int(n_sum + n_cnt), C.c_void_p(peers["flag_ptrs"]),
            peers["rank"], peers["epoch"], _ptr(cent), _ptr(seg_k), _ptr(active), n_seg, k, dim, _ptr(new), _ptr(shift),
            _ptr(n_empty), _ptr(cnts_out), _ptr(chg_out), _stream()), "oodb200_kmeans_update_peers_f32")
        return new, shift, n_empty

    def peer_buffers(self, numel, group):
        """Two symmetric buffers of `numel` floats (alternating between Lloyd iterations) shared with the ranks of `group`
        over NVLink / NVSwitch peer memory, or None when symmetric memory is unavailable (then NCCL reduces).  Cached: the
        rendezvous exchanges memory handles between the ranks."""
        import torch.distributed as dist
        key = (int(numel), id(group))
        hit = getattr(self, "_peers", {}).get(key)
        if hit is not None or key in getattr(self, "_peers", {}):
            return hit
        self._peers = getattr(self, "_peers", {})
        ok = torch.zeros(1, dtype=torch.int32, device=self.device)
        bufs = None
        try:
            # opt-in (OODB200_KMEANS_PEERS=1): measured on 2 x B200 the fused path is SLOWER end to end than NCCL's all-reduce
            # (5.3 vs 2.0 ms per iteration at C3 / 2 although its kernels take 30 us: DESIGN.md section 6), so NCCL is the default
            if os.environ.get("OODB200_KMEANS_PEERS", "0") != "1" or dist.get_backend(group) != "nccl":
                raise RuntimeError("disabled")
            import torch.distributed._symmetric_memory as symm
            t = [symm.empty(int(numel), dtype=torch.float32, device=self.device) for _ in range(2)]
            flags = symm.empty(64, dtype=torch.int32, device=self.device)      # the in-kernel barrier's tickets, one slot per rank
            h = [symm.rendezvous(b, group) for b in t]
            hf = symm.rendezvous(flags, group)
            for b in t:
                b.zero_()
            flags.zero_()
            bufs = dict(t=t, h=h, flags=flags, hf=hf, ptrs=[int(x.buffer_ptrs_dev) for x in h], flag_ptrs=int(hf.buffer_ptrs_dev),
                        world=int(h[0].world_size), rank=int(h[0].rank), epoch=0)
            ok.fill_(1)
        except Exception as e:                                    # every rank must take the same path: agree below
            if str(e) != "disabled":
                import sys
                print(f"kmeans: symmetric memory unavailable ({type(e).__name__}: {e}); the Lloyd all-reduce stays on NCCL", file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)              # also: every rank has zeroed its tickets before anyone signals
        if int(ok.item()) == 0:
            bufs = None
        self._peers[key] = bufs
        return bufs

    def converge(self, n_changed, shift, n_empty, tol_abs, cnts, k, active, state, counts, any_active):
        """Device-side convergence bookkeeping of one iteration (csrc/kmeans.cu::kmeans_converge_kernel)."""
        n_seg = int(active.shape[0])
        as_int = n_changed.dtype == torch.int32
        self._lib.check(self.lib.oodb200_kmeans_converge_f32(
            _ptr(n_changed) if as_int else None, None if as_int else _ptr(n_changed), _ptr(shift), _ptr(n_empty), _ptr(tol_abs),
            _ptr(cnts), n_seg, k, _ptr(active), _ptr(state), _ptr(counts), _ptr(any_active), _stream()),
            "oodb200_kmeans_converge_f32")


@dataclass
class BlockTable:
    """Fixed partition of the (global) row space of every segment into blocks and super-blocks.
    The partition depends only on the GLOBAL segment sizes, never on the number of ranks."""
    seg: torch.Tensor          # [n_blocks] int32 (local blocks)
    row0: torch.Tensor         # [n_blocks] int64 local row range
    row1: torch.Tensor
    n_blocks: int
    super_first: torch.Tensor  # [n_super_local+1] int32: local blocks per local super-block
    n_super_local: int
    super_seg_first: torch.Tensor  # [n_seg+1] int32 over ALL super-blocks (global, rank-major = row order)
    n_super_global: int
    super_owner_counts: List[int]  # super-blocks per rank
    rot: List[int] = field(default_factory=list)   # rotation key of every segment (range_index / ranks_in_row_order)


_BLOCK_CACHE: dict = {}


def auto_block_rows(global_sizes: Sequence[int], world: int, reduce: str = "allreduce") -> int:
    """Rows per CTA block of a fit: BIG_BLOCK_ROWS when every rank still gets >= 6 waves of CTAs on 148 SMs and the result need
    not be identical for every rank count (reduce != "ordered"), else BLOCK_ROWS."""
    per_rank = sum(int(n) for n in global_sizes) // max(int(world), 1)
    return BIG_BLOCK_ROWS if reduce != "ordered" and per_rank >= 6 * 148 * BIG_BLOCK_ROWS else BLOCK_ROWS


def build_blocks(global_sizes: Sequence[int], world: int, rank: int, device, rot: Optional[Sequence[int]] = None,
                 block_rows: Optional[int] = None) -> tuple:
    """Row sharding + block tables.  Rank r owns a contiguous range of super-blocks of every segment.
    rot[g]: rotation key of segment g (default: its index) -- callers that fit a SUBSET of their segments pass a stable id
    (the class index) so that the rows a rank must hold do not depend on which other segments take part.
    The tables only depend on (sizes, world, rank, rot): the last few are kept (a C3 table has 7.8 k blocks built in Python)."""
    rot = list(range(len(global_sizes))) if rot is None else [int(v) for v in rot]
    block_rows = int(block_rows or BLOCK_ROWS)
    assert SUPER_ROWS % block_rows == 0 and block_rows % 128 == 0, "block_rows: a multiple of 128 dividing 4096"
    key = (tuple(int(n) for n in global_sizes), int(world), int(rank), str(device), tuple(rot), block_rows)
    hit = _BLOCK_CACHE.get(key)
    if hit is not None:
        return hit
    out = _build_blocks(global_sizes, world, rank, device, rot, block_rows)
    if len(_BLOCK_CACHE) >= 8:
        _BLOCK_CACHE.pop(next(iter(_BLOCK_CACHE)))
    _BLOCK_CACHE[key] = out
    return out


def _build_blocks(global_sizes: Sequence[int], world: int, rank: int, device, rot, block_rows: int = BLOCK_ROWS) -> tuple:
    seg_l, r0_l, r1_l, sfirst = [], [], [], [0]
    local_sizes, local_off = [], [0]
    super_seg_first = [0]
    owner_counts = [0] * world
    shard = []                                           # per segment: (global start row, rows) owned by this rank
    for g, n in enumerate(global_sizes):
        n_super = (n + SUPER_ROWS - 1) // (SUPER_ROWS)
        bounds = [(n_super * r) // world for r in range(world + 1)]       # super-blocks per range, contiguous
        for r in range(world):
            ri = range_index(rot[g], world, r)
            owner_counts[r] += bounds[ri + 1] - bounds[ri]
        super_seg_first.append(super_seg_first[-1] + n_super)
        ri = range_index(rot[g], world, rank)
        s0, s1 = bounds[ri], bounds[ri + 1]
        row_a = min(n, s0 * SUPER_ROWS)
        row_b = min(n, s1 * SUPER_ROWS)
        shard.append((row_a, row_b - row_a))
        base = local_off[-1]
        for sb in range(s0, s1):
            a = sb * SUPER_ROWS
            b = min(n, a + SUPER_ROWS)
            for blk in range(a, b, block_rows):
                seg_l.append(g)
                r0_l.append(base + blk - row_a)
                r1_l.append(base + min(b, blk + block_rows) - row_a)
            sfirst.append(len(seg_l))
        local_sizes.append(row_b - row_a)
        local_off.append(base + row_b - row_a)
    t = lambda a, dt: torch.tensor(a, dtype=dt, device=device)
    table = BlockTable(seg=t(seg_l, torch.int32), row0=t(r0_l, torch.int64), row1=t(r1_l, torch.int64),
                       n_blocks=len(seg_l), super_first=t(sfirst, torch.int32), n_super_local=len(sfirst) - 1,
                       super_seg_first=t(super_seg_first, torch.int32), n_super_global=super_seg_first[-1],
                       super_owner_counts=owner_counts, rot=list(rot))
    return table, shard, local_off


@dataclass
class KMeansResult:
    labels: torch.Tensor            # [n_local] int32
    centers: torch.Tensor           # [n_seg, k, dim] float32 in the ORIGINAL coordinates
    counts: torch.Tensor            # [n_seg, k]
    n_iter: List[int]
    strict: List[bool]
    n_empty: List[int]
    seconds: dict = field(default_factory=dict)


def _sklearn_first_center(rs: np.random.RandomState, n: int) -> int:
    """`random_state.choice(n_samples, p=sample_weight / sample_weight.sum())` (_kmeans.py:234) with unit float32 weights.
    RandomState.choice(p=...) is `cdf = p.cumsum(); cdf /= cdf[-1]; cdf.searchsorted(random_sample(), side='right')` on
    the float64 copy of p; restated here without choice()'s O(n) validation passes (same stream consumption, same
    arithmetic; tests/test_host_logic.py holds it to rs.choice)."""
    if n >= 1 << 24:                                   # float32 ones no longer sum exactly: take numpy's own path
        w = np.ones(n, dtype=np.float32)
        return int(rs.choice(n, p=w / w.sum()))
    pv = np.float64(np.float32(1.0) / np.float32(n))
    cdf = np.cumsum(np.full(n, pv, dtype=np.float64))
    cdf /= cdf[-1]
    return int(cdf.searchsorted(rs.random_sample(), side="right"))


def kmeans_fit(x_local: torch.Tensor, global_sizes: Sequence[int], k: int, table: BlockTable, local_off: Sequence[int],
               shard: Sequence[tuple], random_state: int = 10, max_iter: int = 300, tol: float = 1e-4,
               backend=None, group=None, reduce: str = "allreduce", seeding: str = "auto",
               seg_k: Optional[Sequence[int]] = None) -> KMeansResult:
    """x_local: this rank's rows [n_local, dim] (segment-major, each segment's shard contiguous), float32.
    global_sizes: rows per segment over all ranks.  Returns labels for the local rows and the global centres.
    seg_k: clusters per segment (each <= k; default k everywhere): every segment is an independent
    `KMeans(n_clusters=seg_k[g], random_state=random_state)` -- the k-search fits all its candidates in one call.

    seeding: "device" -- the whole k-means++ loop runs on the device (csrc/seed.cu): exact sequential float32 cumsum +
             searchsorted, float64 candidate distances, potentials = correctly rounded float32 of the exact sum;
             "host"   -- the scalar decisions run in numpy with sklearn's own expressions (the potentials come from the
             host BLAS `closest @ ones`, like sklearn on this machine), O(N) host work per centre;
             "auto"   -- "device" from DEVICE_SEEDING_MIN_ROWS total rows on, else "host".
    The two differ only in the last bit of a potential, which moves a draw to a neighbouring row with probability
    ~ n * 2^-24 per draw (sklearn itself depends on the BLAS summation order there)."""
    import time
    import torch.distributed as dist
    distributed = group is not None                    # explicit: a single-process fit inside an initialised job stays local
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    dev = x_local.device
    backend = backend or CudaBackend(dev)
    n_seg, dim = len(global_sizes), int(x_local.shape[1])
    if seg_k is not None and (len(seg_k) != len(global_sizes) or any(not 0 < int(kk) <= k for kk in seg_k)):
        raise ValueError("seg_k: one value in 1..k per segment expected")
    seg_k_host = [min(k if seg_k is None else int(seg_k[g]), int(n)) for g, n in enumerate(global_sizes)]
    seg_k = torch.tensor(seg_k_host, dtype=torch.int32, device=dev)
    seg_off_d = torch.tensor(list(local_off), dtype=torch.int64, device=dev)
    timing = {}
    t0 = time.perf_counter()

    def allreduce(t, op=None):
        if distributed:
            dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=group)
        return t

    # ---- mean-centre every segment (KMeans.fit: X -= X.mean(axis=0)) and the sklearn tolerance ----
    x_local = x_local.contiguous()
    mean = allreduce(backend.colsum(x_local, seg_off_d, n_seg))            # float64 column sums, one streaming pass
    gs = torch.tensor([max(int(n), 1) for n in global_sizes], dtype=torch.float64, device=dev)
    mean = (mean / gs[:, None]).to(torch.float32).contiguous()
    x, var = backend.center(x_local, seg_off_d, n_seg, mean)                # x - mean and sum of squares, second pass
    allreduce(var)
    tol_abs = (var / gs / dim * tol).cpu().numpy()                         # mean over features of the variance
    if distributed:
        torch.cuda.synchronize() if dev.type == "cuda" else None
    timing["center"] = time.perf_counter() - t0

    # ---- k-means++ seeding, all segments in lock-step (sklearn _kmeans_plusplus, _kmeans.py:180-278) ----
    t0 = time.perf_counter()
    if seeding == "auto":
        seeding = "device" if (getattr(backend, "supports_device_seeding", False)
                               and sum(int(n) for n in global_sizes) >= DEVICE_SEEDING_MIN_ROWS) else "host"
    if seeding not in ("device", "host"):
        raise ValueError(f"seeding must be 'auto', 'device' or 'host', not {seeding!r}")
    if seeding == "device":
        cent = _seed_on_device(x, global_sizes, local_off, shard, seg_k_host, k, random_state, backend, group,
                               distributed, world, table.rot)
        if dev.type == "cuda":
            torch.cuda.synchronize()
    else:
        cent = _seed_on_host(x, global_sizes, local_off, shard, seg_k_host, k, random_state, backend, group,
                             distributed, world, seg_off_d, table.rot)
    timing["init"] = time.perf_counter() - t0
    timing["seeding"] = seeding

    # ---- Lloyd iterations (sklearn _kmeans_single_lloyd, _kmeans.py:630-758) ----
    # Everything an iteration decides stays on the device: the convergence kernel retires segments (active flags, iteration
    # counts, strict / needs-final-E-step marks) and the host only polls ONE flag, POLL_LAG iterations behind the launches,
    # so that the device never waits for the host loop.  Iterations issued after every segment has stopped are no-ops
    # (inactive segments are skipped by the step, carried through by the update, ignored by the bookkeeping); with N > 1
    # every rank sees the same all-reduced flags and therefore issues the same number of collectives.
    t0 = time.perf_counter()
    cuda = dev.type == "cuda"
    labels = torch.full((x.shape[0],), -1, dtype=torch.int32, device=dev)
    active = torch.tensor([1 if global_sizes[g] > 0 else 0 for g in range(n_seg)], dtype=torch.int32, device=dev)
    state = torch.zeros((4, n_seg), dtype=torch.int32, device=dev)         # iterations, strict, needs final E-step, empties
    any_active = torch.ones(1, dtype=torch.int32, device=dev)
    tol_d = torch.from_numpy(np.ascontiguousarray(tol_abs, dtype=np.float64)).to(dev)
    seg_first = _seg_first_local(table, n_seg, dev)
    counts = torch.zeros((n_seg, k), dtype=torch.float32, device=dev)
    n_sum, n_cnt = n_seg * k * dim, n_seg * k
    flat = torch.zeros(n_sum + n_cnt + n_seg, dtype=torch.float32, device=dev)   # sums | counts | changed labels: ONE all-reduce
    sums_v, cnts_v = flat[:n_sum].view(n_seg, k, dim), flat[n_sum:n_sum + n_cnt].view(n_seg, k)
    chg_f = flat[n_sum + n_cnt:]
    n_changed = torch.zeros(n_seg, dtype=torch.int32, device=dev)
    cent = cent.contiguous()
    other = (torch.empty_like(cent), torch.empty(n_seg, dtype=torch.float32, device=dev),
             torch.empty(n_seg, dtype=torch.int32, device=dev))
    has_into = hasattr(backend, "reduce_into")
    # N > 1 on one NVLink node: the all-reduce is fused into the centre update -- every rank writes its reduced partials into a
    # symmetric buffer and the update kernel adds the peers' values in rank order (kmeans_update_peers); else NCCL.
    peers = backend.peer_buffers(flat.numel(), group) if (distributed and cuda and reduce != "ordered" and has_into
                                                          and hasattr(backend, "peer_buffers")) else None
    if peers is not None:
        cnts_sum = torch.zeros((n_seg, k), dtype=torch.float32, device=dev)
        chg_sum = torch.zeros(n_seg, dtype=torch.float32, device=dev)
    timing["collective"] = "peer-memory update" if peers is not None else ("nccl all-reduce" if distributed else "none")
    polls = []                                                              # (host flag, event) per issued iteration
    # every iteration issued past the converged one is a wasted step: one iteration of lag hides the host loop when an
    # iteration takes >= ~0.2 ms of device time (rows per rank x dim, the same figure on every rank), short ones need two
    lag = 1 if sum(int(n) for n in global_sizes) // world * dim >= 100_000_000 else POLL_LAG
    flag_ring = torch.empty(POLL_LAG + 2, dtype=torch.int32).pin_memory() if cuda else None
    issued = 0
    fused_reduce = hasattr(backend, "reduce_step") and reduce != "ordered" and peers is None and table.n_blocks > 0
    for it in range(max_iter):
        if not fused_reduce:
            n_changed.zero_()
        psums, pcounts = backend.step(x, k, seg_k, cent, table, active, labels, n_changed, True)
        if reduce == "ordered":
            sums, cnts = _ordered_reduce(backend, psums, pcounts, table, n_seg, world, group)
            chg = allreduce(n_changed) if distributed else n_changed
        elif peers is not None:
            buf = peers["t"][it & 1]                                        # alternate: nobody still reads the buffer written now
            b_sums, b_cnts = buf[:n_sum].view(n_seg, k, dim), buf[n_sum:n_sum + n_cnt].view(n_seg, k)
            if table.n_blocks:
                backend.reduce_into(psums, seg_first, n_seg, b_sums)
                backend.reduce_into(pcounts, seg_first, n_seg, b_cnts)
            else:
                buf[:n_sum + n_cnt].zero_()
            buf[n_sum + n_cnt:].copy_(n_changed)                            # exact in float32 below 2^24 rows per segment
            # the update kernel itself waits for every rank's partials (tickets in symmetric memory), then adds them over NVLink
            new_cent, shift, n_empty = backend.update_peers(peers, it & 1, n_sum, n_cnt, cent, seg_k, active, other, cnts_sum, chg_sum)
            backend.converge(chg_sum, shift, n_empty, tol_d, cnts_sum, k, active, state, counts, any_active)
        elif fused_reduce:
            backend.reduce_step(psums, pcounts, seg_first, n_seg, sums_v, cnts_v, n_changed, chg_f)   # also clears n_changed
            if distributed:
                allreduce(flat)                                             # the one collective of the iteration
            sums, cnts, chg = sums_v, cnts_v, chg_f
        else:
            if table.n_blocks and has_into:
                backend.reduce_into(psums, seg_first, n_seg, sums_v)
                backend.reduce_into(pcounts, seg_first, n_seg, cnts_v)
            elif table.n_blocks:
                sums_v.copy_(backend.reduce(psums, seg_first, n_seg))
                cnts_v.copy_(backend.reduce(pcounts, seg_first, n_seg))
            else:
                flat.zero_()
            sums, cnts, chg = sums_v, cnts_v, n_changed
            if distributed:
                chg_f.copy_(n_changed)                                      # exact in float32 below 2^24 rows per segment
                allreduce(flat)                                             # the one collective of the iteration
                chg = chg_f
        if peers is None or reduce == "ordered":
            new_cent, shift, n_empty = backend.update(sums, cnts, cent, seg_k, active, out=other)
            backend.converge(chg, shift, n_empty, tol_d, cnts, k, active, state, counts, any_active)
        other = (cent, shift, n_empty)
        cent = new_cent
        issued += 1
        if cuda:
            flag = flag_ring[it % (POLL_LAG + 2):it % (POLL_LAG + 2) + 1]     # consumed POLL_LAG iterations later: no reuse hazard
            flag.copy_(any_active, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            polls.append((flag, ev))
            if len(polls) > lag:
                f, e = polls[len(polls) - 1 - lag]
                e.synchronize()
                if int(f[0]) == 0:
                    break
        elif int(any_active[0]) == 0:
            break
    if cuda:
        torch.cuda.synchronize()
    st = state.cpu().numpy()
    n_iter = [int(v) for v in st[0]]
    strict = [bool(v) for v in st[1]]
    n_empty_tot = [int(v) for v in st[3]]
    need_final = st[2].astype(bool) | active.cpu().numpy().astype(bool)    # max_iter reached without convergence
    lloyd_iters = max(n_iter + [0])
    if need_final.any():                                                    # E-step with the final centres (:742-754)
        fin = torch.from_numpy(need_final.astype(np.int32)).to(dev)
        dummy = torch.zeros(n_seg, dtype=torch.int32, device=dev)
        backend.step(x, k, seg_k, cent, table, fin, labels, dummy, False)
    if cuda:
        torch.cuda.synchronize()
    timing["lloyd"] = time.perf_counter() - t0
    timing["lloyd_iters"] = lloyd_iters
    timing["lloyd_issued"] = issued
    return KMeansResult(labels=labels, centers=cent + mean[:, None, :], counts=counts.clone(), n_iter=n_iter, strict=strict,
                        n_empty=n_empty_tot, seconds=timing)


def _seed_on_host(x, global_sizes, local_off, shard, seg_k_host, k, random_state, backend, group, distributed, world,
                  seg_off_d, rot=None):
    """k-means++ with the scalar decisions in numpy (sklearn's own expressions, host BLAS potentials)."""
    import torch.distributed as dist
    dev = x.device
    n_seg, dim = len(global_sizes), int(x.shape[1])
    rot = list(range(n_seg)) if not rot else rot

    def allreduce(t, op=None):
        if distributed:
            dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=group)
        return t

    # sklearn: n_local_trials = 2 + int(log(n_clusters)) with the segment's OWN n_clusters = min(k, n)
    trials = [2 + int(np.log(kk)) if kk > 1 else 1 for kk in seg_k_host]
    n_trials = max(trials + [1])
    rs = [np.random.RandomState(random_state) for _ in range(n_seg)]
    cent = torch.zeros((n_seg, k, dim), dtype=torch.float32, device=dev)
    max_rows = max([local_off[g + 1] - local_off[g] for g in range(n_seg)] + [0])

    n_local = int(x.shape[0])
    local_sizes = [local_off[g + 1] - local_off[g] for g in range(n_seg)]
    if distributed:
        meta = torch.tensor([n_local] + local_sizes, dtype=torch.int64, device=dev)
        metas = [torch.empty_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta, group=group)
        metas = [m.cpu().numpy() for m in metas]
        max_local = max(int(m[0]) for m in metas)

    def gather_rows(arr_local):
        """local per-row values [m, n_local] -> list over segments of global arrays [m, n_g] (numpy, row order)."""
        if not distributed:
            a = arr_local.cpu().numpy()
            return [a[:, local_off[g]:local_off[g + 1]] for g in range(n_seg)]
        m = arr_local.shape[0]
        pad = torch.zeros((m, max_local), dtype=arr_local.dtype, device=dev)
        pad[:, :n_local] = arr_local
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)                            # equal-sized: works on nccl and gloo
        host = [b.cpu().numpy() for b in bufs]
        out = []
        for g in range(n_seg):
            parts = []
            for r in ranks_in_row_order(rot[g], world):
                o = int(metas[r][1:1 + g].sum())
                parts.append(host[r][:, o:o + int(metas[r][1 + g])])
            out.append(np.concatenate(parts, axis=1))
        return out

    def fetch_vectors(global_ids):
        """rows (global index within segment) [n_seg, m] -> centred vectors [n_seg, m, dim] on every rank."""
        m = global_ids.shape[1]
        vec = torch.zeros((n_seg, m, dim), dtype=torch.float32, device=dev)
        for g in range(n_seg):
            a, cnt = shard[g]
            for j in range(m):
                gid = int(global_ids[g, j])
                if a <= gid < a + cnt:
                    vec[g, j] = x[local_off[g] + gid - a]
        return allreduce(vec)                                               # exactly one rank contributes each row

    def cand_distances(vectors, closest):
        """min(closest, squared distance to each candidate vector) for the local rows: [m, n_local]."""
        d, _ = backend.sqdist_cand(x, seg_off_d, max_rows, vectors, closest)
        return d

    seg_of_row = torch.repeat_interleave(torch.arange(n_seg, device=dev), torch.tensor(local_sizes, device=dev))
    active_seg = [g for g in range(n_seg) if global_sizes[g] > 0]
    first = np.zeros((n_seg, 1), dtype=np.int64)
    for g in active_seg:
        first[g, 0] = _sklearn_first_center(rs[g], int(global_sizes[g]))
    v0 = fetch_vectors(first)
    cent[:, 0] = v0[:, 0]
    closest = cand_distances(v0, None)[0].contiguous()                      # [n_local]
    closest_g = gather_rows(closest[None])
    pot = [np.float32(0)] * n_seg
    for g in active_seg:
        pot[g] = closest_g[g] @ np.ones(closest_g[g].shape[1], dtype=np.float32)         # _kmeans.py:246, shape (1,)
    for c in range(1, k):
        cand = np.zeros((n_seg, n_trials), dtype=np.int64)
        for g in active_seg:
            if c >= seg_k_host[g]:
                continue
            cd = closest_g[g][0]
            rand_vals = rs[g].uniform(size=trials[g]) * pot[g]                             # _kmeans.py:252
            ids = np.searchsorted(np.cumsum(np.ones(cd.size, dtype=np.float32) * cd), rand_vals)
            np.clip(ids, None, cd.size - 1, out=ids)
            cand[g, :trials[g]] = ids
            cand[g, trials[g]:] = ids[0]                                                   # padding, ignored below
        vec = fetch_vectors(cand)
        newd = cand_distances(vec, closest)                                                # min(closest, d) on device
        newd_g = gather_rows(newd)
        best = np.zeros(n_seg, dtype=np.int64)
        for g in active_seg:
            if c >= seg_k_host[g]:
                continue
            nd = newd_g[g][:trials[g]]
            pots = nd @ np.ones((nd.shape[1], 1), dtype=np.float32)                        # _kmeans.py:268
            bj = int(np.argmin(pots))
            best[g] = bj
            pot[g] = pots[bj]
            closest_g[g] = newd_g[g][bj:bj + 1].copy()
        best_d = torch.from_numpy(best).to(dev)
        take = torch.zeros(n_seg, dtype=torch.bool, device=dev)
        for g in active_seg:
            take[g] = c < seg_k_host[g]
        cent[:, c] = torch.where(take[:, None], vec[torch.arange(n_seg, device=dev), best_d], cent[:, c])
        if x.shape[0]:
            sel = newd[best_d[seg_of_row], torch.arange(x.shape[0], device=dev)]
            closest = torch.where(take[seg_of_row], sel, closest).contiguous()
    return cent



def _sklearn_uniform_stream(global_sizes, seg_k_host, k, random_state):
    """The RandomState draws of `_kmeans_plusplus` do not depend on the data: first centre = rs.choice(n, p=uniform)
    (_kmeans.py:234), then `rs.uniform(size=n_local_trials)` per further centre (:252).  -> (first [n_seg],
    uniform [k, n_seg, n_trials] float64, trials [n_seg])."""
    n_seg = len(global_sizes)
    trials = [2 + int(np.log(kk)) if kk > 1 else 1 for kk in seg_k_host]
    n_trials = max(trials + [1])
    first = np.zeros(n_seg, dtype=np.int64)
    uni = np.zeros((k, n_seg, n_trials), dtype=np.float64)
    memo = {}                                          # every segment restarts RandomState(random_state): (n, k) decides
    for g in range(n_seg):
        if global_sizes[g] <= 0:
            continue
        key = (int(global_sizes[g]), int(seg_k_host[g]))
        if key not in memo:
            rs = np.random.RandomState(random_state)
            f = _sklearn_first_center(rs, key[0])
            u = np.zeros((k, n_trials), dtype=np.float64)
            for c in range(1, key[1]):
                u[c, :trials[g]] = rs.uniform(size=trials[g])
            memo[key] = (f, u)
        first[g], uni[:, g, :] = memo[key]
    return first, uni, trials


def _seed_on_device(x, global_sizes, local_off, shard, seg_k_host, k, random_state, backend, group, distributed, world,
                    rot=None):
    """k-means++ without a host round trip per centre: scan + search, gather, distance pass, pick are stream-ordered
    kernels (csrc/seed.cu); N > 1 adds an all-gather of the closest distances (the scan is global and sequential, every
    rank runs it redundantly on the same bits) and all-reduces of the candidate vectors and potentials."""
    import torch.distributed as dist
    dev = x.device
    n_seg, dim = len(global_sizes), int(x.shape[1])
    rot = list(range(n_seg)) if not rot else rot
    first, uni, trials = _sklearn_uniform_stream(global_sizes, seg_k_host, k, random_state)
    n_trials = uni.shape[2]
    i64 = lambda a: torch.tensor(np.asarray(a), dtype=torch.int64, device=dev)
    i32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.int32, device=dev)
    seg_off_d = i64(list(local_off))
    shard_first = i64([a for a, _ in shard])
    uni_d = torch.from_numpy(uni).to(dev)
    trials_d = i32(trials)
    one_d = i32([1] * n_seg)
    seg_on = i32([[1 if (global_sizes[g] > 0 and c < seg_k_host[g]) else 0 for g in range(n_seg)] for c in range(k)])
    local_sizes = [local_off[g + 1] - local_off[g] for g in range(n_seg)]
    max_rows = max(local_sizes + [0])
    max_rows_global = max([int(n) for n in global_sizes] + [0])
    n_local = int(x.shape[0])
    if distributed:
        meta = torch.tensor([n_local] + local_sizes, dtype=torch.int64, device=dev)
        metas = [torch.empty_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta, group=group)
        metas = [m.cpu().numpy() for m in metas]
        max_local = max(int(m[0]) for m in metas)
        # the pieces of a segment in ROW order (the owner of the j-th range rotates with the segment, ranks_in_row_order)
        piece_off = i64([[r * max_local + int(metas[r][1:1 + g].sum()) for r in ranks_in_row_order(rot[g], world)] for g in range(n_seg)])
        piece_cnt = i64([[int(metas[r][1 + g]) for r in ranks_in_row_order(rot[g], world)] for g in range(n_seg)])
        gathered = torch.zeros((world, max_local), dtype=torch.float32, device=dev)
    else:
        piece_off = i64([[local_off[g]] for g in range(n_seg)])
        piece_cnt = i64([[local_sizes[g]] for g in range(n_seg)])

    def allreduce(t):
        if distributed:
            dist.all_reduce(t, group=group)
        return t

    cent = torch.zeros((n_seg, k, dim), dtype=torch.float32, device=dev)
    closest = torch.zeros(n_local, dtype=torch.float32, device=dev)
    pot = torch.zeros(n_seg, dtype=torch.float32, device=dev)
    cand_id = i64(first.reshape(n_seg, 1))
    vec = allreduce(backend.seed_gather(x, cand_id, seg_off_d, shard_first))
    newd, pots = backend.seed_sqdist(x, seg_off_d, max_rows, vec, None)
    allreduce(pots)
    backend.seed_pick(pots, one_d, seg_on[0], seg_off_d, max_rows, newd, vec, closest, pot, cent, 0)
    cand_id = torch.zeros((n_seg, n_trials), dtype=torch.int64, device=dev)
    for c in range(1, k):
        if distributed:
            pad = torch.zeros(max_local, dtype=torch.float32, device=dev)
            pad[:n_local] = closest
            if dist.get_backend(group) == "nccl":
                dist.all_gather_into_tensor(gathered, pad, group=group)
            else:                                                          # gloo (CPU tests, two ranks on one GPU)
                parts = [torch.empty_like(pad) for _ in range(world)]
                dist.all_gather(parts, pad, group=group)
                gathered.copy_(torch.stack(parts))
            closest_all = gathered
        else:
            closest_all = closest
        backend.seed_scan(closest_all, piece_off, piece_cnt, uni_d[c], pot, trials_d, seg_on[c], max_rows_global, cand_id)
        vec = allreduce(backend.seed_gather(x, cand_id, seg_off_d, shard_first))
        newd, pots = backend.seed_sqdist(x, seg_off_d, max_rows, vec, closest)
        allreduce(pots)
        backend.seed_pick(pots, trials_d, seg_on[c], seg_off_d, max_rows, newd, vec, closest, pot, cent, c)
    return cent


def _seg_first_local(table: BlockTable, n_seg: int, dev) -> torch.Tensor:
    seg = table.seg.cpu().numpy()
    first = np.zeros(n_seg + 1, dtype=np.int32)
    if len(seg):
        first[1:] = np.cumsum(np.bincount(seg, minlength=n_seg))
    return torch.from_numpy(first).to(dev)


def _ordered_reduce(backend, psums, pcounts, table: BlockTable, n_seg: int, world: int, group):
    """Rank-count-invariant reduction: block partials -> super-block partials (fixed order inside a super-block),
    all-gather of the super-block partials in global row order, then a fixed-order sum per segment."""
    import torch.distributed as dist
    dev = table.seg.device
    k, dim = psums.shape[1], psums.shape[2]
    sup_s = backend.reduce(psums, table.super_first, table.n_super_local) if table.n_super_local else psums.new_zeros((0, k, dim))
    sup_c = backend.reduce(pcounts, table.super_first, table.n_super_local) if table.n_super_local else pcounts.new_zeros((0, k))
    packed = torch.cat([sup_s.reshape(sup_s.shape[0], -1), sup_c], dim=1)                 # [n_super_local, k*dim + k]
    cmax = max(table.super_owner_counts)
    pad = torch.zeros((cmax, k * dim + k), dtype=torch.float32, device=dev)
    pad[:packed.shape[0]] = packed
    if world > 1:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
    else:
        bufs = [pad]
    # global order = segment-major, and inside a segment rank-major (= row order)
    per_rank_seg = _super_per_rank_seg(table, n_seg, world)
    rows, offs = [], [0] * world
    for g in range(n_seg):
        for r in ranks_in_row_order(table.rot[g], world):
            c = per_rank_seg[r][g]
            rows.append(bufs[r][offs[r]:offs[r] + c])
            offs[r] += c
    allp = torch.cat(rows, dim=0).contiguous()
    sums = backend.reduce(allp[:, :k * dim].contiguous(), table.super_seg_first, n_seg).reshape(n_seg, k, dim)
    cnts = backend.reduce(allp[:, k * dim:].contiguous(), table.super_seg_first, n_seg).reshape(n_seg, k)
    return sums, cnts


def _super_per_rank_seg(table: BlockTable, n_seg: int, world: int):
    first = table.super_seg_first.cpu().numpy()
    out = [[0] * n_seg for _ in range(world)]
    for g in range(n_seg):
        n_super = int(first[g + 1] - first[g])
        bounds = [(n_super * r) // world for r in range(world + 1)]
        for r in range(world):
            ri = range_index(table.rot[g], world, r)
            out[r][g] = bounds[ri + 1] - bounds[ri]
    return out


def kmeans_fit_sharded(x_local: torch.Tensor, local_sizes: Sequence[int], global_sizes: Sequence[int], k: int, world: int,
                       rank: int, group=None, random_state: int = 10, rot: Optional[Sequence[int]] = None, **kw) -> KMeansResult:
    """N > 1 entry: this rank holds `local_sizes[g]` rows of segment g (segment-major in x_local).  The block table
    only depends on the GLOBAL sizes, and build_blocks prescribes which rows each rank owns; the caller must have
    sharded accordingly (see shard_rows)."""
    table, shard, local_off = build_blocks(global_sizes, world, rank, x_local.device, rot,
                                           auto_block_rows(global_sizes, world, kw.get("reduce", "allreduce")))
    mine = [cnt for _, cnt in shard]
    if list(mine) != [int(v) for v in local_sizes]:
        raise ValueError(f"rank {rank}: row shard {list(local_sizes)} does not match the block table's {mine}; "
                         f"use kmeans.shard_rows() to split the segments")
    return kmeans_fit(x_local, global_sizes, k, table, local_off, shard, random_state=random_state, group=group, **kw)


def shard_rows(global_sizes: Sequence[int], world: int, rank: int, rot: Optional[Sequence[int]] = None) -> List[tuple]:
    """(first row, row count) of every segment owned by `rank`: a contiguous range of whole super-blocks; which of the
    `world` ranges of a segment a rank owns rotates with the segment index (range_index)."""
    out = []
    for g, n in enumerate(global_sizes):
        n_super = (n + SUPER_ROWS - 1) // (SUPER_ROWS)
        ri = range_index(g if rot is None else int(rot[g]), world, rank)
        s0, s1 = (n_super * ri) // world, (n_super * (ri + 1)) // world
        a, b = min(n, s0 * SUPER_ROWS), min(n, s1 * SUPER_ROWS)
        out.append((a, b - a))
    return out


def member_means(x: torch.Tensor, sizes: Sequence[int], labels: Optional[torch.Tensor], k: int, group=None, backend=None,
                 reduce: str = "allreduce", global_sizes: Optional[Sequence[int]] = None, rot: Optional[Sequence[int]] = None):
    """Per segment, per label: mean of the member rows -- `np.mean(X[labels == j], axis=0)` of
    /root/reference/ood_utils.py:2359-2366 (labels=None: one cluster = the segment mean, :2306).
    Sums run through the k-means step kernel in its update==2 mode (block partials in row order + fixed-order
    reduction); N > 1: one all-reduce of the sums and counts, or, with reduce="ordered" (needs the global sizes: the
    local rows are the shard `shard_rows` prescribes), the rank-count-invariant reduction of `_ordered_reduce`, whose
    result is bit-identical for 1 / 2 / 4 / 8 ranks.  -> (means [n_seg, k, dim], counts [n_seg, k])."""
    import torch.distributed as dist
    dev = x.device
    backend = backend or CudaBackend(dev)
    n_seg, dim = len(sizes), int(x.shape[1])
    ordered = reduce == "ordered"
    if ordered:
        world = dist.get_world_size(group) if group is not None else 1
        rank = dist.get_rank(group) if group is not None else 0
        gsz = list(sizes) if global_sizes is None else list(global_sizes)
        table, shard, _ = build_blocks(gsz, world, rank, dev, rot)
        if [cnt for _, cnt in shard] != [int(v) for v in sizes]:
            raise ValueError("member_means(reduce='ordered'): the local rows must be the shard kmeans.shard_rows() prescribes")
    else:
        table, _, _ = build_blocks(sizes, 1, 0, dev)
    if labels is None:
        labels = torch.zeros(x.shape[0], dtype=torch.int32, device=dev)
    labels = labels.to(torch.int32).contiguous()
    seg_k = torch.full((n_seg,), k, dtype=torch.int32, device=dev)
    cent = torch.zeros((n_seg, k, dim), dtype=torch.float32, device=dev)
    dummy = torch.zeros(n_seg, dtype=torch.int32, device=dev)
    if ordered:
        if table.n_blocks:
            psums, pcounts = backend.step(x.contiguous(), k, seg_k, cent, table, None, labels, dummy, 2)
        else:
            psums, pcounts = cent.new_zeros((0, k, dim)), cent.new_zeros((0, k))
        sums, cnts = _ordered_reduce(backend, psums, pcounts, table, n_seg, world, group)
    else:
        if table.n_blocks:
            psums, pcounts = backend.step(x.contiguous(), k, seg_k, cent, table, None, labels, dummy, 2)
            first = _seg_first_local(table, n_seg, dev)
            sums = backend.reduce(psums, first, n_seg)
            cnts = backend.reduce(pcounts, first, n_seg)
        else:
            sums, cnts = cent, torch.zeros((n_seg, k), dtype=torch.float32, device=dev)
        if group is not None:
            flat = torch.cat([sums.reshape(-1), cnts.reshape(-1)])
            dist.all_reduce(flat, group=group)
            sums, cnts = flat[:sums.numel()].reshape(sums.shape), flat[sums.numel():].reshape(cnts.shape)
    means = sums / cnts.clamp_min(1.0)[..., None]
    return means, cnts


def member_medians(x: torch.Tensor, sizes: Sequence[int], labels: Optional[torch.Tensor], k: int):
    """`agg_method='median'`: per segment, per label `np.median(X[labels == j], axis=0)` (/root/reference/ood_utils.py:
    1481-1483, :2306, :2365).  One device sort per (segment, label) (torch.sort = CUB radix sort: memory plumbing, the
    order statistic needs every member row); an even count averages the two middle rows in float32 like numpy does.
    Not on the bandwidth-critical path (the paper's runs use 'mean').  -> (medians [n_seg, k, dim], counts [n_seg, k])."""
    dev = x.device
    n_seg, dim = len(sizes), int(x.shape[1])
    med = torch.zeros((n_seg, k, dim), dtype=torch.float32, device=dev)
    cnt = torch.zeros((n_seg, k), dtype=torch.float32, device=dev)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    for g in range(n_seg):
        xs = x[off[g]:off[g + 1]]
        lab = None if labels is None else labels[off[g]:off[g + 1]]
        for j in range(k if labels is not None else 1):
            rows = xs if lab is None else xs[lab == j]
            m = int(rows.shape[0])
            if m == 0:
                continue
            srt = torch.sort(rows, dim=0).values
            med[g, j] = srt[m // 2] if m % 2 else (srt[m // 2 - 1] + srt[m // 2]) / 2
            cnt[g, j] = m
    return med, cnt


def kmeans_fit_predict_single(x: torch.Tensor, sizes: Sequence[int], k: int, random_state: int = 10, **kw) -> KMeansResult:
    """Single-process convenience wrapper: x [sum(sizes), dim] holds the segments back to back."""
    table, shard, local_off = build_blocks(sizes, 1, 0, x.device, block_rows=auto_block_rows(sizes, 1, kw.get("reduce", "allreduce")))
    return kmeans_fit(x, sizes, k, table, local_off, shard, random_state=random_state, **kw)
