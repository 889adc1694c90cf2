// K1 (RoIAlign 1x1 adaptive pooling) and K1+K2 (pool -> normalise -> distance-min-threshold), sm_100a.
//
// Replaces /root/reference/ultralytics/models/yolo/detect/predict.py:13-90 (torchvision roi_align
// with output 1x1, adaptive sampling grid, aligned=False) and the per-box loop of
// /root/reference/ood_utils.py:2038-2180 (+ :2404-2409 normalize, :2422-2430 pairwise distance + min).
//
// Launches per batch: memset(counters) -> plan_geo_kernel -> items_kernel -> score_kernel.
//  plan_geo_kernel  one CTA per image.  Warp 0: quirk-Q1 class / output slot of every box (ballot prefix ranks) and the
//                position of every box's work items in the item list.  Then one warp per box: ROI geometry and the
//                separable RoIAlign weights.  With a 1x1 output bin
//                    sum_{iy,ix} bilinear(y_iy, x_ix) = sum_r sum_c wy[r] wx[c] v[r,c]
//                because both the bilinear weights and the "sample outside [-1,H]x[-1,W] contributes 0" mask are
//                products of a y-term and an x-term.  Sample coordinates use the float32 operation order of the
//                reference kernel (no FMA contraction).  Every window element is then read ONCE, not ~4 times.
//                Emits self-contained item records: item = (box, slice of kSliceChannels channels).
//  items_kernel  persistent grid, ONE WARP per work item (round-robin, atomic queue for the tail): no block barriers,
//                no tail of big boxes.  A window is addressed in 16-byte chunks (NCHW rows are 16-byte aligned for the
//                usual map widths): lane = chunk, one LDG.128 per lane per channel, 8 channels in flight, 4 FMAs per load
//                against the lane's fixed weight vector, and ONE transposing butterfly per 8 channels instead of 8 warp
//                reductions.  Small windows put 2/4/8 channels into one 32-lane request.  Its prologue groups the boxes
//                by (stride, class) for the score kernel (counting sort on the plan's histogram).
//  score_kernel  CTAs per (stride, class) group: the group's K centroid rows are staged once into shared memory by
//                bulk-async copies; one warp per box: vector in registers, L2 norm, then L1 / L2 / cosine against the K
//                rows in one sweep, first-minimum arg-min and the float64 threshold compare.
// Every pooled element is produced by exactly one warp in a fixed order: results do not depend on scheduling.
// What bounds the gather on B200 (scripts/micro/*.cu): an L2 miss always moves a whole 128-byte line
// from HBM while an NCHW window row is 8..52 bytes, so the HBM traffic of this kernel is the set of LINES the windows touch.
#include "common.cuh"

#include <cooperative_groups.h>

#include <float.h>
#include <limits.h>
#include <stdlib.h>

namespace oodb200 {

namespace cg = cooperative_groups;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxSlices = 255;
#ifndef OODB200_FMAP_MIN_BLOCKS
#define OODB200_FMAP_MIN_BLOCKS 2
#endif
#ifndef OODB200_FMAP_SLICE
#define OODB200_FMAP_SLICE 64
#endif
#ifndef OODB200_FMAP_CU
#define OODB200_FMAP_CU 16                  // channel requests in flight per warp for windows of <= 32 chunks (8 beyond)
#endif
#ifndef OODB200_FMAP_STATIC_PCT
#define OODB200_FMAP_STATIC_PCT 75          // share of the work list handed out round-robin, the rest through the queue
#endif
constexpr int kSliceChannels = OODB200_FMAP_SLICE;   // channels per work item (multiple of 32)
constexpr int kCU = OODB200_FMAP_CU;
#ifndef OODB200_NHWC_SLICE
#define OODB200_NHWC_SLICE 32
#endif
constexpr int kSliceNhwc = OODB200_NHWC_SLICE;       // channels-last: channels per work item (32 = one 128-byte line per cell)
constexpr int kNhwcLanes = kSliceNhwc / 4;           // lanes per cell (128 bits each)
constexpr int kNhwcCells = 32 / kNhwcLanes;          // cells per warp request
static_assert(kSliceNhwc == 32 || kSliceNhwc == 64 || kSliceNhwc == 128, "channels-last slice: 32, 64 or 128 channels");

struct FmapParams {
    const float* const* map_ptrs;
    int C[3], H[3], W[3];
    float scale[3];
    int ns[3];                  // slices per box, per stride
    int nhwc;                   // maps are channels-last: element (c, y, x) of a map at [(y * W + x) * C + c]
    int slice;                  // channels per work item: kSliceChannels (NCHW) / kSliceNhwc (channels-last)
    const float* boxes;
    const int32_t* img_idx;
    const int32_t* stride_idx;
    const int32_t* cls;
    const int32_t* img_start;   // [n_img+1] prefix of boxes per image
    int compat_q1;              // quirk Q1: class by in-stride index, stride-major output
    int n, n_img;
    int metric_mask;
    int normalize;
    const float* cent;
    const float* cent_unit;
    const int64_t* cent_off;
    const int32_t* cent_k;
    int nc;
    const double* thr;
    float* dist;
    int32_t* argmin;
    uint8_t* decision;
    float* pooled_user;         // optional caller buffer [n, pooled_user_ld], rows in output order
    int pooled_user_ld;
    // workspace
    int* counters;              // [0] items emitted by the plan, [1] item queue head, [2] boxes with a valid stride
    int* hist;                  // [3*nc + 1] boxes per (stride, class used) key; last = class outside [0, nc)
    int* cursor;                // [3*nc + 1] scatter cursors of the counting sort
    int32_t* sorted;            // [n] valid boxes grouped by key: the score kernel's order (centroid slices hit in L1)
    int32_t* cls_used;          // [n]
    int32_t* out_index;         // [n]
    int2* ipos;                 // [n] {index of the box's slice-0 item, item stride between its slices}
    float* wts;                 // [n][wstride]: wy[ext_y] | wx padded to chunks [ext_x]
    int ext_y, wstride;
    int4* items;                // [<= n * max ns][2] self-contained item records
    float* pooled;              // [n][pooled_ld] raw pooled vectors (workspace), rows in output order
    int pooled_ld;
    int group_score;            // OODB200_FMAP_GROUP_SCORE: plan -> gather -> score by (stride, class) groups (large tables)
};

struct AxisSample {
    int low, high;
    float l, h;
    bool valid;
};

// torchvision roi_align forward, one axis of one sample (see oracle/roi_align.py::_axis_samples).
// Intrinsics keep the reference's rounding: nvcc must not contract mul+add into fma here.
__device__ __forceinline__ AxisSample axis_sample(float start, float size, int grid, int extent, int i) {
    AxisSample s;
    float c = __fadd_rn(start, __fdiv_rn(__fmul_rn((float)i + 0.5f, size), (float)grid));
    s.valid = !(c < -1.0f || c > (float)extent);
    if (c <= 0.f) c = 0.f;
    int low = (int)c;
    int high;
    if (low >= extent - 1) {
        high = low = extent - 1;
        c = (float)low;
    } else {
        high = low + 1;
    }
    s.low = low;
    s.high = high;
    s.l = __fsub_rn(c, (float)low);
    s.h = __fsub_rn(1.0f, s.l);
    return s;
}

// Summed bilinear weight that row/column `row` receives from the `grid` samples of one axis.
// Samples are visited in increasing order; only the index range that can touch `row` is visited
// (others would add exactly 0), so the result equals the full loop bit for bit.
__device__ __forceinline__ float axis_weight(float start, float size, int grid, int extent, int row) {
    int lo = 0, hi = grid - 1;
    const float inv = (float)grid / size;
    if (row > 0) {
        float f = ((float)(row - 1) - start) * inv - 0.5f;
        if (f > (float)grid) f = (float)grid;
        int v = (int)floorf(f) - 1;                  // one sample of slack for the rounding of f
        lo = v < 0 ? 0 : v;
    }
    if (row < extent - 1) {
        float f = ((float)(row + 1) - start) * inv - 0.5f;
        if (f < -4.f) f = -4.f;
        if (f > (float)grid) f = (float)grid;
        int v = (int)ceilf(f) + 1;
        hi = v > grid - 1 ? grid - 1 : v;
    }
    float acc = 0.f;
    for (int i = lo; i <= hi; ++i) {
        AxisSample s = axis_sample(start, size, grid, extent, i);
        if (s.valid) {
            if (s.low == row) acc = __fadd_rn(acc, s.h);
            if (s.high == row) acc = __fadd_rn(acc, s.l);
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------- plan + geometry
// plan_geo_kernel, one thread-block CLUSTER per image (1-8 CTAs, chosen from the mean number of boxes per image).
//  warp 0 of CTA 0: quirk Q1 (ood_utils.py:2152-2154: the class of the box with the same IN-STRIDE index, stride-major
//  output) and the position of every box in the work list.  Ranks come from ballot prefixes over chunks of 32 boxes;
//  category 3 = "stride outside {0,1,2}" (never pooled by the reference either: answered here).  Results go to global
//  memory (the later kernels read them) and, for the first kPgCap boxes of the image, to CTA 0's shared memory, which the
//  other CTAs of the cluster read through distributed shared memory after the cluster barrier.
//  all warps of the cluster, one box per warp at a time: ROI geometry (predict.py:64-70 -> roi_align, aligned=False), the
//  separable weights and the box's self-contained item records.  The first box's inputs are requested before the plan
//  finishes.
#ifndef OODB200_PG_THREADS
#define OODB200_PG_THREADS 1024
#endif
constexpr int kPgThreads = OODB200_PG_THREADS, kPgWarps = kPgThreads / 32, kPgCap = 1024;

struct BoxGeo {
    int ylo, xa, wh, nxc;
    float count;
    int xoff, ww;                // first live column relative to xa, number of live columns (channels-last path)
};

// ROI geometry + separable weights of one box (one warp); writes the weights (wy[ext_y] | wx padded to 16-byte chunks),
// returns what the item records need.
__device__ __forceinline__ BoxGeo geo_compute_into(const FmapParams& p, int s, float4 bx, float* __restrict__ wy, float* __restrict__ wx) {
    const int lane = threadIdx.x & 31;
    const int H = p.H[s], W = p.W[s];
    const float sc = p.scale[s];
    const float sw = __fmul_rn(bx.x, sc), sh = __fmul_rn(bx.y, sc);
    const float ew = __fmul_rn(bx.z, sc), eh = __fmul_rn(bx.w, sc);
    const float rw = fmaxf(__fsub_rn(ew, sw), 1.0f), rh = fmaxf(__fsub_rn(eh, sh), 1.0f);
    const int gw = (int)ceilf(rw), gh = (int)ceilf(rh);
    int ylo = INT_MAX, yhi = -1, xlo = INT_MAX, xhi = -1;
    for (int k = lane; k < gh; k += 32) {
        const AxisSample a = axis_sample(sh, rh, gh, H, k);
        if (a.valid) { ylo = min(ylo, a.low); yhi = max(yhi, a.high); }
    }
    for (int k = lane; k < gw; k += 32) {
        const AxisSample a = axis_sample(sw, rw, gw, W, k);
        if (a.valid) { xlo = min(xlo, a.low); xhi = max(xhi, a.high); }
    }
    ylo = __reduce_min_sync(kFull, ylo); yhi = __reduce_max_sync(kFull, yhi);
    xlo = __reduce_min_sync(kFull, xlo); xhi = __reduce_max_sync(kFull, xhi);
    BoxGeo g = {0, 0, 0, 0, (float)max(gh * gw, 1), 0, 0};
    if (yhi >= 0 && xhi >= 0) {
        g.ylo = ylo;
        g.wh = yhi - ylo + 1;
        const int ww = xhi - xlo + 1;
        g.xa = xlo & ~3;
        g.xoff = xlo - g.xa;
        g.ww = ww;
        g.nxc = ((xlo + ww + 3) >> 2) - (g.xa >> 2);
        for (int t = lane; t < g.wh + 4 * g.nxc; t += 32) {   // rows and columns share one pass over the lanes
            if (t < g.wh) {
                wy[t] = axis_weight(sh, rh, gh, H, ylo + t);
            } else {
                const int i = t - g.wh, x = g.xa + i;
                wx[i] = (x >= xlo && x <= xhi) ? axis_weight(sw, rw, gw, W, x) : 0.f;
            }
        }
    }
    return g;
}

__device__ __forceinline__ BoxGeo geo_compute(const FmapParams& p, int box, int s, float4 bx) {
    float* wy = p.wts + (size_t)box * p.wstride;
    return geo_compute_into(p, s, bx, wy, wy + p.ext_y);
}

// record: everything an item warp needs except the weights; item(sl) = pos + sl * nbs
__device__ __forceinline__ void geo_emit(const FmapParams& p, int box, int s, unsigned long long mapp, const BoxGeo& g,
                                         int pos, int nbs, int out) {
    const int lane = threadIdx.x & 31;
    for (int sl = lane; sl < p.ns[s]; sl += 32) {
        int4* rec = p.items + 2 * (size_t)(pos + sl * nbs);
        rec[0] = make_int4((int)(((uint32_t)sl << 24) | (uint32_t)box), g.ylo | (g.xa << 16), g.wh | (g.nxc << 16), out);
        rec[1] = make_int4((int)(mapp & 0xffffffffu), (int)(mapp >> 32), __float_as_int(g.count), s | (g.xoff << 2) | (g.ww << 4));
    }
}

__global__ void __launch_bounds__(kPgThreads) plan_geo_kernel(const FmapParams p) {
    __shared__ int s_out[kPgCap], s_pos[kPgCap], s_cls[kPgCap];
    __shared__ signed char s_cat[kPgCap + 32];
    __shared__ int s_meta[8];                          // boxes per category [0..3], base of the image's item range [4]
    cg::cluster_group cluster = cg::this_cluster();
    const int G = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int img = blockIdx.x / G, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = p.img_start[img], m = p.img_start[img + 1] - b0;
    const bool planner = rank == 0 && warp == 0;
    const int stride_b = G * kPgWarps;
    int b = rank * kPgWarps + warp;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = -1;
    unsigned long long mapp = 0;
    auto fetch = [&](int bb, float4& fbx, int& fs, unsigned long long& fmap) {
        fs = p.stride_idx[b0 + bb];
        fbx = *reinterpret_cast<const float4*>(p.boxes + 4 * (size_t)(b0 + bb));
        const int im = p.img_idx[b0 + bb];
        fmap = (fs >= 0 && fs <= 2) ? (unsigned long long)p.map_ptrs[im * 3 + fs] : 0ull;
    };
    if (!planner && b < m) fetch(b, bx, s, mapp);
    if (planner) {
        const unsigned lt = (1u << lane) - 1u;
        int cnt[4] = {0, 0, 0, 0};
        for (int c0 = 0; c0 < m; c0 += 128) {          // 4 chunks of 32 boxes per trip; the raw classes ride along
            int st[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int bb = c0 + 32 * i + lane;
                st[i] = bb < m ? p.stride_idx[b0 + bb] : -2;
                if (bb < m && bb < kPgCap && p.cls) s_cls[bb] = p.cls[b0 + bb];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int bb = c0 + 32 * i + lane;
                const int cat = bb < m ? ((st[i] >= 0 && st[i] <= 2) ? st[i] : 3) : 4;
                if (bb < kPgCap) s_cat[bb] = (signed char)cat;
#pragma unroll
                for (int t = 0; t < 4; ++t) cnt[t] += __popc(__ballot_sync(kFull, cat == t));
            }
        }
        // this image's range of the work list, heaviest stride (most channels) first; items of one stride are slice-major
        const int tot = cnt[2] * p.ns[2] + cnt[1] * p.ns[1] + cnt[0] * p.ns[0];
        int base = 0;
        if (lane == 0 && tot) base = atomicAdd(&p.counters[0], tot);       // awaited only where `base` is used
        if (lane == 0 && p.cent && cnt[0] + cnt[1] + cnt[2]) atomicAdd(&p.counters[2], cnt[0] + cnt[1] + cnt[2]);
        if (lane < 4) s_meta[lane] = cnt[lane];
        __syncwarp();
        int item0[3];                                  // relative to base
        item0[2] = 0;
        item0[1] = cnt[2] * p.ns[2];
        item0[0] = item0[1] + cnt[1] * p.ns[1];
        int run[4] = {0, 0, 0, 0};
        const int base_all = m > kPgCap ? __shfl_sync(kFull, base, 0) : 0;   // awaited only for images beyond the shared-memory cap
        for (int c0 = 0; c0 < m; c0 += 32) {
            const int bb = c0 + lane;
            int cat;
            if (bb < kPgCap) {
                cat = s_cat[bb];
            } else {
                const int st = bb < m ? p.stride_idx[b0 + bb] : -2;
                cat = bb < m ? ((st >= 0 && st <= 2) ? st : 3) : 4;
            }
            int j = 0, before = 0, nb = 0, it0 = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const unsigned mk = __ballot_sync(kFull, cat == t);
                if (cat == t) { j = run[t] + __popc(mk & lt); nb = cnt[t]; if (t < 3) it0 = item0[t]; }
                if (t < cat) before += cnt[t];
                run[t] += __popc(mk);
            }
            if (bb >= m) continue;
            const bool ok = cat < 3;
            int cls_u = p.cls ? (bb < kPgCap ? s_cls[bb] : p.cls[b0 + bb]) : 0, out = b0 + bb;
            if (p.compat_q1) {
                cls_u = ok ? (j < kPgCap ? s_cls[j] : p.cls[b0 + j]) : -1;
                out = b0 + before + j;
            }
            p.cls_used[b0 + bb] = cls_u;
            p.out_index[b0 + bb] = out;
            if (bb < kPgCap) { s_out[bb] = out; s_pos[bb] = it0 + j; }
            if (ok) {
                if (bb >= kPgCap) p.ipos[b0 + bb] = make_int2(base_all + it0 + j, nb);
                if (p.cent) atomicAdd(&p.hist[(cls_u >= 0 && cls_u < p.nc) ? cat * p.nc + cls_u : 3 * p.nc], 1);
            } else if (p.cent) {
                for (int k = 0; k < OODB200_N_METRICS; ++k)
                    if (p.metric_mask >> k & 1) {
                        const size_t o = (size_t)k * p.n + out;
                        p.dist[o] = nanf("");
                        p.argmin[o] = -1;
                        p.decision[o] = 0;
                    }
            }
        }
        if (lane == 0) s_meta[4] = base;
    }
    // The geometry of a warp's first box does not depend on the plan: it overlaps the planner's chain of round trips.
    // The planner arrives at the cluster barrier first and computes its own box between arrive and wait.
    BoxGeo g = {0, 0, 0, 0, 1.f, 0, 0};
    if (!planner && b < m && s >= 0 && s <= 2) g = geo_compute(p, b0 + b, s, bx);
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    if (planner && b < m) {
        fetch(b, bx, s, mapp);
        if (s >= 0 && s <= 2) g = geo_compute(p, b0 + b, s, bx);
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // the plan's global writes are visible too
    const int* __restrict__ r_out = cluster.map_shared_rank(s_out, 0);
    const int* __restrict__ r_pos = cluster.map_shared_rank(s_pos, 0);
    const int* __restrict__ r_meta = cluster.map_shared_rank(s_meta, 0);
    const int base = r_meta[4];
    bool have_geo = true;
    while (b < m) {
        const int nb_ = b + stride_b;
        float4 nbx = make_float4(0.f, 0.f, 0.f, 0.f);
        int nst = -1;
        unsigned long long nmap = 0;
        if (nb_ < m) fetch(nb_, nbx, nst, nmap);
        if (s >= 0 && s <= 2) {
            const int box = b0 + b;
            if (!have_geo) g = geo_compute(p, box, s, bx);
            int pos, out;
            if (b < kPgCap) { pos = base + r_pos[b]; out = r_out[b]; }
            else { pos = p.ipos[box].x; out = p.out_index[box]; }
            geo_emit(p, box, s, mapp, g, pos, r_meta[s], out);
        }
        have_geo = false;
        b = nb_; bx = nbx; s = nst; mapp = nmap;
    }
    cluster.sync();                                    // CTA 0's shared memory stays alive until every reader is done
}

// standalone Q1 plan (same arithmetic as the plan phase of plan_geo_kernel, without the work list)
__global__ void q1_plan_kernel(const int32_t* __restrict__ img_start, const int32_t* __restrict__ stride_idx,
                               const int32_t* __restrict__ cls, int32_t* __restrict__ cls_used,
                               int32_t* __restrict__ out_index) {
    const int img = blockIdx.x;
    const int b0 = img_start[img], m = img_start[img + 1] - b0;
    __shared__ int s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < m; b += blockDim.x) {
        const int s = stride_idx[b0 + b];
        atomicAdd(&s_cnt[(s >= 0 && s <= 2) ? s : 3], 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < m; b += blockDim.x) {
        const int s = stride_idx[b0 + b];
        const bool ok = s >= 0 && s <= 2;
        int j = 0;
        for (int e = 0; e < b; ++e) {
            const int se = stride_idx[b0 + e];
            j += ok ? (se == s) : !(se >= 0 && se <= 2);
        }
        int before = 0;
        for (int t = 0; t < (ok ? s : 3); ++t) before += s_cnt[t];
        cls_used[b0 + b] = ok ? cls[b0 + j] : -1;
        out_index[b0 + b] = b0 + before + j;
    }
}

// ---------------------------------------------------------------------------------------------- gather + pool
template <int N> struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// Segmented transposing butterfly: the 32 lanes are 32/G segments of G = 1<<LG lanes; v[u] holds this lane's partial
// sum of request u.  Afterwards every lane holds the segment total of request u = (lane & (G-1)) >> (LG - log2 CU).
template <int LG, int CU>
__device__ __forceinline__ float seg_butterfly(float (&v)[CU], int lane) {
    static_assert(Log2<CU>::v <= LG, "CU requests need log2(CU) butterfly levels inside a segment");
    int o = 1 << (LG - 1);
#pragma unroll
    for (int n = CU; n > 1; n >>= 1, o >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, o);
        }
    }
    float r = v[0];
#pragma unroll
    for (; o > 0; o >>= 1) r += __shfl_xor_sync(kFull, r, o);
    return r;
}

// Pool channels [c_lo, c_hi) of one window.  G = 1<<LG lanes cover the window's chunks (T chunk slots per lane),
// 32/G channels share one request, CU requests are in flight.
template <int LG, int T, int CU, bool ACC = false>
__device__ __forceinline__ void pool_slice(const float* __restrict__ img, int C, int HW, int W, int4 geo,
                                           const float* __restrict__ wy, const float* __restrict__ wx, float count,
                                           int c_lo, int c_hi, float* __restrict__ out0, float* __restrict__ out1) {
    constexpr int G = 1 << LG, CPR = 32 >> LG, STEP = CU * CPR;
    const int lane = threadIdx.x & 31, l = lane & (G - 1), sub = lane >> LG;
    const int y0 = geo.x, xa = geo.y, nxc = geo.w, nch = geo.z * geo.w;
    float4 w[T];
    int off[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int q = l + G * t;
        if (q < nch) {
            const int r = q / nxc, xc = q - r * nxc;
            const float a = wy[r];
            const float4 b = reinterpret_cast<const float4*>(wx)[xc];
            w[t] = make_float4(a * b.x, a * b.y, a * b.z, a * b.w);
            off[t] = (y0 + r) * W + xa + 4 * xc;
        } else {
            w[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            off[t] = y0 * W + xa;
        }
    }
    const int u_mine = l >> (LG - Log2<CU>::v);
    const bool holder = (l & ((1 << (LG - Log2<CU>::v)) - 1)) == 0;
    for (int cb = c_lo; cb < c_hi; cb += STEP) {
        float4 v[CU][T];
#pragma unroll
        for (int u = 0; u < CU; ++u) {
            const int c = min(cb + u * CPR + sub, C - 1);            // partial last step: clamp, result dropped
            const float* __restrict__ pc = img + (size_t)c * HW;
#pragma unroll
            for (int t = 0; t < T; ++t) v[u][t] = __ldg(reinterpret_cast<const float4*>(pc + off[t]));
        }
        float a[CU];
#pragma unroll
        for (int u = 0; u < CU; ++u) {
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                s = fmaf(w[t].x, v[u][t].x, s); s = fmaf(w[t].y, v[u][t].y, s);
                s = fmaf(w[t].z, v[u][t].z, s); s = fmaf(w[t].w, v[u][t].w, s);
            }
            a[u] = s;
        }
        const float tot = seg_butterfly<LG, CU>(a, lane);
        const int c = cb + u_mine * CPR + sub;
        if (holder && c < c_hi) {
            float val = __fdiv_rn(tot, count);                        // average over the sample grid (roi_align.py:192-196)
            if (ACC) val += out0[c];                                  // later pass of a tiled window (same lane, same channel)
            out0[c] = val;
            if (out1) out1[c] = val;
        }
    }
}

// Maps whose rows are not 16-byte aligned (W % 4 != 0, odd base pointer): plain scalar gather, lanes over the window.
__device__ __noinline__ void pool_scalar(const float* __restrict__ img, int HW, int W, int y0, int xa, int wh, int nxc,
                                         const float* __restrict__ wy, const float* __restrict__ wx, float count,
                                         int c_lo, int c_hi, float* __restrict__ out0, float* __restrict__ out1) {
    const int lane = threadIdx.x & 31;
    const int ww = 4 * nxc, P = wh * ww;
    for (int c = c_lo; c < c_hi; ++c) {
        const float* __restrict__ pc = img + (size_t)c * HW;
        float s = 0.f;
        for (int q = lane; q < P; q += 32) {
            const int r = q / ww, x = q - r * ww;
            const float wgt = wy[r] * wx[x];
            if (wgt != 0.f) s = fmaf(wgt, __ldg(pc + (y0 + r) * W + xa + x), s);   // padding columns may lie outside the row
        }
        s = warp_sum(s);
        if (lane == 0) {
            const float val = __fdiv_rn(s, count);
            out0[c] = val;
            if (out1) out1[c] = val;
        }
    }
}

__device__ __forceinline__ void pool_dispatch(const float* img, int C, int HW, int W, int4 geo, const float* wy,
                                              const float* wx, float count, int c_lo, int c_hi, float* out0, float* out1) {
    const int nch = geo.z * geo.w;
    constexpr int CU1 = kCU, CU2 = kCU >= 8 ? 8 : kCU, CU4 = kCU >= 8 ? 4 : kCU / 2;
    if (nch <= 4) pool_slice<2, 1, 4>(img, C, HW, W, geo, wy, wx, count, c_lo, c_hi, out0, out1);
    else if (nch <= 8) pool_slice<3, 1, 8>(img, C, HW, W, geo, wy, wx, count, c_lo, c_hi, out0, out1);
    else if (nch <= 16) pool_slice<4, 1, CU1>(img, C, HW, W, geo, wy, wx, count, c_lo, c_hi, out0, out1);
    else if (nch <= 32) pool_slice<5, 1, CU1>(img, C, HW, W, geo, wy, wx, count, c_lo, c_hi, out0, out1);
    else if (nch <= 64) pool_slice<5, 2, CU2>(img, C, HW, W, geo, wy, wx, count, c_lo, c_hi, out0, out1);
    else if (nch <= 128) pool_slice<5, 4, CU4>(img, C, HW, W, geo, wy, wx, count, c_lo, c_hi, out0, out1);
    else if (nch <= 256) pool_slice<5, 8, 1>(img, C, HW, W, geo, wy, wx, count, c_lo, c_hi, out0, out1);
    else {
        // very large windows (> 1024 elements per channel): tiles of <= 256 chunks, accumulated in the output row
        const int cols = min(geo.w, 256);
        const int rows = max(256 / cols, 1);
        bool first = true;
        for (int r0 = 0; r0 < geo.z; r0 += rows)
            for (int x0 = 0; x0 < geo.w; x0 += cols) {
                const int4 g2 = make_int4(geo.x + r0, geo.y + 4 * x0, min(rows, geo.z - r0), min(cols, geo.w - x0));
                if (first) pool_slice<5, 8, 1, false>(img, C, HW, W, g2, wy + r0, wx + 4 * x0, count, c_lo, c_hi, out0, out1);
                else pool_slice<5, 8, 1, true>(img, C, HW, W, g2, wy + r0, wx + 4 * x0, count, c_lo, c_hi, out0, out1);
                first = false;
            }
    }
}

// ---------------------------------------------------------------------------------------------- distance phase
struct Best {
    float d[OODB200_N_METRICS];
    int a[OODB200_N_METRICS];
};

__device__ __forceinline__ void write_result(const FmapParams& p, int s, int cls, bool cls_ok, int K, int out, const Best& b) {
    const int lane = threadIdx.x & 31;
    if (lane < OODB200_N_METRICS && (p.metric_mask >> lane & 1)) {
        const int m = lane;
        const float bd = m == 0 ? b.d[0] : (m == 1 ? b.d[1] : b.d[2]);
        const int ba = m == 0 ? b.a[0] : (m == 1 ? b.a[1] : b.a[2]);
        const float d = K > 0 ? bd : 1000.f;                         // no cluster: ood_utils.py:2159-2164
        const double t = cls_ok ? p.thr[(size_t)m * 3 * p.nc + s * p.nc + cls] : nan("");
        const size_t o = (size_t)m * p.n + out;
        p.dist[o] = d;
        p.argmin[o] = K > 0 ? ba : -1;
        p.decision[o] = (t == t && (double)d < t) ? 1 : 0;           // NaN threshold = "no threshold" -> OoD (:2173-2180)
    }
}

// Vector in registers: NJ float4 per lane (C <= 128 * NJ, C % 4 == 0, 16-byte aligned centroid slices).
// Vector path (C % 4 == 0, 16-byte aligned centroid slice).  The normalised vector sits in a warp-private shared-memory
// row; the warp is 4 groups of 8 lanes and every group sweeps ONE centroid row with independent 128-bit loads, so four
// rows (x two tables) are in flight per warp and a row total needs a 3-level reduction instead of 5.  L2 latency, not
// bandwidth, bounds this phase: the point is requests in flight.
__device__ __forceinline__ void finalize_rows(const FmapParams& p, int s, int C, int cls, int out, float* xs) {
    const int lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    const float* __restrict__ row = p.pooled + (size_t)out * p.pooled_ld;
    const bool cls_ok = cls >= 0 && cls < p.nc;
    const int K = cls_ok ? p.cent_k[s * p.nc + cls] : 0;
    const int64_t off = K > 0 ? p.cent_off[s * p.nc + cls] : 0;
    float ss = 0.f;
    for (int d = lane * 4; d < C; d += 128) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(row + d));
        *reinterpret_cast<float4*>(xs + d) = v;
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    const bool want_l1 = p.metric_mask & (1 << OODB200_METRIC_L1);
    const bool want_l2 = p.metric_mask & (1 << OODB200_METRIC_L2);
    const bool want_cos = p.metric_mask & (1 << OODB200_METRIC_COS);
    float n2v = 1.f;
    if (p.normalize || want_cos) {
        float nrm = 1.f;
        if (p.normalize) {                             // ood_utils.py:2409 -> sklearn normalize
            nrm = sqrtf(warp_sum(ss));
            if (nrm < 10.f * FLT_EPSILON) nrm = 1.f;   // _handle_zeros_in_scale
        }
        float s2 = 0.f;
        for (int d = lane * 4; d < C; d += 128) {      // each lane re-reads what it wrote
            float4 v = *reinterpret_cast<float4*>(xs + d);
            if (p.normalize) {
                v.x = __fdiv_rn(v.x, nrm); v.y = __fdiv_rn(v.y, nrm); v.z = __fdiv_rn(v.z, nrm); v.w = __fdiv_rn(v.w, nrm);
                *reinterpret_cast<float4*>(xs + d) = v;
            }
            s2 = fmaf(v.x, v.x, s2); s2 = fmaf(v.y, v.y, s2); s2 = fmaf(v.z, v.z, s2); s2 = fmaf(v.w, v.w, s2);
        }
        if (want_cos) {                                // cosine_distances re-normalises X (pairwise.py:1171-1182)
            n2v = sqrtf(warp_sum(s2));
            if (n2v < 10.f * FLT_EPSILON) n2v = 1.f;
        }
    }
    __syncwarp();
    Best b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {INT_MAX, INT_MAX, INT_MAX}};
    for (int k0 = 0; k0 < K; k0 += 4) {
        const int k = k0 + g;
        const bool have = k < K;
        const float* __restrict__ ck = p.cent + off + (int64_t)(have ? k : K - 1) * C;
        const float* __restrict__ cu = p.cent_unit + off + (int64_t)(have ? k : K - 1) * C;
        float a1 = 0.f, a2 = 0.f, ac = 0.f;
        for (int d0 = j * 4; d0 < C; d0 += 128) {      // batches of 4 x 128-bit loads per table, issued before any use
            float4 c[4], u[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int d = min(d0 + 32 * i, C - 4);
                if (want_l1 || want_l2) c[i] = __ldg(reinterpret_cast<const float4*>(ck + d));
                if (want_cos) u[i] = __ldg(reinterpret_cast<const float4*>(cu + d));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int d = d0 + 32 * i;
                if (d >= C) break;
                const float4 x = *reinterpret_cast<const float4*>(xs + d);
                if (want_l1 || want_l2) {
                    const float e0 = x.x - c[i].x, e1 = x.y - c[i].y, e2 = x.z - c[i].z, e3 = x.w - c[i].w;
                    a1 += (fabsf(e0) + fabsf(e1)) + (fabsf(e2) + fabsf(e3));
                    a2 = fmaf(e0, e0, a2); a2 = fmaf(e1, e1, a2); a2 = fmaf(e2, e2, a2); a2 = fmaf(e3, e3, a2);
                }
                if (want_cos) {
                    ac = fmaf(x.x, u[i].x, ac); ac = fmaf(x.y, u[i].y, ac); ac = fmaf(x.z, u[i].z, ac); ac = fmaf(x.w, u[i].w, ac);
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {              // totals of the group's row
            if (want_l1) a1 += __shfl_xor_sync(kFull, a1, o);
            if (want_l2) a2 += __shfl_xor_sync(kFull, a2, o);
            if (want_cos) ac += __shfl_xor_sync(kFull, ac, o);
        }
        if (have) {
            if (want_l1 && a1 < b.d[0]) { b.d[0] = a1; b.a[0] = k; }
            if (want_l2) { const float v = sqrtf(fmaxf(a2, 0.f)); if (v < b.d[1]) { b.d[1] = v; b.a[1] = k; } }
            if (want_cos) {                            // X / ||X|| applied to the sum (same value to float32 rounding)
                const float v = fminf(fmaxf(1.0f - __fdiv_rn(ac, n2v), 0.f), 2.f);
                if (v < b.d[2]) { b.d[2] = v; b.a[2] = k; }
            }
        }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1)                  // first minimum over the 4 groups: smaller distance, then smaller index
#pragma unroll
        for (int m = 0; m < OODB200_N_METRICS; ++m) {
            const float od = __shfl_xor_sync(kFull, b.d[m], o);
            const int oa = __shfl_xor_sync(kFull, b.a[m], o);
            if (od < b.d[m] || (od == b.d[m] && oa < b.a[m])) { b.d[m] = od; b.a[m] = oa; }
        }
    write_result(p, s, cls, cls_ok, K, out, b);       // lane m reports metric m (all lanes hold the same result)
}

// Any C / alignment: the vector is re-read from the (L1-resident) pooled row for every centroid.
__device__ __noinline__ Best finalize_generic(const float* __restrict__ row, int C, int normalize, int metric_mask,
                                              const float* __restrict__ cent, const float* __restrict__ cent_unit, int K) {
    const int lane = threadIdx.x & 31;
    float nrm = 1.f, n2v = 1.f;
    if (normalize) {
        float ss = 0.f;
        for (int d = lane; d < C; d += 32) { const float v = __ldcg(row + d); ss = fmaf(v, v, ss); }
        nrm = sqrtf(warp_sum(ss));
        if (nrm < 10.f * FLT_EPSILON) nrm = 1.f;
    }
    const bool want_l1 = metric_mask & (1 << OODB200_METRIC_L1);
    const bool want_l2 = metric_mask & (1 << OODB200_METRIC_L2);
    const bool want_cos = metric_mask & (1 << OODB200_METRIC_COS);
    auto xn = [&](int d) { const float v = __ldcg(row + d); return normalize ? __fdiv_rn(v, nrm) : v; };
    if (want_cos) {
        float s2 = 0.f;
        for (int d = lane; d < C; d += 32) { const float v = xn(d); s2 = fmaf(v, v, s2); }
        n2v = sqrtf(warp_sum(s2));
        if (n2v < 10.f * FLT_EPSILON) n2v = 1.f;
    }
    Best b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-1, -1, -1}};
    for (int k = 0; k < K; ++k) {
        const float* __restrict__ ck = cent + (int64_t)k * C;
        const float* __restrict__ cu = cent_unit + (int64_t)k * C;
        float a1 = 0.f, a2 = 0.f, ac = 0.f;
        for (int d = lane; d < C; d += 32) {
            const float x = xn(d);
            if (want_l1 || want_l2) {
                const float df = x - __ldg(ck + d);
                a1 += fabsf(df);
                a2 = fmaf(df, df, a2);
            }
            if (want_cos) ac = fmaf(x, __ldg(cu + d), ac);
        }
        if (want_l1) { a1 = warp_sum(a1); if (a1 < b.d[0]) { b.d[0] = a1; b.a[0] = k; } }
        if (want_l2) { a2 = sqrtf(fmaxf(warp_sum(a2), 0.f)); if (a2 < b.d[1]) { b.d[1] = a2; b.a[1] = k; } }
        if (want_cos) {
            ac = fminf(fmaxf(1.0f - __fdiv_rn(warp_sum(ac), n2v), 0.f), 2.f);
            if (ac < b.d[2]) { b.d[2] = ac; b.a[2] = k; }
        }
    }
    return b;
}

__device__ __forceinline__ void finalize(const FmapParams& p, int s, int cls, int out, float* xs) {
    const int C = p.C[s];
    bool vec = (C % 4 == 0) && (((uintptr_t)p.cent | (uintptr_t)p.cent_unit) & 15) == 0;
    if (vec && cls >= 0 && cls < p.nc) vec = (p.cent_off[s * p.nc + cls] % 4 == 0);
    if (vec) { finalize_rows(p, s, C, cls, out, xs); return; }
    const bool cls_ok = cls >= 0 && cls < p.nc;
    const int K = cls_ok ? p.cent_k[s * p.nc + cls] : 0;
    const int64_t off = K > 0 ? p.cent_off[s * p.nc + cls] : 0;
    const Best b = finalize_generic(p.pooled + (size_t)out * p.pooled_ld, C, p.normalize, p.metric_mask, p.cent + off,
                                    p.cent_unit + off, K);
    write_result(p, s, cls, cls_ok, K, out, b);
}

// ---------------------------------------------------------------------------------------------- channels-last pooling
// Channels-last maps (what a detector run in torch.channels_last hands out): the C values of a cell are contiguous.  A
// work item is a slice of 32 channels = ONE 128-byte line per cell: 8 lanes x 128 bits cover it, so a warp request fetches
// 4 cells, every fetched line is used in full, and a pooled channel stays in its lane (the 4 cell groups are added with
// two shuffles at the very end).  Lanes first build the (offset, weight) pairs of 32 cells of the window in parallel; the
// cells are then visited with shuffles, 8 requests (32 cells) in flight.  Accumulation order is fixed (cell order).
struct NhwcItem {
    const float* img;
    int box, c_lo, s, out, y0, x0, wh, ww, xoff;
    float count;
};

__device__ __forceinline__ NhwcItem nhwc_decode(int4 r0, int4 r1) {
    NhwcItem it;
    it.img = reinterpret_cast<const float*>(((unsigned long long)(uint32_t)r1.y << 32) | (uint32_t)r1.x);
    it.box = r0.x & 0xFFFFFF;
    it.c_lo = (int)((uint32_t)r0.x >> 24) * kSliceNhwc;
    it.s = r1.w & 3;
    it.xoff = (r1.w >> 2) & 3;
    it.ww = r1.w >> 4;
    it.out = r0.w;
    it.y0 = r0.y & 0xFFFF;
    it.x0 = (int)((uint32_t)r0.y >> 16) + it.xoff;     // first live column
    it.wh = r0.z & 0xFFFF;
    it.count = __int_as_float(r1.z);
    return it;
}

// (element offset, weight) of cell q0 + lane of the item's window (0 weight beyond the window)
__device__ __forceinline__ void nhwc_cells(const FmapParams& p, const NhwcItem& it, const float* __restrict__ wy, int q0, float& wq, int& oq) {
    const int q = q0 + (threadIdx.x & 31);
    wq = 0.f;
    oq = 0;
    if (q < it.wh * it.ww) {
        const int r = q / it.ww, x = q - r * it.ww;
        wq = wy[r] * wy[p.ext_y + it.xoff + x];
        oq = ((it.y0 + r) * p.W[it.s] + it.x0 + x) * p.C[it.s];
    }
}

// 8 requests of kNhwcCells cells starting at cell `first` of the 32 built by the lanes (lane group g takes cell
// first + kNhwcCells * u + g); rounds beyond the window cost nothing
__device__ __forceinline__ void nhwc_issue(const float* __restrict__ src, bool mine, int first, int left, float wq, int oq,
                                           float4 (&v)[8], float (&w8)[8]) {
    const int g = (threadIdx.x & 31) / kNhwcLanes;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        w8[u] = 0.f;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (first + kNhwcCells * u < left) {           // warp-uniform
            const int j = (first + kNhwcCells * u + g) & 31;
            const int off = __shfl_sync(kFull, oq, j);
            w8[u] = __shfl_sync(kFull, wq, j);
            if (first + kNhwcCells * u + g < left && w8[u] != 0.f && mine) v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)off));
        }
    }
}
__device__ __forceinline__ void nhwc_consume(int first, int left, const float4 (&v)[8], const float (&w8)[8], float4& acc) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (first + kNhwcCells * u < left) {
            acc.x = fmaf(w8[u], v[u].x, acc.x); acc.y = fmaf(w8[u], v[u].y, acc.y);
            acc.z = fmaf(w8[u], v[u].z, acc.z); acc.w = fmaf(w8[u], v[u].w, acc.w);
        }
}

// any C / alignment: lanes over channels, scalar loads
__device__ __noinline__ void pool_nhwc_scalar(const FmapParams& p, const NhwcItem& it, const float* __restrict__ wy,
                                              float* __restrict__ out0, float* __restrict__ out1) {
    const int lane = threadIdx.x & 31;
    const int C = p.C[it.s], W = p.W[it.s];
    const float* __restrict__ wx = wy + p.ext_y + it.xoff;
    for (int cc = it.c_lo + lane; cc < min(C, it.c_lo + kSliceNhwc); cc += 32) {
        float acc = 0.f;
        for (int r = 0; r < it.wh; ++r)
            for (int x = 0; x < it.ww; ++x) {
                const float w = wy[r] * wx[x];
                if (w != 0.f) acc = fmaf(w, __ldg(it.img + ((size_t)(it.y0 + r) * W + it.x0 + x) * C + cc), acc);
            }
        const float val = __fdiv_rn(acc, it.count);
        out0[cc] = val;
        if (out1) out1[cc] = val;
    }
}

// One channels-last work item: the kSliceNhwc channels from a.c_lo of the window, weights at wy (wy[ext_y] | wx)
__device__ __forceinline__ void nhwc_pool_item(const FmapParams& p, const NhwcItem& a, const float* __restrict__ wy,
                                               float* __restrict__ out0, float* __restrict__ out1) {
    const int lane = threadIdx.x & 31, g = lane / kNhwcLanes, l = lane % kNhwcLanes;
    const int C = p.C[a.s];
    const int ncell = a.wh * a.ww;                 // 0: no sample inside the map (Q5) -> all-zero vector
    if ((C % 4 == 0) && (((uintptr_t)a.img & 15) == 0)) {
        const int c = a.c_lo + l * 4;
        const bool mine = c < C;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q0 = 0; q0 < ncell; q0 += 32) {
            float wq;
            int oq;
            nhwc_cells(p, a, wy, q0, wq, oq);
            const int left = min(32, ncell - q0);
            for (int first = 0; first < left; first += 8 * kNhwcCells) {
                float4 v[8];
                float w8[8];
                nhwc_issue(a.img + c, mine, first, left, wq, oq, v, w8);
                nhwc_consume(first, left, v, w8, acc);
            }
        }
#pragma unroll
        for (int o = kNhwcLanes; o <= 16; o <<= 1) {   // the cell groups
            acc.x += __shfl_xor_sync(kFull, acc.x, o); acc.y += __shfl_xor_sync(kFull, acc.y, o);
            acc.z += __shfl_xor_sync(kFull, acc.z, o); acc.w += __shfl_xor_sync(kFull, acc.w, o);
        }
        if (mine && g == 0) {
            const float4 val = make_float4(__fdiv_rn(acc.x, a.count), __fdiv_rn(acc.y, a.count), __fdiv_rn(acc.z, a.count),
                                           __fdiv_rn(acc.w, a.count));    // average over the sample grid (roi_align.py:192-196)
            *reinterpret_cast<float4*>(out0 + c) = val;
            if (out1) { out1[c] = val.x; out1[c + 1] = val.y; out1[c + 2] = val.z; out1[c + 3] = val.w; }
        }
    } else {
        pool_nhwc_scalar(p, a, wy, out0, out1);
    }
}

// ---------------------------------------------------------------------------------------------- gather kernel
// Counting sort of the boxes by (stride, class used) for the score kernel: position = prefix of the plan's histogram + an
// atomic ticket.  Runs in the gather kernels' prologue because the histogram is complete only after the plan kernel; the
// order inside a group is arbitrary (every box is scored independently).  The list holds OUTPUT slots.
__device__ __forceinline__ void sort_boxes_by_key(const FmapParams& p, int warp_g, int n_warps) {
    if (!p.cent) return;
    const int lane = threadIdx.x & 31;
    for (int box = warp_g; box < p.n; box += n_warps) {
        const int s = p.stride_idx[box];
        if (s < 0 || s > 2) continue;
        const int cu = p.cls_used[box];
        const int out = p.out_index[box];
        const int key = (cu >= 0 && cu < p.nc) ? s * p.nc + cu : 3 * p.nc;
        int before = 0;
        for (int i = lane; i < key; i += 32) before += p.hist[i];
        before = __reduce_add_sync(kFull, before);
        if (lane == 0) p.sorted[before + atomicAdd(&p.cursor[key], 1)] = out;
    }
}

__global__ void __launch_bounds__(kThreads, OODB200_FMAP_MIN_BLOCKS) items_kernel(const __grid_constant__ FmapParams p) {
    const int lane = threadIdx.x & 31;
    const int n_items = p.counters[0];
    // Scheduling: the first kStaticPct % of the work list is handed out round-robin (no atomics: same-address atomics
    // serialise in L2 and their latency under contention is of the order of an item), the tail through an atomic queue
    // so that the last items balance.  The next item's index and record are requested while the current one runs.
    const int warp_g = blockIdx.x * kWarps + (threadIdx.x >> 5), n_warps = gridDim.x * kWarps;
    sort_boxes_by_key(p, warp_g, n_warps);
    const int n_static = (int)((long long)n_items * OODB200_FMAP_STATIC_PCT / 100);
    int q_reg = 0;                                     // lane 0: result of the most recent queue fetch
    bool q_pending = false;
    auto next_index = [&](int cur) {                   // warp-uniform
        int nx = cur + n_warps;
        if (cur < n_static && nx < n_static) return nx;
        if (!q_pending) {                              // first dynamic fetch for this warp: blocking
            if (lane == 0) q_reg = atomicAdd(&p.counters[1], 1);
        }
        nx = n_static + __shfl_sync(kFull, q_reg, 0);
        if (lane == 0) q_reg = atomicAdd(&p.counters[1], 1);       // the one after, in flight during the next item
        q_pending = true;
        return nx;
    };
    int it = warp_g < n_static ? warp_g : n_items;
    if (it >= n_items && n_items > 0) it = next_index(n_static);   // more warps than static items: go to the queue
    int4 r0 = make_int4(0, 0, 0, 0), r1 = r0;
    if (it < n_items) { r0 = __ldg(p.items + 2 * (size_t)it); r1 = __ldg(p.items + 2 * (size_t)it + 1); }
    while (it < n_items) {
        const int nit = next_index(it);
        int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
        if (nit < n_items) { n0 = __ldg(p.items + 2 * (size_t)nit); n1 = __ldg(p.items + 2 * (size_t)nit + 1); }

        const int box = r0.x & 0xFFFFFF, sl = (int)((uint32_t)r0.x >> 24);
        const int s = r1.w & 3;
        const int C = p.C[s], W = p.W[s], HW = p.H[s] * W;
        const int4 geo = make_int4(r0.y & 0xFFFF, (int)((uint32_t)r0.y >> 16), r0.z & 0xFFFF, (int)((uint32_t)r0.z >> 16));
        const int out = r0.w;
        const int c_lo = sl * kSliceChannels, c_hi = min(C, c_lo + kSliceChannels);
        float* __restrict__ out0 = p.pooled + (size_t)out * p.pooled_ld;
        float* __restrict__ out1 = p.pooled_user ? p.pooled_user + (size_t)out * p.pooled_user_ld : nullptr;
        if (geo.w == 0) {                              // no sample inside the map (Q5): all-zero vector
            for (int c = c_lo + lane; c < c_hi; c += 32) { out0[c] = 0.f; if (out1) out1[c] = 0.f; }
        } else {
            const float* __restrict__ img = reinterpret_cast<const float*>(((unsigned long long)(uint32_t)r1.y << 32) | (uint32_t)r1.x);
            const float* __restrict__ wy = p.wts + (size_t)box * p.wstride;
            const float* __restrict__ wx = wy + p.ext_y;
            const float count = __int_as_float(r1.z);
            const bool vec = (W % 4 == 0) && (((uintptr_t)img & 15) == 0) && (HW % 4 == 0);
            if (vec) pool_dispatch(img, C, HW, W, geo, wy, wx, count, c_lo, c_hi, out0, out1);
            else pool_scalar(img, HW, W, geo.x, geo.y, geo.z, geo.w, wy, wx, count, c_lo, c_hi, out0, out1);
        }
        r0 = n0; r1 = n1; it = nit;
    }
}

// Channels-last gather: the same work list and scheduling as items_kernel (record of the next item requested while the
// current one runs).  Measured and not kept: building the next item's (offset, weight) pairs under the current item's
// cell loads with records fetched two items ahead -- 85 us (16 warps/SM) / 98 us (24 warps/SM, spills) against 78 us
// for this loop: the deeper look-ahead takes work out of the balancing queue early.
#ifndef OODB200_FMAP_NHWC_BLOCKS
#define OODB200_FMAP_NHWC_BLOCKS 3
#endif
__global__ void __launch_bounds__(kThreads, OODB200_FMAP_NHWC_BLOCKS) items_nhwc_kernel(const __grid_constant__ FmapParams p) {
    const int lane = threadIdx.x & 31, g = lane / kNhwcLanes, l = lane % kNhwcLanes;
    const int n_items = p.counters[0];
    const int warp_g = blockIdx.x * kWarps + (threadIdx.x >> 5), n_warps = gridDim.x * kWarps;
    sort_boxes_by_key(p, warp_g, n_warps);
    const int n_static = (int)((long long)n_items * OODB200_FMAP_STATIC_PCT / 100);
    int q_reg = 0;
    bool q_pending = false;
    auto next_index = [&](int cur) {                   // warp-uniform (see items_kernel)
        int nx = cur + n_warps;
        if (cur < n_static && nx < n_static) return nx;
        if (!q_pending) {
            if (lane == 0) q_reg = atomicAdd(&p.counters[1], 1);
        }
        nx = n_static + __shfl_sync(kFull, q_reg, 0);
        if (lane == 0) q_reg = atomicAdd(&p.counters[1], 1);
        q_pending = true;
        return nx;
    };
    int it = warp_g < n_static ? warp_g : n_items;
    if (it >= n_items && n_items > 0) it = next_index(n_static);
    int4 r0 = make_int4(0, 0, 0, 0), r1 = r0;
    if (it < n_items) { r0 = __ldg(p.items + 2 * (size_t)it); r1 = __ldg(p.items + 2 * (size_t)it + 1); }
    while (it < n_items) {
        const int nit = next_index(it);
        int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
        if (nit < n_items) { n0 = __ldg(p.items + 2 * (size_t)nit); n1 = __ldg(p.items + 2 * (size_t)nit + 1); }
        const NhwcItem a = nhwc_decode(r0, r1);
        nhwc_pool_item(p, a, p.wts + (size_t)a.box * p.wstride, p.pooled + (size_t)a.out * p.pooled_ld,
                       p.pooled_user ? p.pooled_user + (size_t)a.out * p.pooled_user_ld : nullptr);
        r0 = n0; r1 = n1; it = nit;
    }
}

// ---------------------------------------------------------------------------------------------- score kernel
// score_kernel: one CTA per unit = kUnitBoxes consecutive boxes of one (stride, class) group, one box per warp.
// The group's K centroid rows (and their unit-norm twins for cosine) are staged ONCE per CTA into shared memory with
// bulk-async copies (one mbarrier), while every warp already loads and normalises its first pooled vector into
// registers; the K rows are then swept from shared memory (128-bit, conflict-free), 4 rows per reduction round.  A warp's
// dependent global round trips are: histogram -> list entry -> pooled row (the centroid copy runs beside them).
// Tables larger than the shared-memory budget are staged in chunks of rows (the running best stays in registers); groups
// whose slices are not 16-byte aligned are swept from global memory (finalize) by the same CTAs.
constexpr int kUnitBoxes = kWarps;                     // boxes per score CTA: one per warp

__device__ __forceinline__ uint32_t fs_smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {   // a protocol error traps instead of hanging
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(fs_smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) __trap();
    }
}

// Rows [kb, kb + rc) of the group's centroid tables -> shared memory (one elected thread; completes on `bar`).
template <int MASK>
__device__ __forceinline__ void stage_rows(const FmapParams& p, int64_t off, int C, int kb, int rc, int RC, float* s_dyn, uint64_t* bar) {
    constexpr bool L12 = MASK & ((1 << OODB200_METRIC_L1) | (1 << OODB200_METRIC_L2));
    constexpr bool COS = MASK & (1 << OODB200_METRIC_COS);
    const uint32_t tab = (uint32_t)((size_t)rc * C * sizeof(float));
    float* sc = s_dyn;
    float* su = L12 ? s_dyn + (size_t)RC * C : s_dyn;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fs_smem_u32(bar)), "r"(tab * ((L12 ? 1u : 0u) + (COS ? 1u : 0u))) : "memory");
    if (L12)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fs_smem_u32(sc)),
                     "l"(p.cent + off + (int64_t)kb * C), "r"(tab), "r"(fs_smem_u32(bar)) : "memory");
    if (COS)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fs_smem_u32(su)),
                     "l"(p.cent_unit + off + (int64_t)kb * C), "r"(tab), "r"(fs_smem_u32(bar)) : "memory");
}

// One box against the K rows of its group, staged RC rows at a time (one chunk when the tables fit the shared-memory
// budget; the running best lives in registers across chunks).  NJ = float4 per lane (C <= 128 * NJ, only the last one can be partial), MASK = the
// requested metrics: no run-time flag or bound check inside the sweep.
template <int NJ, int MASK>
__device__ __forceinline__ void score_box_smem(const FmapParams& p, int s, int C, int cls, int K, int out, bool have,
                                               int64_t off, int RC, float* s_dyn, uint64_t* bar) {
    constexpr bool L1 = MASK & (1 << OODB200_METRIC_L1), L2 = MASK & (1 << OODB200_METRIC_L2);
    constexpr bool COS = MASK & (1 << OODB200_METRIC_COS), L12 = L1 || L2;
    constexpr int NM = (L1 ? 1 : 0) + (L2 ? 1 : 0) + (COS ? 1 : 0);
    const int lane = threadIdx.x & 31;
    const float* __restrict__ row = p.pooled + (size_t)out * p.pooled_ld;
    const bool last_ok = (NJ - 1) * 128 + lane * 4 < C;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* __restrict__ sc = s_dyn;
    const float* __restrict__ su = L12 ? s_dyn + (size_t)RC * C : s_dyn;
    float4 x[NJ];
#pragma unroll
    for (int t = 0; t < NJ; ++t)
        x[t] = (have && (t < NJ - 1 || last_ok)) ? __ldcg(reinterpret_cast<const float4*>(row + t * 128 + lane * 4)) : zero;
    float ss = 0.f;
#pragma unroll
    for (int t = 0; t < NJ; ++t) {
        ss = fmaf(x[t].x, x[t].x, ss); ss = fmaf(x[t].y, x[t].y, ss);
        ss = fmaf(x[t].z, x[t].z, ss); ss = fmaf(x[t].w, x[t].w, ss);
    }
    float n2v = 1.f;
    if (p.normalize) {                                 // ood_utils.py:2409 -> sklearn normalize
        float nrm = sqrtf(warp_sum(ss));
        if (nrm < 10.f * FLT_EPSILON) nrm = 1.f;       // _handle_zeros_in_scale
        ss = 0.f;
#pragma unroll
        for (int t = 0; t < NJ; ++t) {
            x[t].x = __fdiv_rn(x[t].x, nrm); x[t].y = __fdiv_rn(x[t].y, nrm);
            x[t].z = __fdiv_rn(x[t].z, nrm); x[t].w = __fdiv_rn(x[t].w, nrm);
            if (COS) {
                ss = fmaf(x[t].x, x[t].x, ss); ss = fmaf(x[t].y, x[t].y, ss);
                ss = fmaf(x[t].z, x[t].z, ss); ss = fmaf(x[t].w, x[t].w, ss);
            }
        }
    }
    if (COS) {                                         // cosine_distances re-normalises X (pairwise.py:1171-1182)
        n2v = sqrtf(warp_sum(ss));
        if (n2v < 10.f * FLT_EPSILON) n2v = 1.f;
    }
    Best b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-1, -1, -1}};
    int kb = 0;                                        // first row of the staged chunk
    auto batch = [&](int k0, int nr) {                 // staged rows k0 .. k0 + nr - 1 (nr <= 4): 4 x NM independent reductions
        float acc[4][NM];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float a1 = 0.f, a2 = 0.f, ac = 0.f;
            const int k = k0 + (r < nr ? r : 0);       // short batch: repeat row k0, result dropped
            const float* __restrict__ ck = sc + (size_t)k * C + lane * 4;
            const float* __restrict__ cu = su + (size_t)k * C + lane * 4;
#pragma unroll
            for (int t = 0; t < NJ; ++t) {
                const bool ld = t < NJ - 1 || last_ok;
                if (L12) {
                    const float4 c = ld ? *reinterpret_cast<const float4*>(ck + t * 128) : zero;
                    const float e0 = x[t].x - c.x, e1 = x[t].y - c.y, e2 = x[t].z - c.z, e3 = x[t].w - c.w;
                    if (L1) a1 += (fabsf(e0) + fabsf(e1)) + (fabsf(e2) + fabsf(e3));
                    if (L2) { a2 = fmaf(e0, e0, a2); a2 = fmaf(e1, e1, a2); a2 = fmaf(e2, e2, a2); a2 = fmaf(e3, e3, a2); }
                }
                if (COS) {
                    const float4 u = ld ? *reinterpret_cast<const float4*>(cu + t * 128) : zero;
                    ac = fmaf(x[t].x, u.x, ac); ac = fmaf(x[t].y, u.y, ac); ac = fmaf(x[t].z, u.z, ac); ac = fmaf(x[t].w, u.w, ac);
                }
            }
            int i = 0;
            if (L1) acc[r][i++] = a1;
            if (L2) acc[r][i++] = a2;
            if (COS) acc[r][i++] = ac;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < NM; ++i) acc[r][i] += __shfl_xor_sync(kFull, acc[r][i], o);
#pragma unroll
        for (int r = 0; r < 4; ++r) {                  // rows in increasing order, strict '<': first minimum
            if (r < nr) {
                const int k = kb + k0 + r;
                int i = 0;
                if (L1) { const float v = acc[r][i++]; if (v < b.d[0]) { b.d[0] = v; b.a[0] = k; } }
                if (L2) { const float v = sqrtf(fmaxf(acc[r][i++], 0.f)); if (v < b.d[1]) { b.d[1] = v; b.a[1] = k; } }
                if (COS) {                             // X / ||X|| applied to the sum (same value to float32 rounding)
                    const float v = fminf(fmaxf(1.0f - __fdiv_rn(acc[r][i++], n2v), 0.f), 2.f);
                    if (v < b.d[2]) { b.d[2] = v; b.a[2] = k; }
                }
            }
        }
    };
    uint32_t phase = 0;
    for (; kb < K; kb += RC) {
        const int rc = min(RC, K - kb);
        if (kb > 0) {
            __syncthreads();                           // every warp is done with the previous chunk
            if (threadIdx.x == 0) stage_rows<MASK>(p, off, C, kb, rc, RC, s_dyn, bar);
        }
        mbar_wait_bounded(bar, phase);                 // the chunk has landed
        phase ^= 1u;
        if (have) {
            int k0 = 0;
            for (; k0 + 4 <= rc; k0 += 4) batch(k0, 4);
            if (k0 < rc) batch(k0, rc - k0);
        }
    }
    if (have) write_result(p, s, cls, true, K, out, b);   // lane m reports metric m (all lanes hold the same result)
}

// Groups that cannot be staged: sweep from global memory, warp-private vector in shared memory (kept out of line: the
// staged path stays small in the instruction cache).
__device__ __noinline__ void score_unit_global(const FmapParams& p, int s, int cls, int out, float* xs) {
    finalize(p, s, cls, out, xs);
}

template <int MASK>
__global__ void __launch_bounds__(kThreads, 3) score_kernel(const __grid_constant__ FmapParams p, int smem_bytes) {
    extern __shared__ __align__(128) float s_dyn[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_unit[4];                          // key, chunk, start, boxes of the group
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fs_smem_u32(&s_bar)), "r"(1) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // unit u = blockIdx.x -> (group key, chunk of kUnitBoxes boxes inside the group): blocked prefix over the histogram
        const int nkeys = 3 * p.nc + 1, q = (nkeys + 31) >> 5;
        int ub = 0, bb = 0;
        for (int i = 0; i < q; ++i) {
            const int k = lane * q + i;
            if (k < nkeys) { const int h = p.hist[k]; ub += (h + kUnitBoxes - 1) / kUnitBoxes; bb += h; }
        }
        int ue = ub, be = bb;                          // inclusive scans over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tu = __shfl_up_sync(kFull, ue, o), tb = __shfl_up_sync(kFull, be, o);
            if (lane >= o) { ue += tu; be += tb; }
        }
        const int u = blockIdx.x;
        if (lane == 0) s_unit[0] = -1;                 // beyond the last unit
        __syncwarp();
        if (u >= ue - ub && u < ue) {                  // exactly one lane owns the unit
            int uacc = ue - ub, bacc = be - bb;
            for (int i = 0; i < q; ++i) {
                const int k = lane * q + i;
                const int h = k < nkeys ? p.hist[k] : 0;
                const int nu = (h + kUnitBoxes - 1) / kUnitBoxes;
                if (u < uacc + nu) { s_unit[0] = k; s_unit[1] = u - uacc; s_unit[2] = bacc; s_unit[3] = h; break; }
                uacc += nu; bacc += h;
            }
        }
    }
    __syncthreads();
    const int key = s_unit[0];
    if (key < 0) return;
    const int chunk = s_unit[1], start = s_unit[2], m = s_unit[3];
    const int bi = chunk * kUnitBoxes + warp;          // this warp's box inside the group
    const bool have = bi < m;
    const int out = have ? p.sorted[start + bi] : 0;
    const bool cls_ok = key < 3 * p.nc;
    const int s = cls_ok ? key / p.nc : 0;
    const int cls = cls_ok ? key - s * p.nc : -1;
    if (!cls_ok) {                                     // class outside [0, nc): no clusters, no threshold
        const Best b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-1, -1, -1}};
        if (have) write_result(p, 0, -1, false, 0, out, b);
        return;
    }
    const int C = p.C[s];
    const int K = p.cent_k[key];
    const int64_t off = K > 0 ? p.cent_off[key] : 0;
    constexpr bool L12 = MASK & ((1 << OODB200_METRIC_L1) | (1 << OODB200_METRIC_L2));
    constexpr bool COS = MASK & (1 << OODB200_METRIC_COS);
    const int ntab = (L12 ? 1 : 0) + (COS ? 1 : 0);
    const int rows_fit = (int)((size_t)smem_bytes / ((size_t)C * sizeof(float) * ntab));
    const int RC = min(K, min(rows_fit, (int)((1u << 19) / ((size_t)C * sizeof(float)))));   // rows per staged chunk
    const int nj = (C + 127) >> 7;
    const bool staged = K > 0 && C % 4 == 0 && nj <= 8 && RC >= 1 &&
                        off % 4 == 0 && ((((uintptr_t)p.cent) | ((uintptr_t)p.cent_unit)) & 15) == 0;
    if (!staged) {
        if (have) score_unit_global(p, s, cls, out, s_dyn + (size_t)warp * p.pooled_ld);
        return;
    }
    if (threadIdx.x == 0) stage_rows<MASK>(p, off, C, 0, RC, RC, s_dyn, &s_bar);
    switch (nj) {                                      // block-uniform: every warp runs the same chunk loop (barriers inside)
        case 1: score_box_smem<1, MASK>(p, s, C, cls, K, out, have, off, RC, s_dyn, &s_bar); break;
        case 2: score_box_smem<2, MASK>(p, s, C, cls, K, out, have, off, RC, s_dyn, &s_bar); break;
        case 3: score_box_smem<3, MASK>(p, s, C, cls, K, out, have, off, RC, s_dyn, &s_bar); break;
        case 4: score_box_smem<4, MASK>(p, s, C, cls, K, out, have, off, RC, s_dyn, &s_bar); break;
        case 5: score_box_smem<5, MASK>(p, s, C, cls, K, out, have, off, RC, s_dyn, &s_bar); break;
        case 6: score_box_smem<6, MASK>(p, s, C, cls, K, out, have, off, RC, s_dyn, &s_bar); break;
        default: score_box_smem<8, MASK>(p, s, C, cls, K, out, have, off, RC, s_dyn, &s_bar); break;
    }
}

typedef void (*ScoreKernel)(const FmapParams, int);
static ScoreKernel score_kernel_for(int mask) {
    switch (mask) {
        case 1: return score_kernel<1>;
        case 2: return score_kernel<2>;
        case 3: return score_kernel<3>;
        case 4: return score_kernel<4>;
        case 5: return score_kernel<5>;
        case 6: return score_kernel<6>;
        default: return score_kernel<7>;
    }
}

// =============================================================================================== channels-last pipeline
// pipe_nhwc_kernel: ONE launch for the whole path on channels-last maps when the centroid tables are small enough to be
// swept from L2 per box (the usual K <= 16).  In [H, W, C] memory a window row is ONE contiguous run of ww * C floats
// (3 .. 40 KB): the window is staged into shared memory by the copy engine (cp.async.bulk, TMA) in pieces of <= 7 KB, and
// every fetched 128-byte line is used in full.
// Persistent grid; a CTA is kPipeGroups independent two-warp pipelines, each with its own 3-stage ring (8 pipelines = 144 KB
// of copies in flight per SM):
//   planner warp   next box from the atomic queue; quirk-Q1 class / output slot from ballots over the image's stride list;
//                  RoIAlign geometry + separable weights into one of two descriptor slots (it runs ONE BOX AHEAD of the
//                  consumer); then one bulk copy per piece of the window as ring stages free up.
//   consumer warp  waits for a piece (mbarrier), adds w[cell] * v[cell][c] for its channels (lane = 4 channels x NJ, one
//                  conflict-free LDS.128 per cell, no cross-lane reduction, fixed cell order), frees the stage; after the
//                  last piece it normalises the vector and sweeps the K centroid rows from L2 (4 rows in flight), then
//                  writes distance / arg-min / decision.  Planning, copies and scoring of consecutive boxes overlap.
// Same per-lane partition and reduction tree as score_box_smem / vec_score_fast_kernel: fit-time and decision-time distances
// of the same vector are bit-identical.
constexpr int kPipeGroups = 4;                         // two-warp pipelines per CTA
constexpr int kPipeThreads = kPipeGroups * 64;
constexpr int kPipeStages = 3;
constexpr int kPipeStageBytes = 7168;
constexpr int kPipeMaxNJ = 5;                          // C <= 640 (beyond: the plan / gather / score sequence)

struct PipeDesc {                                      // one box, written by the planner, read by the consumer
    int box, s, cls_u, out;                            // box < 0: no more work
    int y0, x0, wh, ww, xoff;                          // first live row / column, live rows / columns
    float count;
    int direct;                                        // 1: map not 16-byte aligned -> the consumer pools from global memory
    unsigned long long mapp;
};

__device__ __forceinline__ void mb_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fs_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fs_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fs_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fs_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(fs_smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kPipeThreads, 2) pipe_nhwc_kernel(const __grid_constant__ FmapParams p, int group_bytes) {
    extern __shared__ __align__(128) unsigned char psm[];
    __shared__ __align__(8) uint64_t s_bar[kPipeGroups][2 * kPipeStages + 4];   // ring full / empty, descriptor full / empty
    __shared__ PipeDesc s_desc[kPipeGroups][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp >> 1;
    const bool planner = (warp & 1) == 0;
    unsigned char* gbase = psm + (size_t)grp * group_bytes;
    float* ring = reinterpret_cast<float*>(gbase);                               // kPipeStages x kPipeStageBytes
    float* s_wts = reinterpret_cast<float*>(gbase + kPipeStages * kPipeStageBytes);   // 2 x wstride: wy[ext_y] | wx
    float* xs = s_wts + 2 * (size_t)p.wstride;                                   // pooled_ld floats
    uint64_t* ring_full = &s_bar[grp][0];
    uint64_t* ring_empty = &s_bar[grp][kPipeStages];
    uint64_t* desc_full = &s_bar[grp][2 * kPipeStages];
    uint64_t* desc_empty = &s_bar[grp][2 * kPipeStages + 2];
    if (threadIdx.x == 0) {
        for (int g = 0; g < kPipeGroups; ++g)
            for (int i = 0; i < 2 * kPipeStages + 4; ++i) mb_init(&s_bar[g][i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (planner) {
        int pc = 0;                                    // pieces issued by this pipeline so far
        for (int bc = 0;; ++bc) {
            const int slot = bc & 1;
            int b = 0;
            if (lane == 0) b = atomicAdd(&p.counters[1], 1);
            b = __shfl_sync(kFull, b, 0);
            const bool more = b < p.n;
            int img = 0, s = -1;
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            if (more) {
                img = p.img_idx[b];
                s = p.stride_idx[b];
                bx = *reinterpret_cast<const float4*>(p.boxes + 4 * (size_t)b);
            }
            const bool ok = more && s >= 0 && s <= 2;
            int cls_u = 0, out = b;
            if (more) {                                // quirk Q1 (ood_utils.py:2152-2154) from the image's own stride list
                const int b0 = p.img_start[img], m = p.img_start[img + 1] - b0, bb = b - b0;
                const int mycat = ok ? s : 3;
                int cnt[4] = {0, 0, 0, 0};
                int j = 0;
                for (int c0 = 0; c0 < m; c0 += 32) {
                    const int e = c0 + lane;
                    const int st = e < m ? p.stride_idx[b0 + e] : -2;
                    const int cat = e < m ? ((st >= 0 && st <= 2) ? st : 3) : 4;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const unsigned mk = __ballot_sync(kFull, cat == t);
                        cnt[t] += __popc(mk);
                        if (t == mycat) {
                            if (c0 + 32 <= bb) j += __popc(mk);
                            else if (c0 <= bb) j += __popc(mk & ((1u << (bb - c0)) - 1u));
                        }
                    }
                }
                int before = 0;
#pragma unroll
                for (int t = 0; t < 3; ++t) if (t < mycat) before += cnt[t];
                cls_u = p.cls ? p.cls[b] : 0;
                if (p.compat_q1) {
                    cls_u = ok ? p.cls[b0 + j] : -1;
                    out = b0 + before + j;
                }
            }
            if (bc >= 2) mbar_wait_bounded(&desc_empty[slot], (uint32_t)((bc >> 1) - 1) & 1u);   // the consumer is done with this slot
            float* wy = s_wts + (size_t)slot * p.wstride;
            BoxGeo g = {0, 0, 0, 0, 1.f, 0, 0};
            unsigned long long mapp = 0;
            if (ok) {
                g = geo_compute_into(p, s, bx, wy, wy + p.ext_y);
                mapp = (unsigned long long)p.map_ptrs[img * 3 + s];
            }
            const int C = ok ? p.C[s] : 4;
            const int direct = ok && ((mapp & 15) != 0 || (C & 3) != 0);
            if (lane == 0) {
                PipeDesc d;
                d.box = more ? b : -1; d.s = s; d.cls_u = cls_u; d.out = out;
                d.y0 = g.ylo; d.x0 = g.xa + g.xoff; d.wh = g.nxc ? g.wh : 0; d.ww = g.ww; d.xoff = g.xoff;
                d.count = g.count; d.direct = direct; d.mapp = mapp;
                s_desc[grp][slot] = d;
                if (more) { p.cls_used[b] = cls_u; p.out_index[b] = out; }
            }
            __syncwarp();
            if (lane == 0) mb_arrive(&desc_full[slot]);
            if (!more) break;
            if (ok && !direct && g.nxc && lane == 0) {                            // the window, row by row, in pieces of <= cps cells
                const int W = p.W[s];
                const int cps = kPipeStageBytes / (C * 4);
                const float* __restrict__ src = reinterpret_cast<const float*>(mapp);
                for (int r = 0; r < g.wh; ++r)
                    for (int a = 0; a < g.ww; a += cps) {
                        const int st = pc % kPipeStages;
                        if (pc >= kPipeStages) mbar_wait_bounded(&ring_empty[st], (uint32_t)((pc / kPipeStages) - 1) & 1u);
                        const uint32_t bytes = (uint32_t)min(cps, g.ww - a) * (uint32_t)C * 4u;
                        mb_expect_tx(&ring_full[st], bytes);
                        bulk_g2s(ring + (size_t)st * (kPipeStageBytes / 4),
                                 src + ((size_t)(g.ylo + r) * W + (g.xa + g.xoff + a)) * C, bytes, &ring_full[st]);
                        ++pc;
                    }
            }
            pc = __shfl_sync(kFull, pc, 0);
        }
        return;
    }
    // ---- consumer warp
    int pc = 0;
    for (int bc = 0;; ++bc) {
        const int slot = bc & 1;
        mbar_wait_bounded(&desc_full[slot], (uint32_t)(bc >> 1) & 1u);
        const PipeDesc d = s_desc[grp][slot];
        if (d.box < 0) break;
        const int s = d.s, out = d.out;
        if (s < 0 || s > 2) {                          // never pooled by the reference either: answered here
            if (p.cent && lane < OODB200_N_METRICS && (p.metric_mask >> lane & 1)) {
                const size_t o = (size_t)lane * p.n + out;
                p.dist[o] = nanf("");
                p.argmin[o] = -1;
                p.decision[o] = 0;
            }
            __syncwarp();
            if (lane == 0) mb_arrive(&desc_empty[slot]);
            continue;
        }
        const int C = p.C[s];
        const int nj = (C + 127) >> 7;
        const float* __restrict__ wy = s_wts + (size_t)slot * p.wstride;
        const float* __restrict__ wx = wy + p.ext_y + d.xoff;
        float* __restrict__ out1 = p.pooled_user ? p.pooled_user + (size_t)out * p.pooled_user_ld : nullptr;
        if (d.direct) {                                // odd alignment / channel count: lanes over channels, scalar loads
            NhwcItem a;
            a.img = reinterpret_cast<const float*>(d.mapp); a.box = d.box; a.s = s; a.out = out; a.y0 = d.y0; a.x0 = d.x0;
            a.wh = d.wh; a.ww = d.ww; a.xoff = d.xoff; a.count = d.count;
            for (a.c_lo = 0; a.c_lo < C; a.c_lo += kSliceNhwc) pool_nhwc_scalar(p, a, wy, xs, out1);
        } else {
            float4 acc[kPipeMaxNJ];
#pragma unroll
            for (int t = 0; t < kPipeMaxNJ; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int cps = kPipeStageBytes / (C * 4);
            for (int r = 0; r < d.wh; ++r) {
                const float wr = wy[r];
                for (int a = 0; a < d.ww; a += cps) {
                    const int st = pc % kPipeStages;
                    mbar_wait_bounded(&ring_full[st], (uint32_t)(pc / kPipeStages) & 1u);
                    const float* __restrict__ piece = ring + (size_t)st * (kPipeStageBytes / 4) + lane * 4;
                    const int nc_ = min(cps, d.ww - a);
                    for (int x = 0; x < nc_; ++x) {
                        const float w = wr * wx[a + x];
#pragma unroll
                        for (int t = 0; t < kPipeMaxNJ; ++t)
                            if (t < nj && lane * 4 + 128 * t < C) {
                                const float4 v = *reinterpret_cast<const float4*>(piece + (size_t)x * C + 128 * t);
                                acc[t].x = fmaf(w, v.x, acc[t].x); acc[t].y = fmaf(w, v.y, acc[t].y);
                                acc[t].z = fmaf(w, v.z, acc[t].z); acc[t].w = fmaf(w, v.w, acc[t].w);
                            }
                    }
                    __syncwarp();
                    if (lane == 0) mb_arrive(&ring_empty[st]);
                    ++pc;
                }
            }
#pragma unroll
            for (int t = 0; t < kPipeMaxNJ; ++t) {
                const int c = lane * 4 + 128 * t;
                if (t < nj && c < C) {                 // average over the sample grid (roi_align.py:192-196); no sample inside the map (Q5): 0
                    const float4 val = d.wh ? make_float4(__fdiv_rn(acc[t].x, d.count), __fdiv_rn(acc[t].y, d.count),
                                                          __fdiv_rn(acc[t].z, d.count), __fdiv_rn(acc[t].w, d.count))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(xs + c) = val;
                    if (out1) { out1[c] = val.x; out1[c + 1] = val.y; out1[c + 2] = val.z; out1[c + 3] = val.w; }
                }
            }
        }
        __syncwarp();
        if (p.cent) {
            // ---- distance to the nearest centroid of (class used, stride), threshold: one warp, 4 rows in flight
            const int cls_u = d.cls_u;
            const bool cls_ok = cls_u >= 0 && cls_u < p.nc;
            const int K = cls_ok ? p.cent_k[s * p.nc + cls_u] : 0;
            const int64_t off = K > 0 ? p.cent_off[s * p.nc + cls_u] : 0;
            const bool vecs = (C % 4 == 0) && (off % 4 == 0) && ((((uintptr_t)p.cent) | ((uintptr_t)p.cent_unit)) & 15) == 0;
            if (!vecs) {                               // odd shapes: through the workspace row, like finalize()
                float* row = p.pooled + (size_t)out * p.pooled_ld;
                for (int c = lane; c < C; c += 32) row[c] = xs[c];
                __threadfence_block();
                __syncwarp();
                const Best bb = finalize_generic(row, C, p.normalize, p.metric_mask, p.cent + off, p.cent_unit + off, K);
                write_result(p, s, cls_u, cls_ok, K, out, bb);
            } else {
                const bool want_l1 = p.metric_mask & (1 << OODB200_METRIC_L1);
                const bool want_l2 = p.metric_mask & (1 << OODB200_METRIC_L2);
                const bool want_cos = p.metric_mask & (1 << OODB200_METRIC_COS);
                float4 x[kPipeMaxNJ];
                float ss = 0.f;
#pragma unroll
                for (int t = 0; t < kPipeMaxNJ; ++t) {
                    const int c = lane * 4 + 128 * t;
                    x[t] = (t < nj && c < C) ? *reinterpret_cast<const float4*>(xs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    ss = fmaf(x[t].x, x[t].x, ss); ss = fmaf(x[t].y, x[t].y, ss); ss = fmaf(x[t].z, x[t].z, ss); ss = fmaf(x[t].w, x[t].w, ss);
                }
                float n2v = 1.f;
                if (p.normalize) {                     // ood_utils.py:2409 -> sklearn normalize
                    float nrm = sqrtf(warp_sum(ss));
                    if (nrm < 10.f * FLT_EPSILON) nrm = 1.f;       // _handle_zeros_in_scale
                    ss = 0.f;
#pragma unroll
                    for (int t = 0; t < kPipeMaxNJ; ++t) {
                        x[t].x = __fdiv_rn(x[t].x, nrm); x[t].y = __fdiv_rn(x[t].y, nrm);
                        x[t].z = __fdiv_rn(x[t].z, nrm); x[t].w = __fdiv_rn(x[t].w, nrm);
                        ss = fmaf(x[t].x, x[t].x, ss); ss = fmaf(x[t].y, x[t].y, ss); ss = fmaf(x[t].z, x[t].z, ss); ss = fmaf(x[t].w, x[t].w, ss);
                    }
                }
                if (want_cos) {                        // cosine_distances re-normalises X (pairwise.py:1171-1182)
                    n2v = sqrtf(warp_sum(ss));
                    if (n2v < 10.f * FLT_EPSILON) n2v = 1.f;
                }
                Best bst = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-1, -1, -1}};
                for (int k0 = 0; k0 < K; k0 += 4) {    // rows k0 .. k0 + 3: independent reductions
                    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f}, ac[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int t = 0; t < kPipeMaxNJ; ++t) {
                        const int c = lane * 4 + 128 * t;
                        if (t < nj && c < C) {
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const int k = min(k0 + r, K - 1);                 // short batch: repeat the last row, result dropped
                                if (want_l1 || want_l2) {
                                    const float4 cc = __ldg(reinterpret_cast<const float4*>(p.cent + off + (int64_t)k * C + c));
                                    const float e0 = x[t].x - cc.x, e1 = x[t].y - cc.y, e2 = x[t].z - cc.z, e3 = x[t].w - cc.w;
                                    a1[r] += (fabsf(e0) + fabsf(e1)) + (fabsf(e2) + fabsf(e3));
                                    a2[r] = fmaf(e0, e0, a2[r]); a2[r] = fmaf(e1, e1, a2[r]); a2[r] = fmaf(e2, e2, a2[r]); a2[r] = fmaf(e3, e3, a2[r]);
                                }
                                if (want_cos) {
                                    const float4 u = __ldg(reinterpret_cast<const float4*>(p.cent_unit + off + (int64_t)k * C + c));
                                    ac[r] = fmaf(x[t].x, u.x, ac[r]); ac[r] = fmaf(x[t].y, u.y, ac[r]);
                                    ac[r] = fmaf(x[t].z, u.z, ac[r]); ac[r] = fmaf(x[t].w, u.w, ac[r]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (want_l1) a1[r] += __shfl_xor_sync(kFull, a1[r], o);
                            if (want_l2) a2[r] += __shfl_xor_sync(kFull, a2[r], o);
                            if (want_cos) ac[r] += __shfl_xor_sync(kFull, ac[r], o);
                        }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {      // rows in increasing order, strict '<': first minimum
                        const int k = k0 + r;
                        if (k < K) {
                            if (want_l1 && a1[r] < bst.d[0]) { bst.d[0] = a1[r]; bst.a[0] = k; }
                            if (want_l2) { const float v = sqrtf(fmaxf(a2[r], 0.f)); if (v < bst.d[1]) { bst.d[1] = v; bst.a[1] = k; } }
                            if (want_cos) {            // X / ||X|| applied to the sum (same value to float32 rounding)
                                const float v = fminf(fmaxf(1.0f - __fdiv_rn(ac[r], n2v), 0.f), 2.f);
                                if (v < bst.d[2]) { bst.d[2] = v; bst.a[2] = k; }
                            }
                        }
                    }
                }
                write_result(p, s, cls_u, cls_ok, K, out, bst);
            }
        }
        __syncwarp();
        if (lane == 0) mb_arrive(&desc_empty[slot]);
    }
}

struct WorkspaceLayout {
    size_t counters, hist, cursor, sorted, cls_used, out_index, ipos, wts, items, pooled, total;
    size_t zero_bytes;
    int ext_y, wstride, pooled_ld, max_ns;
};

static WorkspaceLayout layout_of(int n, int nc, const int32_t* map_chw) {
    WorkspaceLayout L;
    int hmax = 1, wmax = 1, cmax = 1;
    for (int s = 0; s < 3; ++s) {
        cmax = max(cmax, map_chw[3 * s]);
        hmax = max(hmax, map_chw[3 * s + 1]);
        wmax = max(wmax, map_chw[3 * s + 2]);
    }
    L.ext_y = (hmax + 3) & ~3;
    L.wstride = L.ext_y + ((wmax + 3) & ~3) + 4;
    L.pooled_ld = (cmax + 3) & ~3;
    constexpr int kFine = kSliceNhwc < kSliceChannels ? kSliceNhwc : kSliceChannels;   // the finer of the two slicings
    L.max_ns = (cmax + kFine - 1) / kFine;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    const size_t nk = 3 * (size_t)(nc > 0 ? nc : 1) + 1;
    L.counters = o; o += 16;                          // counters, hist and cursor are zeroed by one memset
    L.hist = o; o += 4 * nk;
    L.cursor = o; o += 4 * nk;
    L.zero_bytes = o;
    o = up(o);
    L.sorted = o; o = up(o + 4 * nn);
    L.cls_used = o; o = up(o + 4 * nn);
    L.out_index = o; o = up(o + 4 * nn);
    L.ipos = o; o = up(o + 8 * nn);
    L.wts = o; o = up(o + 4 * nn * L.wstride);
    L.items = o; o = up(o + 32 * nn * L.max_ns);
    L.pooled = o; o = up(o + 4 * nn * L.pooled_ld);
    L.total = o;
    return L;
}

struct DeviceInfo { bool init; int sms, items_per_sm, items_per_sm_nhwc; size_t score_attr[8], pipe_attr; };
static DeviceInfo g_dev[64];

static DeviceInfo& device_info() {                     // per device: SM count, occupancy, attributes already raised
    int dev = 0;
    cudaGetDevice(&dev);
    DeviceInfo& d = g_dev[dev & 63];
    if (!d.init) {
        if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || d.sms <= 0) d.sms = 148;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.items_per_sm, items_kernel, kThreads, 0) != cudaSuccess || d.items_per_sm <= 0)
            d.items_per_sm = OODB200_FMAP_MIN_BLOCKS;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.items_per_sm_nhwc, items_nhwc_kernel, kThreads, 0) != cudaSuccess ||
            d.items_per_sm_nhwc <= 0)
            d.items_per_sm_nhwc = OODB200_FMAP_MIN_BLOCKS;
        d.init = true;
    }
    return d;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int launch_fmap(FmapParams& p, const int32_t* map_chw, const float* scale, int32_t* cls_used_out,
                       int32_t* out_index_out, void* workspace, int64_t workspace_bytes, void* stream, const char* what) {
    for (int s = 0; s < 3; ++s) {
        p.C[s] = map_chw[3 * s];
        p.H[s] = map_chw[3 * s + 1];
        p.W[s] = map_chw[3 * s + 2];
        p.scale[s] = scale[s];
        OODB200_REQUIRE(p.C[s] > 0 && p.H[s] > 0 && p.W[s] > 0, "%s: map %d has non-positive shape", what, s);
        p.slice = p.nhwc ? kSliceNhwc : kSliceChannels;
        p.ns[s] = (p.C[s] + p.slice - 1) / p.slice;
        OODB200_REQUIRE(p.ns[s] <= kMaxSlices, "%s: map %d has too many channels (%d)", what, s, p.C[s]);
        OODB200_REQUIRE((long long)p.H[s] * p.W[s] < (1LL << 30) && p.H[s] < 65536 && p.W[s] < 65536, "%s: map %d too large", what, s);
        OODB200_REQUIRE(!p.nhwc || (long long)p.H[s] * p.W[s] * p.C[s] < (1LL << 31), "%s: channels-last map %d too large", what, s);
    }
    OODB200_REQUIRE(p.n < (1 << 24), "%s: at most %d boxes per call", what, (1 << 24) - 1);
    if (p.n == 0) return OODB200_OK;
    const WorkspaceLayout L = layout_of(p.n, p.cent ? p.nc : 0, map_chw);
    OODB200_REQUIRE(workspace && workspace_bytes >= (int64_t)L.total, "%s: workspace too small (%lld < %zu bytes)", what,
                    (long long)workspace_bytes, L.total);
    OODB200_REQUIRE(((uintptr_t)workspace & 255) == 0, "%s: workspace must be 256-byte aligned", what);
    char* ws = (char*)workspace;
    p.counters = (int*)(ws + L.counters);
    p.hist = (int*)(ws + L.hist);
    p.cursor = (int*)(ws + L.cursor);
    p.sorted = (int32_t*)(ws + L.sorted);
    p.cls_used = cls_used_out ? cls_used_out : (int32_t*)(ws + L.cls_used);
    p.out_index = out_index_out ? out_index_out : (int32_t*)(ws + L.out_index);
    p.ipos = (int2*)(ws + L.ipos);
    p.wts = (float*)(ws + L.wts);
    p.ext_y = L.ext_y;
    p.wstride = L.wstride;
    p.items = (int4*)(ws + L.items);
    p.pooled = (float*)(ws + L.pooled);
    p.pooled_ld = L.pooled_ld;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceInfo& dv = device_info();
    cudaError_t e;
    int rc;
    static const int force_group = env_int("OODB200_FMAP_GROUP_SCORE", -1);    // 1 / 0: override the caller's mode (A/B runs)
    if (force_group >= 0) p.group_score = force_group;
    static const int no_pipe = env_int("OODB200_FMAP_NO_PIPE", 0);              // 1: the plan / gather / score sequence (A/B runs)
    int cmax = 0;
    for (int s = 0; s < 3; ++s) cmax = max(cmax, p.C[s]);
    const size_t group_bytes = ((size_t)kPipeStages * kPipeStageBytes + sizeof(float) * (2 * (size_t)L.wstride + L.pooled_ld) + 127) & ~(size_t)127;
    const size_t pipe_smem = group_bytes * kPipeGroups;
    if (p.nhwc && !p.group_score && !no_pipe && cmax <= 128 * kPipeMaxNJ && pipe_smem <= 110 * 1024) {
        // ---- channels-last, small tables: one persistent launch, windows staged by TMA, plan / pool / score pipelined per box
        e = cudaMemsetAsync(p.counters, 0, 16, st);
        if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
        if (pipe_smem > dv.pipe_attr) {
            e = cudaFuncSetAttribute(pipe_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem);
            if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
            dv.pipe_attr = pipe_smem;
        }
        long long ctas = 2LL * dv.sms;
        const long long need = ((long long)p.n + kPipeGroups - 1) / kPipeGroups;
        if (ctas > need) ctas = need;
        pipe_nhwc_kernel<<<(int)ctas, kPipeThreads, pipe_smem, st>>>(p, (int)group_bytes);
        return check_launch(what);
    }
    // ---- large tables: memset -> plan + geometry -> gather -> score by (stride, class) groups
    e = cudaMemsetAsync(p.counters, 0, L.zero_bytes, st);
    if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    {
        int G = 1;                                                 // CTAs per image: one warp per box in one round if possible
        while (G < 8 && (long long)G * kPgWarps * p.n_img < p.n) G *= 2;
        static const int g_force = env_int("OODB200_PLAN_CLUSTER", 0);   // tuning override (1, 2, 4 or 8)
        if (g_force == 1 || g_force == 2 || g_force == 4 || g_force == 8) G = g_force;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(p.n_img * G), 1, 1);
        cfg.blockDim = dim3(kPgThreads, 1, 1);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)G;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, plan_geo_kernel, p);
        if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
        rc = check_launch(what);
        if (rc) return rc;
    }
    long long max_items = (long long)p.n * L.max_ns;
    long long grid = (long long)dv.sms * (p.nhwc ? dv.items_per_sm_nhwc : dv.items_per_sm);   // persistent: every resident warp pulls from the queue
    if (grid * kWarps > max_items) grid = (max_items + kWarps - 1) / kWarps;
    if (p.nhwc) items_nhwc_kernel<<<(int)grid, kThreads, 0, st>>>(p);
    else items_kernel<<<(int)grid, kThreads, 0, st>>>(p);
    rc = check_launch(what);
    if (rc) return rc;
    if (p.cent) {
        // shared memory per score CTA: the centroid tables of one (stride, class) group when they fit this budget
        // (OODB200_SCORE_SMEM_KB, default 72 KB = 3 CTAs per SM), at least the warp-private vectors of the global sweep
        static const int budget_env = env_int("OODB200_SCORE_SMEM_KB", 72);
        const int budget_kb = budget_env < 16 ? 16 : (budget_env > 200 ? 200 : budget_env);
        size_t smem = sizeof(float) * (size_t)kWarps * L.pooled_ld;
        OODB200_REQUIRE(smem <= 200 * 1024, "%s: too many channels for the score kernel (%d)", what, L.pooled_ld);
        if (smem < (size_t)budget_kb * 1024) smem = (size_t)budget_kb * 1024;
        const ScoreKernel kern = score_kernel_for(p.metric_mask);
        if (smem > 48 * 1024 && smem > dv.score_attr[p.metric_mask & 7]) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
            dv.score_attr[p.metric_mask & 7] = smem;
        }
        const int nkeys = 3 * p.nc + 1;
        const int units = (p.n + kUnitBoxes - 1) / kUnitBoxes + nkeys;   // >= sum over groups of ceil(m / kUnitBoxes)
        kern<<<units, kThreads, smem, st>>>(p, (int)smem);
        rc = check_launch(what);
    }
    return rc;
}

}  // namespace oodb200

using namespace oodb200;

extern "C" int64_t oodb200_fmap_workspace_bytes(int n, int nc, const int32_t* map_chw) {
    if (!map_chw || nc < 0) return -1;
    return (int64_t)layout_of(n, nc, map_chw).total;
}

static int roi_pool_impl(int nhwc, const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                         const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                         const int32_t* img_start, int n,
                         float* out, int out_ld, void* workspace, int64_t workspace_bytes, void* stream) {
    OODB200_REQUIRE(n >= 0 && n_img >= 0, "roi_pool: negative size");
    OODB200_REQUIRE(map_chw && scale, "roi_pool: map_chw/scale must be host arrays");
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(map_ptrs && boxes && img_idx && stride_idx && img_start && out, "roi_pool: null pointer");
    FmapParams p = {};
    p.map_ptrs = map_ptrs; p.boxes = boxes; p.img_idx = img_idx; p.stride_idx = stride_idx; p.img_start = img_start;
    p.n = n; p.n_img = n_img; p.pooled_user = out; p.pooled_user_ld = out_ld; p.nhwc = nhwc;
    return launch_fmap(p, map_chw, scale, nullptr, nullptr, workspace, workspace_bytes, stream, "roi_pool");
}

extern "C" int oodb200_roi_pool_f32(const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                                    const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                                    const int32_t* img_start, int n,
                                    float* out, int out_ld, void* workspace, int64_t workspace_bytes, void* stream) {
    return roi_pool_impl(0, map_ptrs, map_chw, scale, n_img, boxes, img_idx, stride_idx, img_start, n, out, out_ld, workspace,
                         workspace_bytes, stream);
}

extern "C" int oodb200_roi_pool_nhwc_f32(const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                                         const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                                         const int32_t* img_start, int n,
                                         float* out, int out_ld, void* workspace, int64_t workspace_bytes, void* stream) {
    return roi_pool_impl(1, map_ptrs, map_chw, scale, n_img, boxes, img_idx, stride_idx, img_start, n, out, out_ld, workspace,
                         workspace_bytes, stream);
}

static int fmap_score_impl(int nhwc, const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                           const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                           const int32_t* cls, const int32_t* img_start, int flags, int n,
                           int metric_mask, int normalize,
                           const float* cent, const float* cent_unit, const int64_t* cent_off,
                           const int32_t* cent_k, int nc, const double* thr,
                           float* dist, int32_t* argmin, uint8_t* decision,
                           float* pooled, int pooled_ld, int32_t* cls_used_out, int32_t* out_index_out,
                           void* workspace, int64_t workspace_bytes, void* stream) {
    OODB200_REQUIRE(n >= 0 && n_img >= 0 && nc > 0, "fmap_score: bad size");
    OODB200_REQUIRE(map_chw && scale, "fmap_score: map_chw/scale must be host arrays");
    OODB200_REQUIRE(metric_mask > 0 && metric_mask < (1 << OODB200_N_METRICS), "fmap_score: metric_mask %d", metric_mask);
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(map_ptrs && boxes && img_idx && stride_idx && cls && img_start, "fmap_score: null input pointer");
    OODB200_REQUIRE(cent && cent_off && cent_k && thr, "fmap_score: null centroid/threshold table");
    OODB200_REQUIRE(!(metric_mask & (1 << OODB200_METRIC_COS)) || cent_unit, "fmap_score: cosine needs cent_unit");
    OODB200_REQUIRE(dist && argmin && decision, "fmap_score: null output pointer");
    FmapParams p = {};
    p.map_ptrs = map_ptrs; p.boxes = boxes; p.img_idx = img_idx; p.stride_idx = stride_idx;
    p.cls = cls; p.img_start = img_start; p.compat_q1 = flags & OODB200_FMAP_COMPAT_Q1 ? 1 : 0;
    p.group_score = flags & OODB200_FMAP_GROUP_SCORE ? 1 : 0; p.n = n; p.n_img = n_img;
    p.metric_mask = metric_mask; p.normalize = normalize;
    p.cent = cent; p.cent_unit = cent_unit ? cent_unit : cent; p.cent_off = cent_off; p.cent_k = cent_k; p.nc = nc; p.thr = thr;
    p.dist = dist; p.argmin = argmin; p.decision = decision; p.pooled_user = pooled; p.pooled_user_ld = pooled_ld;
    p.nhwc = nhwc;
    return launch_fmap(p, map_chw, scale, cls_used_out, out_index_out, workspace, workspace_bytes, stream, "fmap_score");
}

#define OODB200_FMAP_SCORE_ARGS                                                                                       \
    const float *const *map_ptrs, const int32_t *map_chw, const float *scale, int n_img, const float *boxes,          \
        const int32_t *img_idx, const int32_t *stride_idx, const int32_t *cls, const int32_t *img_start, int flags, \
        int n, int metric_mask, int normalize, const float *cent, const float *cent_unit, const int64_t *cent_off,    \
        const int32_t *cent_k, int nc, const double *thr, float *dist, int32_t *argmin, uint8_t *decision,            \
        float *pooled, int pooled_ld, int32_t *cls_used_out, int32_t *out_index_out, void *workspace,                 \
        int64_t workspace_bytes, void *stream
#define OODB200_FMAP_SCORE_PASS                                                                                        \
    map_ptrs, map_chw, scale, n_img, boxes, img_idx, stride_idx, cls, img_start, flags, n, metric_mask, normalize, \
        cent, cent_unit, cent_off, cent_k, nc, thr, dist, argmin, decision, pooled, pooled_ld, cls_used_out,           \
        out_index_out, workspace, workspace_bytes, stream

extern "C" int oodb200_fmap_score_f32(OODB200_FMAP_SCORE_ARGS) { return fmap_score_impl(0, OODB200_FMAP_SCORE_PASS); }
extern "C" int oodb200_fmap_score_nhwc_f32(OODB200_FMAP_SCORE_ARGS) { return fmap_score_impl(1, OODB200_FMAP_SCORE_PASS); }

extern "C" int oodb200_q1_plan_i32(const int32_t* img_start, const int32_t* stride_idx, const int32_t* cls, int n_img,
                                   int32_t* cls_used, int32_t* out_index, void* stream) {
    OODB200_REQUIRE(n_img >= 0, "q1_plan: negative n_img");
    if (n_img == 0) return OODB200_OK;
    OODB200_REQUIRE(img_start && stride_idx && cls && cls_used && out_index, "q1_plan: null pointer");
    q1_plan_kernel<<<n_img, 128, 0, (cudaStream_t)stream>>>(img_start, stride_idx, cls, cls_used, out_index);
    return check_launch("q1_plan");
}
