// K1 (RoIAlign 1x1 adaptive pooling) and K1+K2 (pool -> normalise -> distance-min-threshold), sm_100a.
//
// Replaces /root/reference/ultralytics/models/yolo/detect/predict.py:13-90 (torchvision roi_align
// with output 1x1, adaptive sampling grid, aligned=False) and the per-box loop of
// /root/reference/ood_utils.py:2038-2180 (+ :2404-2409 normalize, :2422-2430 pairwise distance + min).
//
// Two launches per batch:
//  plan_kernel  one small CTA per image: quirk-Q1 class / output slot of every box, and three per-stride box
//               lists so that the main grid runs the heaviest boxes (largest stride = most channels) first and
//               keeps boxes of one image adjacent (overlapping windows then hit in L2).
//  fmap_kernel  one CTA (8 warps) per detection.
//   1. The RoIAlign sample grid factorises: sum_{iy,ix} bilinear(y_iy, x_ix) = sum_r sum_c wy[r] wx[c] v[r,c],
//      because both the bilinear weights and the "sample outside [-1,H]x[-1,W] contributes 0" mask are
//      products of a y-term and an x-term.  Each CTA builds wy[], wx[] (sample coordinates are evaluated
//      with the exact float32 operation order of the reference kernel, no FMA contraction), so every
//      feature-map element of the window is read ONCE instead of ~4 times.
//   2. Window gather: lanes run over the flattened window (x fastest -> consecutive lanes read consecutive
//      addresses of an NCHW row), warps run over groups of CU channels.  Loads of the next channel group are
//      issued before the current group is reduced (double buffer in registers); channel offsets are immediates
//      of the load for the usual map sizes; one butterfly reduction per CU channels instead of CU full
//      warp reductions.  NCHW rows of a box are short (8..52 B), so the unit of DRAM traffic is the 32-byte
//      sector; see DESIGN.md for the roofline accounting.
//   3. Pooled vector stays in shared memory: L2 norm, then every centroid of (class, stride) is streamed
//      once (128-bit loads, L2-resident table), L1 / L2 / cosine evaluated in the same sweep, first-minimum
//      arg-min, threshold compare in float64.
#include "common.cuh"

#include <float.h>
#include <limits.h>

namespace oodb200 {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
#ifndef OODB200_FMAP_DOUBLE_BUFFER
#define OODB200_FMAP_DOUBLE_BUFFER 0
#endif
#ifndef OODB200_FMAP_MIN_BLOCKS
#define OODB200_FMAP_MIN_BLOCKS 6
#endif

struct FmapParams {
    const float* const* map_ptrs;
    int C[3], H[3], W[3];
    float scale[3];
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
    float* pooled;
    int pooled_ld;
    // plan (workspace or caller-provided)
    int* counts;                // [4] boxes per stride (device counters, zeroed by a memset)
    int* lists;                 // [3][n] box ids per stride, image-major
    int32_t* cls_used;          // [n]
    int32_t* out_index;         // [n]
    int ext_pad;                // smem floats reserved for each of wy / wx
    int c_pad;                  // smem floats reserved for each of xs / xu
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
        int v = (int)floorf(f) - 2;
        lo = v < 0 ? 0 : v;
    }
    if (row < extent - 1) {
        float f = ((float)(row + 1) - start) * inv - 0.5f;
        if (f < -4.f) f = -4.f;
        if (f > (float)grid) f = (float)grid;
        int v = (int)ceilf(f) + 2;
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

// Reduce N per-lane partial sums to N channel totals; afterwards every lane of group (lane >> (5 - log2 N))
// holds the total of channel (lane >> (5 - log2 N)).
template <int N>
__device__ __forceinline__ float butterfly(float (&v)[N], int lane) {
    int o = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, o >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, o);
        }
    }
    float r = v[0];
    for (; o > 0; o >>= 1) r += __shfl_xor_sync(kFull, r, o);
    return r;
}

template <int N> struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// Accumulate sum_p w_p * v[c, p] for every channel into acc_s[c].  R window positions per lane (chunks of 32*R),
// CU channels per warp step, HWC = H*W when known at compile time (0: run-time stride).
template <int R, int CU, int HWC>
__device__ __forceinline__ void pool_window(const float* __restrict__ img, int C, int hw_rt, int W, int y0, int x0,
                                            int wh, int ww, const float* wy, const float* wx, float* acc_s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int HW = HWC ? HWC : hw_rt;
    const int P = wh * ww;
    const float inv_ww = 1.0f / (float)ww;
    constexpr int kStep = kWarps * CU;
    for (int base = 0; base < P; base += 32 * R) {
        int off[R];
        float wt[R];
#pragma unroll
        for (int t = 0; t < R; ++t) {
            const int q = base + lane + 32 * t;
            if (q < P) {
                int y = (int)(((float)q + 0.5f) * inv_ww);
                if (y * ww > q) --y;
                else if ((y + 1) * ww <= q) ++y;
                const int x = q - y * ww;
                off[t] = (y0 + y) * W + (x0 + x);
                wt[t] = wy[y] * wx[x];
            } else {
                off[t] = y0 * W + x0;
                wt[t] = 0.f;
            }
        }
        auto load = [&](float (&v)[CU][R], int c0) {
            const float* __restrict__ b = img + (size_t)c0 * HW;
            if (c0 + CU <= C) {
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    const float* __restrict__ pt = b + off[t];
#pragma unroll
                    for (int u = 0; u < CU; ++u) v[u][t] = ldg_f32(pt + u * HW);
                }
            } else {                                     // C % CU != 0: clamp (results of the extra channels are dropped)
#pragma unroll
                for (int t = 0; t < R; ++t)
#pragma unroll
                    for (int u = 0; u < CU; ++u) v[u][t] = ldg_f32(b + min(u, C - 1 - c0) * HW + off[t]);
            }
        };
        auto consume = [&](float (&v)[CU][R], int c0) {
            float a[CU];
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                float s = 0.f;
#pragma unroll
                for (int t = 0; t < R; ++t) s = fmaf(wt[t], v[u][t], s);
                a[u] = s;
            }
            const float tot = butterfly<CU>(a, lane);
            const int c = c0 + (lane >> (5 - Log2<CU>::v));
            if ((lane & (32 / CU - 1)) == 0 && c < C) acc_s[c] = (base == 0) ? tot : acc_s[c] + tot;
        };
#if OODB200_FMAP_DOUBLE_BUFFER
        float va[CU][R], vb[CU][R];
        int c0 = warp * CU;
        if (c0 < C) load(va, c0);
        while (c0 < C) {                                 // ping-pong: next group's loads fly under this group's math
            int cn = c0 + kStep;
            if (cn < C) load(vb, cn);
            consume(va, c0);
            c0 = cn;
            if (c0 >= C) break;
            cn = c0 + kStep;
            if (cn < C) load(va, cn);
            consume(vb, c0);
            c0 = cn;
        }
#else
        for (int c0 = warp * CU; c0 < C; c0 += kStep) {
            float va[CU][R];
            load(va, c0);
            consume(va, c0);
        }
#endif
    }
}

template <int HWC>
__device__ __forceinline__ void pool_dispatch(const float* img, int C, int hw, int W, int y0, int x0, int wh, int ww,
                                              const float* wy, const float* wx, float* acc_s) {
    const int P = wh * ww;
    if (P <= 32) pool_window<1, 8, HWC>(img, C, hw, W, y0, x0, wh, ww, wy, wx, acc_s);
    else if (P <= 64) pool_window<2, 4, HWC>(img, C, hw, W, y0, x0, wh, ww, wy, wx, acc_s);
    else if (P <= 128) pool_window<4, 2, HWC>(img, C, hw, W, y0, x0, wh, ww, wy, wx, acc_s);
    else pool_window<8, 2, HWC>(img, C, hw, W, y0, x0, wh, ww, wy, wx, acc_s);
}

__global__ void __launch_bounds__(kThreads, OODB200_FMAP_MIN_BLOCKS) fmap_kernel(const FmapParams p) {
    extern __shared__ __align__(16) float smem[];
    float* wy = smem;
    float* wx = wy + p.ext_pad;
    float* xs = wx + p.ext_pad;          // pooled -> normalised vector
    float* xu = xs + p.c_pad;            // unit vector for cosine
    __shared__ float s_red[kWarps];
    __shared__ int s_win[4];
    __shared__ float s_wmin[OODB200_N_METRICS][kWarps];
    __shared__ int s_warg[OODB200_N_METRICS][kWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // heaviest stride first (largest-processing-time-first keeps the tail short)
    int i = blockIdx.x, box;
    const int n2 = p.counts[2], n1 = p.counts[1], n0 = p.counts[0];
    if (i < n2) box = p.lists[2 * p.n + i];
    else if ((i -= n2) < n1) box = p.lists[p.n + i];
    else if ((i -= n1) < n0) box = p.lists[i];
    else return;                                      // boxes with an invalid stride were answered by the plan kernel
    const int s = p.stride_idx[box];
    const int img = p.img_idx[box];
    const int out = p.out_index[box];
    const int C = p.C[s], H = p.H[s], W = p.W[s];

    // ---- ROI geometry (predict.py:64-70 -> roi_align, aligned=False) ----
    const float sc = p.scale[s];
    const float4 bx = *reinterpret_cast<const float4*>(p.boxes + 4 * (size_t)box);
    const float sw = __fmul_rn(bx.x, sc), sh = __fmul_rn(bx.y, sc);
    const float ew = __fmul_rn(bx.z, sc), eh = __fmul_rn(bx.w, sc);
    const float rw = fmaxf(__fsub_rn(ew, sw), 1.0f), rh = fmaxf(__fsub_rn(eh, sh), 1.0f);
    const int gw = (int)ceilf(rw), gh = (int)ceilf(rh);
    const float count = (float)max(gh * gw, 1);

    if (tid == 0) {
        s_win[0] = INT_MAX; s_win[1] = -1; s_win[2] = INT_MAX; s_win[3] = -1;
    }
    __syncthreads();
    for (int k = tid; k < gh + gw; k += kThreads) {
        const bool isy = k < gh;
        const AxisSample a = isy ? axis_sample(sh, rh, gh, H, k) : axis_sample(sw, rw, gw, W, k - gh);
        if (a.valid) {
            atomicMin(&s_win[isy ? 0 : 2], a.low);
            atomicMax(&s_win[isy ? 1 : 3], a.high);
        }
    }
    __syncthreads();
    const int y0 = s_win[0], x0 = s_win[2];
    const int wh = s_win[1] - y0 + 1, ww = s_win[3] - x0 + 1;
    const bool empty = (s_win[1] < 0) || (s_win[3] < 0);

    if (!empty) {
        for (int r = tid; r < wh + ww; r += kThreads) {
            if (r < wh) wy[r] = axis_weight(sh, rh, gh, H, y0 + r);
            else wx[r - wh] = axis_weight(sw, rw, gw, W, x0 + (r - wh));
        }
    }
    __syncthreads();

    // ---- gather + pool ----
    if (empty) {
        for (int c = tid; c < C; c += kThreads) xs[c] = 0.f;
    } else {
        const float* img_base = p.map_ptrs[img * 3 + s];
        const int HW = H * W;
        switch (HW) {                                 // map sizes of 320/640/1280-pixel inputs: immediate channel offsets
            case 400: pool_dispatch<400>(img_base, C, HW, W, y0, x0, wh, ww, wy, wx, xs); break;
            case 1600: pool_dispatch<1600>(img_base, C, HW, W, y0, x0, wh, ww, wy, wx, xs); break;
            case 6400: pool_dispatch<6400>(img_base, C, HW, W, y0, x0, wh, ww, wy, wx, xs); break;
            case 25600: pool_dispatch<25600>(img_base, C, HW, W, y0, x0, wh, ww, wy, wx, xs); break;
            default: pool_dispatch<0>(img_base, C, HW, W, y0, x0, wh, ww, wy, wx, xs); break;
        }
    }
    __syncthreads();
    // ---- average (roi_align.py:192-196) and, for K2, the squared norm in the same sweep ----
    float ss = 0.f;
    for (int c = tid; c < C; c += kThreads) {
        const float v = empty ? 0.f : __fdiv_rn(xs[c], count);
        xs[c] = v;
        ss = fmaf(v, v, ss);
        if (p.pooled) p.pooled[(size_t)out * p.pooled_ld + c] = v;
    }
    if (p.cent == nullptr) return;                    // K1 only

    // ---- K2: normalise (ood_utils.py:2409 -> sklearn normalize) ----
    if (p.normalize) {
        ss = block_sum<kWarps>(ss, s_red);
        float nrm = sqrtf(ss);
        if (nrm < 10.f * FLT_EPSILON) nrm = 1.f;      // _handle_zeros_in_scale
        for (int c = tid; c < C; c += kThreads) xs[c] = __fdiv_rn(xs[c], nrm);
    }
    __syncthreads();
    const bool want_l1 = p.metric_mask & (1 << OODB200_METRIC_L1);
    const bool want_l2 = p.metric_mask & (1 << OODB200_METRIC_L2);
    const bool want_cos = p.metric_mask & (1 << OODB200_METRIC_COS);
    if (want_cos) {                                   // cosine_distances re-normalises X (pairwise.py:1171-1182)
        float s2 = 0.f;
        for (int c = tid; c < C; c += kThreads) s2 = fmaf(xs[c], xs[c], s2);
        s2 = block_sum<kWarps>(s2, s_red);
        float n2v = sqrtf(s2);
        if (n2v < 10.f * FLT_EPSILON) n2v = 1.f;
        for (int c = tid; c < C; c += kThreads) xu[c] = __fdiv_rn(xs[c], n2v);
        __syncthreads();
    }

    // ---- distances to the centroids of (class, stride); warps stride over centroids ----
    const int cls = p.cls_used[box];
    const bool cls_ok = cls >= 0 && cls < p.nc;
    const int K = cls_ok ? p.cent_k[s * p.nc + cls] : 0;
    float best[OODB200_N_METRICS] = {FLT_MAX, FLT_MAX, FLT_MAX};
    int barg[OODB200_N_METRICS] = {-1, -1, -1};
    if (K > 0) {
        const int64_t off = p.cent_off[s * p.nc + cls];
        const bool vec = (C % 4 == 0) && (off % 4 == 0);
        for (int k = warp; k < K; k += kWarps) {
            const float* __restrict__ ck = p.cent + off + (int64_t)k * C;
            const float* __restrict__ cu = want_cos ? p.cent_unit + off + (int64_t)k * C : nullptr;
            float a1 = 0.f, a2 = 0.f, ac = 0.f;
            if (vec) {
                for (int d = lane * 4; d < C; d += 128) {
                    const float4 x4 = *reinterpret_cast<const float4*>(xs + d);
                    if (want_l1 || want_l2) {
                        const float4 c4 = __ldg(reinterpret_cast<const float4*>(ck + d));
                        const float d0 = x4.x - c4.x, d1 = x4.y - c4.y, d2 = x4.z - c4.z, d3 = x4.w - c4.w;
                        a1 += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
                        a2 = fmaf(d0, d0, a2); a2 = fmaf(d1, d1, a2); a2 = fmaf(d2, d2, a2); a2 = fmaf(d3, d3, a2);
                    }
                    if (want_cos) {
                        const float4 u4 = *reinterpret_cast<const float4*>(xu + d);
                        const float4 c4 = __ldg(reinterpret_cast<const float4*>(cu + d));
                        ac = fmaf(u4.x, c4.x, ac); ac = fmaf(u4.y, c4.y, ac); ac = fmaf(u4.z, c4.z, ac); ac = fmaf(u4.w, c4.w, ac);
                    }
                }
            } else {
                for (int d = lane; d < C; d += 32) {
                    const float x = xs[d];
                    if (want_l1 || want_l2) {
                        const float df = x - __ldg(ck + d);
                        a1 += fabsf(df);
                        a2 = fmaf(df, df, a2);
                    }
                    if (want_cos) ac = fmaf(xu[d], __ldg(cu + d), ac);
                }
            }
            if (want_l1) {
                a1 = warp_sum(a1);
                if (a1 < best[0]) { best[0] = a1; barg[0] = k; }
            }
            if (want_l2) {
                a2 = sqrtf(fmaxf(warp_sum(a2), 0.f));
                if (a2 < best[1]) { best[1] = a2; barg[1] = k; }
            }
            if (want_cos) {
                ac = fminf(fmaxf(1.0f - warp_sum(ac), 0.f), 2.f);
                if (ac < best[2]) { best[2] = ac; barg[2] = k; }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int m = 0; m < OODB200_N_METRICS; ++m) { s_wmin[m][warp] = best[m]; s_warg[m][warp] = barg[m]; }
    }
    __syncthreads();
    if (tid < OODB200_N_METRICS && (p.metric_mask >> tid & 1)) {
        const int m = tid;
        float d = 1000.f;                             // ood_utils.py:2159-2164
        int a = -1;
        if (K > 0) {
            d = FLT_MAX;
            for (int w = 0; w < kWarps; ++w) {        // first minimum: smaller distance, then smaller index
                const int wa = s_warg[m][w];
                if (wa >= 0 && (s_wmin[m][w] < d || (s_wmin[m][w] == d && wa < a))) { d = s_wmin[m][w]; a = wa; }
            }
        }
        const double t = cls_ok ? p.thr[(size_t)m * 3 * p.nc + s * p.nc + cls] : nan("");
        const size_t o = (size_t)m * p.n + out;
        p.dist[o] = d;
        p.argmin[o] = a;
        p.decision[o] = (t == t && (double)d < t) ? 1 : 0;   // NaN threshold = "no threshold" -> OoD (:2173-2180)
    }
}

// One CTA per image: quirk Q1 (ood_utils.py:2152-2154) and the per-stride box lists (image-major inside a stride).
__global__ void __launch_bounds__(128) plan_kernel(const FmapParams p) {
    const int img = blockIdx.x;
    const int b0 = p.img_start[img], m = p.img_start[img + 1] - b0;
    __shared__ int s_cnt[4], s_base[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < m; b += blockDim.x) {
        const int s = p.stride_idx[b0 + b];
        atomicAdd(&s_cnt[(s >= 0 && s <= 2) ? s : 3], 1);
    }
    __syncthreads();
    if (threadIdx.x < 3) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&p.counts[threadIdx.x], s_cnt[threadIdx.x]) : 0;
    __syncthreads();
    for (int b = threadIdx.x; b < m; b += blockDim.x) {
        const int s = p.stride_idx[b0 + b];
        const bool ok = s >= 0 && s <= 2;
        int j = 0;                                    // rank among earlier boxes of the same kind (m <= 300)
        for (int e = 0; e < b; ++e) {
            const int se = p.stride_idx[b0 + e];
            j += ok ? (se == s) : !(se >= 0 && se <= 2);
        }
        int before = 0;
        for (int t = 0; t < (ok ? s : 3); ++t) before += s_cnt[t];
        int cls_u = p.cls ? p.cls[b0 + b] : 0, out = b0 + b;
        if (p.compat_q1) {
            cls_u = ok ? p.cls[b0 + j] : -1;
            out = b0 + before + j;
        }
        p.cls_used[b0 + b] = cls_u;
        p.out_index[b0 + b] = out;
        if (ok) {
            p.lists[s * p.n + s_base[s] + j] = b0 + b;
        } else if (p.cent) {                          // never pooled by the reference either: answered here
            for (int k = 0; k < OODB200_N_METRICS; ++k)
                if (p.metric_mask >> k & 1) {
                    const size_t o = (size_t)k * p.n + out;
                    p.dist[o] = nanf("");
                    p.argmin[o] = -1;
                    p.decision[o] = 0;
                }
        }
    }
}

// standalone Q1 plan (same arithmetic as plan_kernel, without the lists)
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

struct WorkspaceLayout {
    size_t counts, lists, cls_used, out_index, total;
};

static WorkspaceLayout layout_of(int n) {
    WorkspaceLayout L;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    L.counts = o; o = up(o + 16);
    L.lists = o; o = up(o + sizeof(int) * 3 * (size_t)n);
    L.cls_used = o; o = up(o + sizeof(int) * (size_t)n);
    L.out_index = o; o = up(o + sizeof(int) * (size_t)n);
    L.total = o;
    return L;
}

static int launch_fmap(FmapParams& p, const int32_t* map_chw, const float* scale, int32_t* cls_used_out,
                       int32_t* out_index_out, void* workspace, int64_t workspace_bytes, void* stream, const char* what) {
    int ext = 1, cmax = 1;
    for (int s = 0; s < 3; ++s) {
        p.C[s] = map_chw[3 * s];
        p.H[s] = map_chw[3 * s + 1];
        p.W[s] = map_chw[3 * s + 2];
        p.scale[s] = scale[s];
        OODB200_REQUIRE(p.C[s] > 0 && p.H[s] > 0 && p.W[s] > 0, "%s: map %d has non-positive shape", what, s);
        ext = max(ext, max(p.H[s], p.W[s]));
        cmax = max(cmax, p.C[s]);
    }
    p.ext_pad = (ext + 3) & ~3;
    p.c_pad = (cmax + 3) & ~3;
    const size_t smem = sizeof(float) * (2 * (size_t)p.ext_pad + 2 * (size_t)p.c_pad);
    OODB200_REQUIRE(smem <= 200 * 1024, "%s: maps too large for the shared-memory layout (%zu B)", what, smem);
    if (p.n == 0) return OODB200_OK;
    const WorkspaceLayout L = layout_of(p.n);
    OODB200_REQUIRE(workspace && workspace_bytes >= (int64_t)L.total, "%s: workspace too small (%lld < %zu bytes)", what,
                    (long long)workspace_bytes, L.total);
    OODB200_REQUIRE(((uintptr_t)workspace & 255) == 0, "%s: workspace must be 256-byte aligned", what);
    char* ws = (char*)workspace;
    p.counts = (int*)(ws + L.counts);
    p.lists = (int*)(ws + L.lists);
    p.cls_used = cls_used_out ? cls_used_out : (int32_t*)(ws + L.cls_used);
    p.out_index = out_index_out ? out_index_out : (int32_t*)(ws + L.out_index);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(p.counts, 0, 16, st);
    if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    plan_kernel<<<p.n_img, 128, 0, st>>>(p);
    int rc = check_launch(what);
    if (rc) return rc;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(fmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    }
    fmap_kernel<<<p.n, kThreads, smem, st>>>(p);
    return check_launch(what);
}

}  // namespace oodb200

using namespace oodb200;

extern "C" int64_t oodb200_fmap_workspace_bytes(int n, const int32_t* map_chw, int need_pooled) {
    (void)map_chw; (void)need_pooled;
    if (n <= 0) return 256;
    return (int64_t)layout_of(n).total;
}

extern "C" int oodb200_roi_pool_f32(const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                                    const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                                    const int32_t* img_start, int n,
                                    float* out, int out_ld, void* workspace, int64_t workspace_bytes, void* stream) {
    OODB200_REQUIRE(n >= 0 && n_img >= 0, "roi_pool: negative size");
    OODB200_REQUIRE(map_chw && scale, "roi_pool: map_chw/scale must be host arrays");
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(map_ptrs && boxes && img_idx && stride_idx && img_start && out, "roi_pool: null pointer");
    FmapParams p = {};
    p.map_ptrs = map_ptrs; p.boxes = boxes; p.img_idx = img_idx; p.stride_idx = stride_idx; p.img_start = img_start;
    p.n = n; p.n_img = n_img; p.pooled = out; p.pooled_ld = out_ld;
    return launch_fmap(p, map_chw, scale, nullptr, nullptr, workspace, workspace_bytes, stream, "roi_pool");
}

extern "C" int oodb200_fmap_score_f32(const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                                      const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                                      const int32_t* cls, const int32_t* img_start, int compat_q1, int n,
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
    p.cls = cls; p.img_start = img_start; p.compat_q1 = compat_q1; p.n = n; p.n_img = n_img;
    p.metric_mask = metric_mask; p.normalize = normalize;
    p.cent = cent; p.cent_unit = cent_unit; p.cent_off = cent_off; p.cent_k = cent_k; p.nc = nc; p.thr = thr;
    p.dist = dist; p.argmin = argmin; p.decision = decision; p.pooled = pooled; p.pooled_ld = pooled_ld;
    return launch_fmap(p, map_chw, scale, cls_used_out, out_index_out, workspace, workspace_bytes, stream, "fmap_score");
}

extern "C" int oodb200_q1_plan_i32(const int32_t* img_start, const int32_t* stride_idx, const int32_t* cls, int n_img,
                                   int32_t* cls_used, int32_t* out_index, void* stream) {
    OODB200_REQUIRE(n_img >= 0, "q1_plan: negative n_img");
    if (n_img == 0) return OODB200_OK;
    OODB200_REQUIRE(img_start && stride_idx && cls && cls_used && out_index, "q1_plan: null pointer");
    q1_plan_kernel<<<n_img, 128, 0, (cudaStream_t)stream>>>(img_start, stride_idx, cls, cls_used, out_index);
    return check_launch("q1_plan");
}
