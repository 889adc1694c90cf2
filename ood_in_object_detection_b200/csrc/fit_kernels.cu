// Fit-stage kernels, sm_100a:
//   K2 standalone  vec_score   -- normalise + distance-min of already pooled vectors, segmented by (class, stride)
//                                 (/root/reference/ood_utils.py:2000-2036, :2404-2430)
//   K5             radix_hist  -- one pass of an exact radix select over float32 score bit patterns
//                                 (np.percentile(scores, q, method='lower'), /root/reference/ood_utils.py:613,626)
// Both are bandwidth-bound streams over N rows / N scores; see DESIGN.md for the byte accounting.
#include "common.cuh"

#include <float.h>
#include <limits.h>
#include <stdlib.h>

namespace oodb200 {

constexpr int kVecThreads = 256;
constexpr int kVecWarps = kVecThreads / 32;

struct VecParams {
    const float* x;
    int64_t ld;
    int dim;
    const int64_t* seg_off;
    int n_seg;
    int64_t n_rows;
    int metric_mask, normalize;
    const float* cent;
    const float* cent_unit;
    const int64_t* cent_row_off;
    const int32_t* cent_k;
    float* dist;
    int32_t* argmin;
    int d_pad;
    const double* thr;        // optional [3][n_seg], NaN = no threshold
    uint8_t* decision;        // optional [3][n_rows]
};

__device__ __forceinline__ int find_segment(const int64_t* __restrict__ seg_off, int n_seg, int64_t row) {
    int lo = 0, hi = n_seg;                       // seg_off[lo] <= row < seg_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(seg_off + mid) <= row) lo = mid; else hi = mid;
    }
    return lo;
}

// one warp per row; the row lives in shared memory (xs: normalised, xu: unit vector for cosine)
__global__ void __launch_bounds__(kVecThreads) vec_score_kernel(const VecParams p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* xs = smem + (size_t)warp * 2 * p.d_pad;
    float* xu = xs + p.d_pad;
    const int D = p.dim;
    const bool want_l1 = p.metric_mask & (1 << OODB200_METRIC_L1);
    const bool want_l2 = p.metric_mask & (1 << OODB200_METRIC_L2);
    const bool want_cos = p.metric_mask & (1 << OODB200_METRIC_COS);
    for (int64_t row = (int64_t)blockIdx.x * kVecWarps + warp; row < p.n_rows; row += (int64_t)gridDim.x * kVecWarps) {
        const float* __restrict__ xr = p.x + row * p.ld;
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float v = __ldg(xr + d);
            xs[d] = v;
            ss = fmaf(v, v, ss);
        }
        if (p.normalize) {                         // sklearn normalize (ood_utils.py:2409)
            float nrm = sqrtf(warp_sum(ss));
            if (nrm < 10.f * FLT_EPSILON) nrm = 1.f;
            for (int d = lane; d < D; d += 32) xs[d] = __fdiv_rn(xs[d], nrm);
        }
        __syncwarp();
        if (want_cos) {
            float s2 = 0.f;
            for (int d = lane; d < D; d += 32) s2 = fmaf(xs[d], xs[d], s2);
            float n2 = sqrtf(warp_sum(s2));
            if (n2 < 10.f * FLT_EPSILON) n2 = 1.f;
            for (int d = lane; d < D; d += 32) xu[d] = __fdiv_rn(xs[d], n2);
            __syncwarp();
        }
        const int g = find_segment(p.seg_off, p.n_seg, row);
        const int K = p.cent_k[g];
        const int64_t off = p.cent_row_off[g] * D;
        float best[OODB200_N_METRICS] = {FLT_MAX, FLT_MAX, FLT_MAX};
        int barg[OODB200_N_METRICS] = {-1, -1, -1};
        const bool vec = (D % 4 == 0) && (off % 4 == 0);
        for (int k = 0; k < K; ++k) {
            const float* __restrict__ ck = p.cent + off + (int64_t)k * D;
            const float* __restrict__ cu = want_cos ? p.cent_unit + off + (int64_t)k * D : nullptr;
            float a1 = 0.f, a2 = 0.f, ac = 0.f;
            if (vec) {
                for (int d = lane * 4; d < D; d += 128) {
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
                for (int d = lane; d < D; d += 32) {
                    if (want_l1 || want_l2) {
                        const float df = xs[d] - __ldg(ck + d);
                        a1 += fabsf(df);
                        a2 = fmaf(df, df, a2);
                    }
                    if (want_cos) ac = fmaf(xu[d], __ldg(cu + d), ac);
                }
            }
            if (want_l1) { a1 = warp_sum(a1); if (a1 < best[0]) { best[0] = a1; barg[0] = k; } }
            if (want_l2) { a2 = sqrtf(fmaxf(warp_sum(a2), 0.f)); if (a2 < best[1]) { best[1] = a2; barg[1] = k; } }
            if (want_cos) { ac = fminf(fmaxf(1.0f - warp_sum(ac), 0.f), 2.f); if (ac < best[2]) { best[2] = ac; barg[2] = k; } }
        }
        if (lane < OODB200_N_METRICS && (p.metric_mask >> lane & 1)) {
            const size_t o = (size_t)lane * p.n_rows + row;
            const float dv = K > 0 ? best[lane] : 1000.f;          // no cluster: ood_utils.py:2159-2164
            p.dist[o] = dv;
            p.argmin[o] = K > 0 ? barg[lane] : -1;
            if (p.decision) {                                        // :2173-2180 (NaN = falsy threshold -> OoD)
                const double t = p.thr[(size_t)lane * p.n_seg + g];
                p.decision[o] = (t == t && (double)dv < t) ? 1 : 0;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- fast path (one metric)
// vec_score_fast_kernel<NJ, METRIC>: D % 4 == 0, D <= 128 * NJ, 16-byte aligned rows.  A CTA owns a contiguous range of
// rows; per segment inside that range the segment's centroid rows (the unit-norm twins for cosine) are staged ONCE into
// shared memory (chunks of RC rows when K rows do not fit), then every warp takes groups of 4 consecutive rows: the 4
// rows live in registers (NJ float4 per lane and row), a centroid row is read once from shared memory (128-bit,
// conflict-free) for all 4 rows, and the 4 x 8 partial sums of 8 centroid rows are reduced with ONE transposing
// butterfly (31 shuffles for 32 totals instead of 160).  The per-lane accumulation order and the pairing tree of the
// reduction are those of score_box_smem (csrc/fmap_score.cu): given the same pooled vector, fit-time and decision-time
// distances are bit-identical (thresholds are compared with distances of the same arithmetic).
#ifndef OODB200_VF_THREADS_WIDE
#define OODB200_VF_THREADS_WIDE 384
#endif
constexpr int kVfRows = 4;                          // rows per warp and group
constexpr int kVfK = 8;                             // centroid rows per butterfly
__host__ __device__ constexpr int vf_threads(int nj) { return nj <= 2 ? 256 : OODB200_VF_THREADS_WIDE; }   // wide rows: 3 warps per scheduler under a 168-register cap

__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
    // afterwards lane j holds the warp-wide total of value index j; the pairing tree is the xor butterfly 16, 8, 4, 2, 1
    int o = 16;
#pragma unroll
    for (int n = 32; n > 1; n >>= 1, o >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// mbarrier / bulk-copy primitives of the row staging below
__device__ __forceinline__ uint32_t vf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vf_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vf_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void vf_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vf_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vf_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(vf_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(vf_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void vf_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(vf_smem_u32(bar)), "r"(parity)
                     : "memory");
    }
}

template <int NJ, int METRIC>
__global__ void __launch_bounds__(vf_threads(NJ), NJ <= 2 ? 2 : 1) vec_score_fast_kernel(const VecParams p, int rows_per_cta, int smem_floats) {
    constexpr int kVecThreads = vf_threads(NJ), kVecWarps = kVecThreads / 32;   // this kernel's own CTA shape
    extern __shared__ __align__(16) float s_cent[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.dim;
    const float* __restrict__ table = METRIC == OODB200_METRIC_COS ? p.cent_unit : p.cent;
    const int64_t cta0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t cta1 = cta0 + rows_per_cta < p.n_rows ? cta0 + rows_per_cta : p.n_rows;
    const bool last_ok = (NJ - 1) * 128 + lane * 4 < D;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // Row staging: every warp owns a buffer of 4 rows at the END of the shared memory and one mbarrier.  The bulk copy of the
    // NEXT group of rows is issued as soon as the current group sits in registers, so it lands under the current group's
    // distance loop (measured at D = 576, K = 16, 12 warps: 3.4 ms per 2 M rows staged, 4.1 ms with plain loads).  A segment
    // whose centroid table does not fit beside the buffers gives them up (plain loads, the table may use all of it) rather
    // than sweeping the rows once per table chunk.
    const int rows_floats = kVecWarps * kVfRows * D;
    float* __restrict__ wbuf = s_cent + (smem_floats - rows_floats) + (size_t)warp * kVfRows * D;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_cent + smem_floats) + warp;
    if (lane == 0) {
        vf_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    const uint32_t row_bytes = (uint32_t)D * 4u;
    auto stage_rows = [&](int64_t gr, int64_t r1) {                  // lane 0: rows gr .. gr+3 (clamped to the segment) -> wbuf
        vf_mbar_expect_tx(bar, kVfRows * row_bytes);
        if (p.ld == D && gr + kVfRows <= r1) {
            vf_bulk_g2s(wbuf, p.x + gr * p.ld, kVfRows * row_bytes, bar);
        } else {
#pragma unroll
            for (int r = 0; r < kVfRows; ++r) {
                const int64_t row = gr + r < r1 ? gr + r : r1 - 1;   // short group: repeat the last row, result dropped
                vf_bulk_g2s(wbuf + (size_t)r * D, p.x + row * p.ld, row_bytes, bar);
            }
        }
    };
    int64_t r0 = cta0;
    while (r0 < cta1) {                                              // block-uniform loop over the segments of the range
        const int g = find_segment(p.seg_off, p.n_seg, r0);
        const int64_t seg_end = __ldg(p.seg_off + g + 1);
        const int64_t r1 = seg_end < cta1 ? seg_end : cta1;
        const int K = p.cent_k[g];
        const int64_t coff = p.cent_row_off[g] * (int64_t)D;
        const bool staged = (int64_t)K * D + rows_floats <= smem_floats && rows_floats < smem_floats;
        const int rc_max = (smem_floats - (staged ? rows_floats : 0)) / D;
        const int RC = K < rc_max ? K : rc_max;
        for (int kb = 0; kb == 0 || kb < K; kb += (RC > 0 ? RC : 1)) {
            const int rc = K - kb < RC ? K - kb : RC;
            const bool first = kb == 0, final = kb + rc >= K;
            __syncthreads();                                         // previous table no longer read
            for (int i = threadIdx.x * 4; i < rc * D; i += kVecThreads * 4)
                *reinterpret_cast<float4*>(s_cent + i) = __ldg(reinterpret_cast<const float4*>(table + coff + (int64_t)kb * D + i));
            __syncthreads();
            if (staged && r0 + (int64_t)warp * kVfRows < r1 && lane == 0) stage_rows(r0 + (int64_t)warp * kVfRows, r1);
            for (int64_t gr = r0 + (int64_t)warp * kVfRows; gr < r1; gr += (int64_t)kVecWarps * kVfRows) {
                float4 x[kVfRows][NJ];
                float n2v[kVfRows];
                if (staged) {                                        // block-uniform
                    vf_mbar_wait(bar, phase);
                    phase ^= 1u;
#pragma unroll
                    for (int r = 0; r < kVfRows; ++r) {
                        const float* __restrict__ xr = wbuf + (size_t)r * D + lane * 4;
#pragma unroll
                        for (int t = 0; t < NJ; ++t)
                            x[r][t] = (t < NJ - 1 || last_ok) ? *reinterpret_cast<const float4*>(xr + t * 128) : zero;
                    }
                    __syncwarp();                                    // every lane holds its part: the buffer is free again
                    if (gr + (int64_t)kVecWarps * kVfRows < r1 && lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        stage_rows(gr + (int64_t)kVecWarps * kVfRows, r1);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < kVfRows; ++r) {
                        const int64_t row = gr + r < r1 ? gr + r : r1 - 1;    // short group: repeat the last row, result dropped
                        const float* __restrict__ xr = p.x + row * p.ld + lane * 4;
#pragma unroll
                        for (int t = 0; t < NJ; ++t)
                            x[r][t] = (t < NJ - 1 || last_ok) ? __ldg(reinterpret_cast<const float4*>(xr + t * 128)) : zero;
                    }
                }
#pragma unroll
                for (int r = 0; r < kVfRows; ++r) {
                    float ss = 0.f;
#pragma unroll
                    for (int t = 0; t < NJ; ++t) {
                        ss = fmaf(x[r][t].x, x[r][t].x, ss); ss = fmaf(x[r][t].y, x[r][t].y, ss);
                        ss = fmaf(x[r][t].z, x[r][t].z, ss); ss = fmaf(x[r][t].w, x[r][t].w, ss);
                    }
                    if (p.normalize) {                               // ood_utils.py:2409 -> sklearn normalize
                        float nrm = sqrtf(warp_sum(ss));
                        if (nrm < 10.f * FLT_EPSILON) nrm = 1.f;     // _handle_zeros_in_scale
                        ss = 0.f;
#pragma unroll
                        for (int t = 0; t < NJ; ++t) {
                            x[r][t].x = __fdiv_rn(x[r][t].x, nrm); x[r][t].y = __fdiv_rn(x[r][t].y, nrm);
                            x[r][t].z = __fdiv_rn(x[r][t].z, nrm); x[r][t].w = __fdiv_rn(x[r][t].w, nrm);
                            if (METRIC == OODB200_METRIC_COS) {
                                ss = fmaf(x[r][t].x, x[r][t].x, ss); ss = fmaf(x[r][t].y, x[r][t].y, ss);
                                ss = fmaf(x[r][t].z, x[r][t].z, ss); ss = fmaf(x[r][t].w, x[r][t].w, ss);
                            }
                        }
                    }
                    n2v[r] = 1.f;
                    if (METRIC == OODB200_METRIC_COS) {              // cosine_distances re-normalises X (pairwise.py:1171-1182)
                        n2v[r] = sqrtf(warp_sum(ss));
                        if (n2v[r] < 10.f * FLT_EPSILON) n2v[r] = 1.f;
                    }
                }
                // lane j = r + 4 * kk receives the total of (row r, centroid row k0 + kk)
                const int my_r = lane & 3, my_kk = lane >> 2;
                const float my_n2v = my_r == 0 ? n2v[0] : (my_r == 1 ? n2v[1] : (my_r == 2 ? n2v[2] : n2v[3]));
                float best = FLT_MAX;
                int barg = INT_MAX;
                for (int k0 = 0; k0 < rc; k0 += kVfK) {
                    float acc[32];
#pragma unroll
                    for (int kk = 0; kk < kVfK; ++kk) {
                        const int k = k0 + kk < rc ? k0 + kk : rc - 1;        // short batch: repeat the last row, result dropped
                        const float* __restrict__ ck = s_cent + (size_t)k * D + lane * 4;
                        float a[kVfRows] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int t = 0; t < NJ; ++t) {
                            const float4 c = (t < NJ - 1 || last_ok) ? *reinterpret_cast<const float4*>(ck + t * 128) : zero;
#pragma unroll
                            for (int r = 0; r < kVfRows; ++r) {
                                if (METRIC == OODB200_METRIC_COS) {
                                    a[r] = fmaf(x[r][t].x, c.x, a[r]); a[r] = fmaf(x[r][t].y, c.y, a[r]);
                                    a[r] = fmaf(x[r][t].z, c.z, a[r]); a[r] = fmaf(x[r][t].w, c.w, a[r]);
                                } else {
                                    const float e0 = x[r][t].x - c.x, e1 = x[r][t].y - c.y, e2 = x[r][t].z - c.z, e3 = x[r][t].w - c.w;
                                    if (METRIC == OODB200_METRIC_L1) a[r] += (fabsf(e0) + fabsf(e1)) + (fabsf(e2) + fabsf(e3));
                                    else { a[r] = fmaf(e0, e0, a[r]); a[r] = fmaf(e1, e1, a[r]); a[r] = fmaf(e2, e2, a[r]); a[r] = fmaf(e3, e3, a[r]); }
                                }
                            }
                        }
#pragma unroll
                        for (int r = 0; r < kVfRows; ++r) acc[r + 4 * kk] = a[r];
                    }
                    const float tot = transpose_reduce32(acc, lane);
                    if (k0 + my_kk < rc) {
                        float v = tot;
                        if (METRIC == OODB200_METRIC_L2) v = sqrtf(fmaxf(tot, 0.f));
                        if (METRIC == OODB200_METRIC_COS) v = fminf(fmaxf(1.0f - __fdiv_rn(tot, my_n2v), 0.f), 2.f);
                        if (v < best) { best = v; barg = kb + k0 + my_kk; }                    // increasing k per lane, strict '<'
                    }
                }
                // first minimum over the 8 lanes of a row: smaller value, then smaller index
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oa = __shfl_xor_sync(0xffffffffu, barg, o);
                    if (od < best || (od == best && oa < barg)) { best = od; barg = oa; }
                }
                if (lane < kVfRows && gr + lane < r1) {
                    const size_t o = (size_t)METRIC * p.n_rows + (size_t)(gr + lane);
                    if (!first) {                                    // merge with the earlier chunks (kept un-finalised in the outputs)
                        const float pd = p.dist[o];
                        const int pa = p.argmin[o];
                        if (pd <= best) { best = pd; barg = pa; }    // earlier rows win ties
                    }
                    if (final) {
                        const float dv = K > 0 ? best : 1000.f;   // no cluster: ood_utils.py:2159-2164
                        p.dist[o] = dv;
                        p.argmin[o] = K > 0 ? barg : -1;
                        if (p.decision) {                            // :2173-2180 (NaN = falsy threshold -> OoD)
                            const double t = p.thr[(size_t)METRIC * p.n_seg + g];
                            p.decision[o] = (t == t && (double)dv < t) ? 1 : 0;
                        }
                    } else {
                        p.dist[o] = best;
                        p.argmin[o] = barg;
                    }
                }
            }
            if (K == 0) break;
        }
        r0 = r1;
    }
}

typedef void (*VecFastKernel)(const VecParams, int, int);
template <int METRIC>
static VecFastKernel vec_fast_for(int nj) {
    switch (nj) {
        case 1: return vec_score_fast_kernel<1, METRIC>;
        case 2: return vec_score_fast_kernel<2, METRIC>;
        case 3: return vec_score_fast_kernel<3, METRIC>;
        case 4: return vec_score_fast_kernel<4, METRIC>;
        default: return vec_score_fast_kernel<5, METRIC>;
    }
}

// sklearn.preprocessing.normalize(x, axis=1) for float32 rows: one warp per row
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ x, int64_t ld, int dim, int64_t n_rows,
                                                             float* __restrict__ out, int64_t out_ld) {
    const int lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += (int64_t)gridDim.x * 8) {
        const float* __restrict__ xr = x + row * ld;
        float ss = 0.f;
        for (int d = lane; d < dim; d += 32) { const float v = __ldg(xr + d); ss = fmaf(v, v, ss); }
        float nrm = sqrtf(warp_sum(ss));
        if (nrm < 10.f * FLT_EPSILON) nrm = 1.f;
        for (int d = lane; d < dim; d += 32) out[row * out_ld + d] = __fdiv_rn(__ldg(xr + d), nrm);
    }
}

// DistanceMethod.compute_indness as the reference INTENDS it (ood_utils.py:1599-1604; the shipped code always
// returns -1, SURVEY.md Q2 -- the host keeps that behaviour in compat mode and never calls this kernel then).
// Piecewise linear through (min_dist, +1), (thr, 0), (max_dist, -1), python-float arithmetic.
__global__ void dist_indness_kernel(const float* __restrict__ dist, const int32_t* __restrict__ slot, int64_t n,
                                    const double* __restrict__ thr, const double* __restrict__ dmin,
                                    const double* __restrict__ dmax, int clip, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = slot[i];
    float r = -1.f;
    if (g >= 0) {
        const double t = thr[g], s = (double)dist[i];
        if (t == t) {
            double a = 0.0, b = 0.0;
            if (s > t) { a = -1.0 / (dmax[g] - t); b = t / (dmax[g] - t); }
            else if (s < t) { a = 1.0 / (dmin[g] - t); b = -t / (dmin[g] - t); }
            double v = a * s + b;
            if (clip) v = fmax(-1.0, fmin(v, 1.0));
            r = (float)v;
        }
    }
    out[i] = r;
}

// order-preserving map float32 -> uint32 (negative floats reversed, positives above them)
__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// one CTA-range per segment slice; shared-memory histogram of (1<<bits) bins, flushed with one atomic per bin
__global__ void __launch_bounds__(256) radix_hist_kernel(const float* __restrict__ scores, const int64_t* __restrict__ seg_off,
                                                         int n_seg, const uint32_t* __restrict__ prefix, int shift, int bits,
                                                         uint32_t* __restrict__ hist, uint32_t* __restrict__ minmax) {
    extern __shared__ uint32_t s_hist[];
    const int g = blockIdx.y;
    const int nb = 1 << bits;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t r0 = seg_off[g], r1 = seg_off[g + 1];
    const bool top = shift + bits >= 32;
    const uint32_t pfx = top ? 0u : prefix[g];
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
    for (int64_t r = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t key = float_key(scores[r]);
        kmin = min(kmin, key);
        kmax = max(kmax, key);
        if (top || (key >> (shift + bits)) == pfx) atomicAdd(&s_hist[(key >> shift) & (nb - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[(size_t)g * nb + i], s_hist[i]);
    if (minmax) {                                   // integer min / max of the keys: exact and order-independent
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        if ((threadIdx.x & 31) == 0 && kmin <= kmax) {
            atomicMin(&minmax[2 * g], kmin);
            atomicMax(&minmax[2 * g + 1], kmax);
        }
    }
}

}  // namespace oodb200

using namespace oodb200;

extern "C" int oodb200_vec_score_f32(const float* x, int64_t ld, int dim, const int64_t* seg_off, int n_seg, int64_t n_rows,
                                     int metric_mask, int normalize,
                                     const float* cent, const float* cent_unit, const int64_t* cent_row_off,
                                     const int32_t* cent_k, float* dist, int32_t* argmin, const double* thr,
                                     uint8_t* decision, void* stream) {
    OODB200_REQUIRE(dim > 0 && n_seg >= 0 && n_rows >= 0 && ld >= dim, "vec_score: bad size");
    OODB200_REQUIRE(metric_mask > 0 && metric_mask < (1 << OODB200_N_METRICS), "vec_score: metric_mask %d", metric_mask);
    if (n_rows == 0 || n_seg == 0) return OODB200_OK;
    OODB200_REQUIRE(x && seg_off && cent && cent_row_off && cent_k && dist && argmin, "vec_score: null pointer");
    OODB200_REQUIRE(!(metric_mask & (1 << OODB200_METRIC_COS)) || cent_unit, "vec_score: cosine needs cent_unit");
    OODB200_REQUIRE(!decision || thr, "vec_score: decision output needs thresholds");
    VecParams p = {x, ld, dim, seg_off, n_seg, n_rows, metric_mask, normalize, cent, cent_unit, cent_row_off, cent_k,
                   dist, argmin, (dim + 3) & ~3, thr, decision};
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    static int use_fast = -1;                                       // OODB200_VEC_FAST=0: one-row-per-warp kernel (A/B runs)
    if (use_fast < 0) { const char* env = getenv("OODB200_VEC_FAST"); use_fast = (env && atoi(env) == 0) ? 0 : 1; }
    const bool aligned = dim % 4 == 0 && ld % 4 == 0 && (((uintptr_t)x | (uintptr_t)cent | (uintptr_t)(cent_unit ? cent_unit : cent)) & 15) == 0;
    if (use_fast && aligned && dim <= 640) {
        // one launch per requested metric: rows in registers, centroid rows staged in shared memory
        const int nj = (dim + 127) / 128;
        const int vthreads = vf_threads(nj), vwarps = vthreads / 32;
        const size_t total = (nj <= 2 ? 110 : 220) * 1024;          // 2 CTAs / 1 CTA per SM: centroid table + row staging, then the barriers
        const int smem_floats = (int)(total / sizeof(float));
        OODB200_REQUIRE(smem_floats >= dim, "vec_score: dim %d too large", dim);
        const size_t smem = total + vwarps * sizeof(uint64_t);
        const int ctas = sms * (nj <= 2 ? 2 : 1);
        long long per = (n_rows + ctas - 1) / ctas;
        per = (per + vwarps * kVfRows - 1) / (vwarps * kVfRows) * (vwarps * kVfRows);
        const int grid = (int)((n_rows + per - 1) / per);
        for (int m = 0; m < OODB200_N_METRICS; ++m) {
            if (!(metric_mask >> m & 1)) continue;
            const VecFastKernel kern = m == OODB200_METRIC_L1 ? vec_fast_for<OODB200_METRIC_L1>(nj)
                                     : (m == OODB200_METRIC_L2 ? vec_fast_for<OODB200_METRIC_L2>(nj) : vec_fast_for<OODB200_METRIC_COS>(nj));
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: cheap
            if (e != cudaSuccess) { set_error("vec_score: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
            kern<<<grid, vthreads, smem, st>>>(p, (int)per, smem_floats);
            const int rc = check_launch("vec_score");
            if (rc) return rc;
        }
        return OODB200_OK;
    }
    const size_t smem = sizeof(float) * 2 * (size_t)p.d_pad * kVecWarps;
    OODB200_REQUIRE(smem <= 200 * 1024, "vec_score: dim %d too large", dim);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(vec_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("vec_score: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    }
    long long grid = (n_rows + kVecWarps - 1) / kVecWarps;
    if (grid > (long long)sms * 32) grid = (long long)sms * 32;
    vec_score_kernel<<<(int)grid, kVecThreads, smem, st>>>(p);
    return check_launch("vec_score");
}

extern "C" int oodb200_dist_indness_f32(const float* dist, const int32_t* slot, int64_t n, const double* thr,
                                        const double* dmin, const double* dmax, int clip, float* out, void* stream) {
    OODB200_REQUIRE(n >= 0, "dist_indness: negative n");
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(dist && slot && thr && dmin && dmax && out, "dist_indness: null pointer");
    dist_indness_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dist, slot, n, thr, dmin, dmax, clip, out);
    return check_launch("dist_indness");
}

extern "C" int oodb200_normalize_rows_f32(const float* x, int64_t ld, int dim, int64_t n_rows, float* out, int64_t out_ld,
                                          void* stream) {
    OODB200_REQUIRE(dim > 0 && n_rows >= 0 && ld >= dim && out_ld >= dim, "normalize_rows: bad size");
    if (n_rows == 0) return OODB200_OK;
    OODB200_REQUIRE(x && out, "normalize_rows: null pointer");
    long long grid = (n_rows + 7) / 8;
    if (grid > 148LL * 32) grid = 148LL * 32;
    normalize_rows_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(x, ld, dim, n_rows, out, out_ld);
    return check_launch("normalize_rows");
}

extern "C" int oodb200_radix_hist_u32(const float* scores, const int64_t* seg_off, int n_seg, int64_t n_rows,
                                      const uint32_t* prefix, int shift, int bits, uint32_t* hist, uint32_t* minmax,
                                      void* stream) {
    OODB200_REQUIRE(n_seg >= 0 && n_rows >= 0, "radix_hist: negative size");
    OODB200_REQUIRE(bits >= 1 && bits <= 12 && shift >= 0 && shift + bits <= 32, "radix_hist: shift %d bits %d", shift, bits);
    if (n_seg == 0 || n_rows == 0) return OODB200_OK;
    OODB200_REQUIRE(scores && seg_off && hist, "radix_hist: null pointer");
    OODB200_REQUIRE(shift + bits >= 32 || prefix, "radix_hist: prefix needed below the top pass");
    OODB200_REQUIRE(n_seg <= 65535, "radix_hist: too many segments");
    long long per_seg = (n_rows / (n_seg > 0 ? n_seg : 1) + 256 * 16 - 1) / (256 * 16);
    if (per_seg < 1) per_seg = 1;
    if (per_seg > 64) per_seg = 64;
    dim3 grid((unsigned)per_seg, (unsigned)n_seg);
    radix_hist_kernel<<<grid, 256, sizeof(uint32_t) << bits, (cudaStream_t)stream>>>(scores, seg_off, n_seg, prefix, shift,
                                                                                     bits, hist, minmax);
    return check_launch("radix_hist");
}
