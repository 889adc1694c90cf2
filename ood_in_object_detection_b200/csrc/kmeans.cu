// K4: k-means (Lloyd) over segmented data -- every (class, stride) problem in one launch, sm_100a.
//
// Replaces sklearn's `lloyd_iter_chunked_dense` (+ `_average_centers`, `_center_shift`) behind
// `KMeans(n_clusters=k, random_state=10).fit_predict(X)`  (/root/reference/cluster_utils.py:62-73, called per
// (class, stride) from /root/reference/ood_utils.py:2345), and the distance passes of `_kmeans_plusplus`.
//
//  kmeans_step    one CTA per fixed block of rows of one segment.  Centroids of the segment are resident in
//                 shared memory.  Each warp takes one row into registers (128-bit coalesced loads), computes
//                 ||c||^2 - 2 x.c against all centroids (16 partial dots per lane, one butterfly reduction per
//                 16 centroids) and the first-minimum label.  The 8 rows of a chunk are then added into the
//                 block's shared-memory sums column-wise (thread = column, rows in order): no atomics, so the
//                 block partial is bit-reproducible.  Partials go to global memory.
//  kmeans_reduce  sums consecutive block partials in a FIXED order -> the result does not depend on how blocks
//                 were scheduled, nor (with the two-level scheme of kmeans.py) on the number of GPUs.
//  kmeans_update  new centres (sklearn `_average_centers` arithmetic), squared centre shift per segment.
//  sqdist_cand    squared distances of candidate seeds to every row, float64 expansion cast to float32 like
//                 sklearn's `_euclidean_distances_upcast`, fused with the min against `closest` and the
//                 per-candidate potential.
// HBM-bound stream over X (4*D bytes per row per iteration); K=16 needs 2*K flop per 4 bytes = 8 flop/B, below the
// FP32 ridge, so the dot products stay on the FP32 pipe.
#include "common.cuh"

#include <float.h>
#include <limits.h>

namespace oodb200 {

constexpr int kKmThreads = 256;
constexpr int kKmWarps = kKmThreads / 32;
constexpr int kKG = 16;                       // centroids per butterfly group
constexpr unsigned kFullMask = 0xffffffffu;

struct KmParams {
    const float* x;
    int dim, d_pad;
    int k;                                    // centroids per segment (table stride)
    const int32_t* seg_k;                     // [n_seg] centroids actually used (<= k)
    const float* cent;                        // [n_seg, k, dim]
    const int32_t* block_seg;                 // [n_blocks]
    const int64_t* block_row0;                // [n_blocks + 1]... row range = [row0[b], row_end[b])
    const int64_t* block_row1;
    const int32_t* active;                    // [n_seg] or null
    int32_t* labels;
    float* psums;                             // [n_blocks, k, dim]
    float* pcounts;                           // [n_blocks, k]
    int32_t* n_changed;                       // [n_seg]
    int update;
};

__device__ __forceinline__ float butterfly16(float (&v)[kKG], int lane) {
#pragma unroll
    for (int o = 16, n = kKG; n > 1; o >>= 1, n >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(kFullMask, send, o);
        }
    }
    return v[0] + __shfl_xor_sync(kFullMask, v[0], 1);   // lanes 2u, 2u+1 hold value u
}

template <int NJ>
__global__ void __launch_bounds__(kKmThreads) kmeans_step_kernel(const KmParams p) {
    extern __shared__ __align__(16) float smem[];
    const int D = p.dim, Dp = p.d_pad, K = p.k;
    float* s_cent = smem;                          // [K][Dp]
    float* s_sum = s_cent + (size_t)K * Dp;        // [K][Dp]
    float* s_stage = s_sum + (size_t)K * Dp;       // [8][Dp]
    float* s_csn = s_stage + (size_t)kKmWarps * Dp; // [K]
    float* s_cnt = s_csn + K;                      // [K]
    __shared__ int s_lab[kKmWarps];
    __shared__ int s_changed;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int g = p.block_seg[b];
    if (p.active && !p.active[g]) return;
    const int Kg = p.seg_k[g];
    const int64_t r0 = p.block_row0[b], r1 = p.block_row1[b];
    const float* __restrict__ cg = p.cent + (size_t)g * K * D;
    for (int i = tid; i < K * Dp; i += kKmThreads) {
        const int k = i / Dp, d = i - k * Dp;
        s_cent[i] = (k < Kg && d < D) ? __ldg(cg + (size_t)k * D + d) : 0.f;
        s_sum[i] = 0.f;
    }
    if (tid < K) s_cnt[tid] = 0.f;
    if (tid == 0) s_changed = 0;
    __syncthreads();
    for (int k = warp; k < K; k += kKmWarps) {     // ||c||^2 (row_norms(centers, squared=True))
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(s_cent[k * Dp + d], s_cent[k * Dp + d], s);
        s = warp_sum(s);
        if (lane == 0) s_csn[k] = s;
    }
    __syncthreads();
    const int n_groups = (Kg + kKG - 1) / kKG;
    for (int64_t base = r0; base < r1; base += kKmWarps) {
        const int64_t r = base + warp;
        const bool have = r < r1;
        if (have) {
            const float* __restrict__ xr = p.x + r * D;
            float4 xv[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int d = lane * 4 + 128 * j;
                if (d + 3 < D) xv[j] = __ldg(reinterpret_cast<const float4*>(xr + d));
                else {
                    xv[j].x = d < D ? __ldg(xr + d) : 0.f;
                    xv[j].y = d + 1 < D ? __ldg(xr + d + 1) : 0.f;
                    xv[j].z = d + 2 < D ? __ldg(xr + d + 2) : 0.f;
                    xv[j].w = 0.f;
                }
                if (p.update && d < Dp) *reinterpret_cast<float4*>(s_stage + (size_t)warp * Dp + d) = xv[j];
            }
            float best = FLT_MAX;
            int barg = 0;
            if (p.update == 2) {                     // sums of the GIVEN labels (member means), no re-assignment
                barg = p.labels[r];
                if (barg < 0 || barg >= Kg) barg = 0;
            }
            for (int grp = 0; grp < (p.update == 2 ? 0 : n_groups); ++grp) {
                float dot[kKG];
#pragma unroll
                for (int u = 0; u < kKG; ++u) {
                    const float* __restrict__ ck = s_cent + (size_t)min(grp * kKG + u, K - 1) * Dp;
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int d = lane * 4 + 128 * j;
                        if (d < Dp) {
                            const float4 c4 = *reinterpret_cast<const float4*>(ck + d);
                            s = fmaf(xv[j].x, c4.x, s); s = fmaf(xv[j].y, c4.y, s);
                            s = fmaf(xv[j].z, c4.z, s); s = fmaf(xv[j].w, c4.w, s);
                        }
                    }
                    dot[u] = s;
                }
                const float tot = butterfly16(dot, lane);
                const int kk = grp * kKG + (lane >> 1);
                float pd = kk < Kg ? fmaf(-2.0f, tot, s_csn[kk]) : FLT_MAX;   // gemm(alpha=-2, beta=1) on ||c||^2
                int pk = kk < Kg ? kk : INT_MAX;
#pragma unroll
                for (int o = 16; o > 1; o >>= 1) {                           // first minimum across the group
                    const float od = __shfl_xor_sync(kFullMask, pd, o);
                    const int ok = __shfl_xor_sync(kFullMask, pk, o);
                    if (od < pd || (od == pd && ok < pk)) { pd = od; pk = ok; }
                }
                if (pd < best) { best = pd; barg = pk; }                     // strict <: earlier group wins ties
            }
            if (lane == 0) {
                s_lab[warp] = barg;
                if (p.labels[r] != barg) atomicAdd(&s_changed, 1);
                p.labels[r] = barg;
            }
        } else if (lane == 0) {
            s_lab[warp] = -1;
        }
        if (!p.update) continue;
        __syncthreads();
        // column-wise accumulation of the chunk's rows, in row order: deterministic
        for (int d = tid; d < D; d += kKmThreads) {
#pragma unroll
            for (int w = 0; w < kKmWarps; ++w) {
                const int l = s_lab[w];
                if (l >= 0) s_sum[(size_t)l * Dp + d] += s_stage[(size_t)w * Dp + d];
            }
        }
        if (tid == 0) {
#pragma unroll
            for (int w = 0; w < kKmWarps; ++w)
                if (s_lab[w] >= 0) s_cnt[s_lab[w]] += 1.f;
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0 && s_changed) atomicAdd(&p.n_changed[g], s_changed);
    if (p.update) {
        float* __restrict__ ps = p.psums + (size_t)b * K * D;
        for (int i = tid; i < K * D; i += kKmThreads) {
            const int k = i / D, d = i - k * D;
            ps[i] = s_sum[(size_t)k * Dp + d];
        }
        if (tid < K) p.pcounts[(size_t)b * K + tid] = s_cnt[tid];
    }
}

// ---- fast path for K <= 16: register tiling.
// A warp takes R = 4 rows at a time (x in registers: R * NJ float4 per lane), so every centroid value read from shared
// memory feeds 4 rows (the one-row kernel above is bound by shared-memory bandwidth: 1 LDS.128 per 4 FMAs).  The
// R * 8 partial dots of a group of 8 centroids are reduced with ONE transposing butterfly (31 shuffles for 32 totals).
// Partial sums live in REGISTERS: thread t owns columns t, t + 256, ... of every centroid; the label of a row is
// warp-uniform, so `switch (label)` picks the accumulator without divergence and rows are added in row order
// (bit-reproducible block partials, same order as the one-row kernel).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    }
}

constexpr int kFastR = 4;
constexpr int kFastRows = kKmWarps * kFastR;          // rows per chunk

template <int N>
__device__ __forceinline__ float transpose_reduce(float (&v)[N], int lane) {   // lane l ends with the total of v[l], N == 32
    int o = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, o >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(kFullMask, send, o);
        }
    }
    return v[0];
}

template <int NJ>
__global__ void __launch_bounds__(kKmThreads, 1) kmeans_step_fast_kernel(const KmParams p) {
    extern __shared__ __align__(128) float smem[];
    constexpr int NC = (NJ * 128 + kKmThreads - 1) / kKmThreads;      // columns per thread
    const int D = p.dim, K = p.k;                                     // D % 4 == 0 (row pitch = D), K <= 16
    float* s_stage = smem;                                            // [2][kFastRows][D]: TMA landing zone, double buffered
    float* s_cent = s_stage + (size_t)2 * kFastRows * D;              // [16][D]
    float* s_csn = s_cent + (size_t)16 * D;                           // [16]
    __shared__ __align__(8) uint64_t s_full[2];
    __shared__ int s_lab[kFastRows];
    __shared__ unsigned s_mask[16];
    __shared__ int s_changed;
    static_assert(kFastRows == 32, "one ballot covers the chunk");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int g = p.block_seg[b];
    if (p.active && !p.active[g]) return;
    const int Kg = p.seg_k[g];
    const int64_t r0 = p.block_row0[b], r1 = p.block_row1[b];
    const int n_chunks = (int)((r1 - r0 + kFastRows - 1) / kFastRows);
    auto issue = [&](int c) {                                         // thread 0: one bulk copy = the chunk's rows (contiguous)
        const int64_t ra = r0 + (int64_t)c * kFastRows;
        const uint32_t bytes = (uint32_t)(min((int64_t)kFastRows, r1 - ra) * D * 4);
        mbar_expect_tx(&s_full[c & 1], bytes);
        bulk_g2s(s_stage + (size_t)(c & 1) * kFastRows * D, p.x + ra * D, bytes, &s_full[c & 1]);
    };
    if (tid == 0) {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_changed = 0;
        issue(0);
        if (n_chunks > 1) issue(1);
    }
    const float* __restrict__ cg = p.cent + (size_t)g * K * D;
    for (int i = tid; i < 16 * D; i += kKmThreads) {
        const int k = i / D, d = i - k * D;
        s_cent[i] = k < Kg ? __ldg(cg + (size_t)k * D + d) : 0.f;
    }
    __syncthreads();
    for (int k = warp; k < 16; k += kKmWarps) {                       // ||c||^2 (row_norms(centers, squared=True))
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(s_cent[k * D + d], s_cent[k * D + d], s);
        s = warp_sum(s);
        if (lane == 0) s_csn[k] = s;
    }
    float acc[16][NC];
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[k][c] = 0.f;
    float cnt = 0.f;                                                  // thread k (< 16) counts the members of cluster k
    __syncthreads();
    for (int c = 0; c < n_chunks; ++c) {
        const float* __restrict__ st = s_stage + (size_t)(c & 1) * kFastRows * D;
        const int64_t rw = r0 + (int64_t)c * kFastRows + warp * kFastR;
        mbar_wait(&s_full[c & 1], (uint32_t)(c >> 1) & 1u);
        float4 xv[kFastR][NJ];
#pragma unroll
        for (int r = 0; r < kFastR; ++r)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int d = lane * 4 + 128 * j;
                xv[r][j] = (rw + r < r1 && d < D) ? *reinterpret_cast<const float4*>(st + (size_t)(warp * kFastR + r) * D + d)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        float best = FLT_MAX;                                         // lane l: row l >> 3 of this warp
        int barg = 0;
        if (p.update == 2) {                                          // sums of the GIVEN labels (member means)
            const int64_t rr = rw + (lane >> 3);
            barg = rr < r1 ? p.labels[rr] : 0;
            if (barg < 0 || barg >= Kg) barg = 0;
        } else {
            for (int kg = 0; kg < Kg; kg += 8) {
                float dot[kFastR * 8];
#pragma unroll
                for (int i = 0; i < kFastR * 8; ++i) dot[i] = 0.f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int d = lane * 4 + 128 * j;
                    if (d < D) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float4 c4 = *reinterpret_cast<const float4*>(s_cent + (size_t)(kg + u) * D + d);
#pragma unroll
                            for (int r = 0; r < kFastR; ++r) {
                                float s = dot[r * 8 + u];
                                s = fmaf(xv[r][j].x, c4.x, s); s = fmaf(xv[r][j].y, c4.y, s);
                                s = fmaf(xv[r][j].z, c4.z, s); s = fmaf(xv[r][j].w, c4.w, s);
                                dot[r * 8 + u] = s;
                            }
                        }
                    }
                }
                const float tot = transpose_reduce<kFastR * 8>(dot, lane);   // lane l: row l >> 3, centroid kg + (l & 7)
                const int kk = kg + (lane & 7);
                float pd = kk < Kg ? fmaf(-2.0f, tot, s_csn[kk]) : FLT_MAX;  // gemm(alpha=-2, beta=1) on ||c||^2
                int pk = kk < Kg ? kk : INT_MAX;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {                            // first minimum over the 8 centroids of the group
                    const float od = __shfl_xor_sync(kFullMask, pd, o);
                    const int ok = __shfl_xor_sync(kFullMask, pk, o);
                    if (od < pd || (od == pd && ok < pk)) { pd = od; pk = ok; }
                }
                if (pd < best) { best = pd; barg = pk; }                     // strict <: earlier group wins ties
            }
        }
        if ((lane & 7) == 0) {
            const int r = lane >> 3;
            const int64_t rr = rw + r;
            if (rr < r1) {
                s_lab[warp * kFastR + r] = barg;
                if (p.labels[rr] != barg) atomicAdd(&s_changed, 1);
                p.labels[rr] = barg;
            } else {
                s_lab[warp * kFastR + r] = -1;
            }
        }
        __syncthreads();
        if (p.update) {
            // Register accumulation, cluster by cluster: the rows of the chunk that carry label k come from a ballot
            // (warp 0), are visited in row order (bit-reproducible block partial), and k is a compile-time index of
            // the accumulator array -- no label-dependent branch, no shared-memory read-modify-write.
            if (warp == 0) {
                const int l = s_lab[lane];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const unsigned m = __ballot_sync(kFullMask, l == k);
                    if (lane == k) s_mask[k] = m;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                unsigned m = s_mask[k];
                if (tid == k) cnt += (float)__popc(m);
                while (m) {                                           // 4 rows per trip: their loads overlap, the adds stay in row order
                    int w[4];
                    float v[4][NC];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        w[q] = m ? __ffs(m) - 1 : -1;
                        m &= m - 1;                                   // 0 & anything stays 0
#pragma unroll
                        for (int cc = 0; cc < NC; ++cc) {
                            const int d = tid + cc * kKmThreads;
                            v[q][cc] = (w[q] >= 0 && d < D) ? st[(size_t)w[q] * D + d] : 0.f;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int cc = 0; cc < NC; ++cc)
                            if (w[q] >= 0) acc[k][cc] += v[q][cc];
                }
            }
            __syncthreads();                                          // the buffer is free again
        }
        if (tid == 0 && c + 2 < n_chunks) issue(c + 2);
    }
    __syncthreads();
    if (tid == 0 && s_changed) atomicAdd(&p.n_changed[g], s_changed);
    if (p.update) {
        float* __restrict__ ps = p.psums + (size_t)b * K * D;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < K)
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) {
                    const int d = tid + cc * kKmThreads;
                    if (d < D) ps[(size_t)k * D + d] = acc[k][cc];
                }
        if (tid < K) p.pcounts[(size_t)b * K + tid] = cnt;
    }
}

// out[grp, e] = sum over blocks b in [first[grp], first[grp+1]) of in[b, e], sequentially in b
__global__ void kmeans_reduce_kernel(const float* __restrict__ in, const int32_t* __restrict__ first, int n_groups,
                                     int64_t elems, float* __restrict__ out) {
    const int grp = blockIdx.y;
    const int b0 = first[grp], b1 = first[grp + 1];
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < elems; e += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int b = b0; b < b1; ++b) s += in[(size_t)b * elems + e];
        out[(size_t)grp * elems + e] = s;
    }
}

// The block partials of one Lloyd iteration in ONE launch: sums and counts reduced in fixed block order (like
// kmeans_reduce_kernel), the changed-label counters converted to float32 next to them (they ride in the all-reduce buffer)
// and cleared for the next iteration.
__global__ void kmeans_reduce_step_kernel(const float* __restrict__ psums, const float* __restrict__ pcounts,
                                          const int32_t* __restrict__ first, int64_t elems_s, int64_t elems_c,
                                          float* __restrict__ out_s, float* __restrict__ out_c, int32_t* __restrict__ n_changed,
                                          float* __restrict__ chg_f) {
    const int grp = blockIdx.y;
    const int b0 = first[grp], b1 = first[grp + 1];
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < elems_s + elems_c; e += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        if (e < elems_s) {
            for (int b = b0; b < b1; ++b) s += psums[(size_t)b * elems_s + e];
            out_s[(size_t)grp * elems_s + e] = s;
        } else {
            const int64_t ec = e - elems_s;
            for (int b = b0; b < b1; ++b) s += pcounts[(size_t)b * elems_c + ec];
            out_c[(size_t)grp * elems_c + ec] = s;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        chg_f[grp] = (float)n_changed[grp];            // exact below 2^24 rows per segment
        n_changed[grp] = 0;
    }
}

// sklearn _average_centers + _center_shift (squared, summed per segment); empty clusters keep their old centre.
// One CTA per segment, one WARP per cluster (clusters w, w + n_warps, ...): the K rows are independent, only the shift is
// summed over the clusters, in cluster order (fixed: bit-reproducible).  An inactive (converged) segment copies its centres
// through, so that the caller can ping-pong two buffers.
// PEERS > 0: the sums / counts are not in `sums` / `counts` but in the symmetric buffers of the `PEERS` ranks of the
// fit (peer[r] + sums_off / counts_off), read straight over NVLink and added in rank order -- the all-reduce of the Lloyd
// iteration fused into the update; every rank computes identical bits.  The summed counts and changed-label counts are
// written to cnts_out / chg_out for the convergence kernel.
constexpr int kUpdThreads = 512;
constexpr int kUpdMaxPeers = 8;                        // peers whose loads are kept in flight together (more: in groups of 8)

__global__ void __launch_bounds__(kUpdThreads) kmeans_update_kernel(const float* __restrict__ sums, const float* __restrict__ counts,
                                     const float* const* __restrict__ peer, int n_peers, int64_t counts_off, int64_t chg_off,
                                     uint32_t* const* __restrict__ peer_flags, int my_rank, uint32_t epoch,
                                     const float* __restrict__ cent_old, const int32_t* __restrict__ seg_k,
                                     const int32_t* __restrict__ active, int k, int dim, float* __restrict__ cent_new,
                                     float* __restrict__ shift_sq, int32_t* __restrict__ n_empty,
                                     float* __restrict__ cnts_out, float* __restrict__ chg_out) {
    const int g = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    if (n_peers > 0 && peer_flags) {
        // Cross-GPU barrier of the iteration, inside the kernel: this rank's partials were written by earlier launches of the
        // stream; CTA 0 publishes `epoch` in every peer's flag array (release, system scope), every CTA waits until all peers
        // have published theirs in the LOCAL array (acquire).  Bounded: a missing peer traps instead of hanging the GPU.
        if (blockIdx.x == 0 && threadIdx.x < n_peers) {
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[threadIdx.x] + my_rank), "r"(epoch) : "memory");
        }
        if (threadIdx.x < n_peers) {
            const uint32_t* f = peer_flags[my_rank] + threadIdx.x;
            uint32_t v = 0;
            const long long t0 = clock64();
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if ((int32_t)(v - epoch) >= 0) break;
                if (clock64() - t0 > 8000000000LL) __trap();
            }
        }
        __syncthreads();
    }
    if (n_peers > 0 && threadIdx.x == 0 && chg_out) {   // changed labels of the segment over all ranks (exact: small integers in float32)
        float c = 0.f;
        for (int r = 0; r < n_peers; ++r) c += peer[r][chg_off + g];
        chg_out[g] = c;
    }
    const size_t seg = (size_t)g * k * dim;
    if (active && !active[g]) {
        for (int i = threadIdx.x; i < k * dim; i += blockDim.x) cent_new[seg + i] = cent_old[seg + i];
        if (threadIdx.x == 0) { shift_sq[g] = 0.f; n_empty[g] = 0; }
        return;
    }
    extern __shared__ float s_upd[];                   // [k] counts, [k] squared shift per cluster, [k * dim / 4] partial shifts
    float* s_w = s_upd;
    float* s_shift = s_upd + k;
    float* s_part = s_upd + 2 * k;
    const int Kg = seg_k[g];
    for (int kk = threadIdx.x; kk < k; kk += blockDim.x) {            // cluster sizes (over the ranks, in rank order)
        float w = 0.f;
        if (kk < Kg) {
            if (n_peers > 0) {
                for (int r = 0; r < n_peers; ++r) w += peer[r][counts_off + (size_t)g * k + kk];
                if (cnts_out) cnts_out[(size_t)g * k + kk] = w;
            } else {
                w = counts[(size_t)g * k + kk];
            }
        }
        s_w[kk] = w;
    }
    __syncthreads();
    const bool vec = (dim & 3) == 0;
    const int nv = vec ? k * dim / 4 : 0;
    // every thread owns float4 elements of the segment's K x dim block: all its loads (x peers) are in flight together
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const int kk = (4 * i) / dim;
        const float4 old = *reinterpret_cast<const float4*>(cent_old + seg + 4 * (size_t)i);
        float4 c = old;
        float part = 0.f;
        if (kk < Kg) {
            float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n_peers > 0) {
                for (int r0 = 0; r0 < n_peers; r0 += kUpdMaxPeers) {
                    float4 v[kUpdMaxPeers];
#pragma unroll
                    for (int r = 0; r < kUpdMaxPeers; ++r)
                        v[r] = r0 + r < n_peers ? *reinterpret_cast<const float4*>(peer[r0 + r] + seg + 4 * (size_t)i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int r = 0; r < kUpdMaxPeers; ++r)
                        if (r0 + r < n_peers) { sm.x += v[r].x; sm.y += v[r].y; sm.z += v[r].z; sm.w += v[r].w; }
                }
            } else {
                sm = *reinterpret_cast<const float4*>(sums + seg + 4 * (size_t)i);
            }
            const float w = s_w[kk];
            if (w > 0.f) {
                const float alpha = (float)(1.0 / (double)w);
                c = make_float4(sm.x * alpha, sm.y * alpha, sm.z * alpha, sm.w * alpha);
            }
            const float d0 = c.x - old.x, d1 = c.y - old.y, d2 = c.z - old.z, d3 = c.w - old.w;
            part = fmaf(d0, d0, part); part = fmaf(d1, d1, part); part = fmaf(d2, d2, part); part = fmaf(d3, d3, part);
        }
        *reinterpret_cast<float4*>(cent_new + seg + 4 * (size_t)i) = c;
        s_part[i] = part;
    }
    __syncthreads();
    for (int kk = warp; kk < k; kk += n_warps) {       // squared shift of every cluster: fixed lane partition + butterfly
        float sh = 0.f;
        if (vec) {
            const int per = dim / 4;
            for (int j = lane; j < per; j += 32) sh += s_part[kk * per + j];
        } else if (kk < Kg) {                          // dim % 4 != 0: scalar path
            const float w = s_w[kk];
            const float alpha = w > 0.f ? (float)(1.0 / (double)w) : 0.f;
            for (int d = lane; d < dim; d += 32) {
                const size_t e = seg + (size_t)kk * dim + d;
                float sm = 0.f;
                if (n_peers > 0) { for (int r = 0; r < n_peers; ++r) sm += peer[r][e]; } else sm = sums[e];
                const float c = w > 0.f ? sm * alpha : cent_old[e];
                cent_new[e] = c;
                const float df = c - cent_old[e];
                sh = fmaf(df, df, sh);
            }
        } else {
            for (int d = lane; d < dim; d += 32) cent_new[seg + (size_t)kk * dim + d] = cent_old[seg + (size_t)kk * dim + d];
        }
        sh = warp_sum(sh);                             // (sqrt(t))^2 in sklearn; equal up to one rounding
        if (lane == 0) s_shift[kk] = sh;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        int empties = 0;
        for (int kk = 0; kk < Kg; ++kk) { tot += s_shift[kk]; empties += !(s_w[kk] > 0.f); }
        shift_sq[g] = tot;
        n_empty[g] = empties;
    }
}

// Convergence bookkeeping of one Lloyd iteration on the device (sklearn _kmeans_single_lloyd, _kmeans.py:712-740): a
// segment stops when no label changed (strict convergence) or when its squared centre shift is <= tol; state rows:
// [0] iterations run, [1] strict, [2] needs the final E-step, [3] empty clusters seen.  The host only polls any_active.
__global__ void kmeans_converge_kernel(const int32_t* __restrict__ n_changed_i, const float* __restrict__ n_changed_f,
                                       const float* __restrict__ shift, const int32_t* __restrict__ n_empty,
                                       const double* __restrict__ tol_abs, const float* __restrict__ cnts, int n_seg, int k,
                                       int32_t* __restrict__ active, int32_t* __restrict__ state, float* __restrict__ counts,
                                       int32_t* __restrict__ any_active) {
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    int any = 0;
    for (int g = threadIdx.x; g < n_seg; g += blockDim.x) {
        if (!active[g]) continue;
        const bool changed = n_changed_i ? n_changed_i[g] != 0 : n_changed_f[g] != 0.f;
        state[g] += 1;
        state[3 * n_seg + g] += n_empty[g];
        for (int j = 0; j < k; ++j) counts[(size_t)g * k + j] = cnts[(size_t)g * k + j];
        if (!changed) {
            state[n_seg + g] = 1;
            active[g] = 0;
        } else if ((double)shift[g] <= tol_abs[g]) {
            state[2 * n_seg + g] = 1;
            active[g] = 0;
        } else {
            any = 1;
        }
    }
    if (any) atomicOr(&s_any, 1);
    __syncthreads();
    if (threadIdx.x == 0) any_active[0] = s_any;
}

// d[j, r] = max(float32(||y_j||^2 - 2 x_r.y_j + ||x_r||^2 in float64), 0); newd = min(closest, d); pot[j] += sum newd
struct CandParams {
    const float* x;
    int dim;
    const int64_t* seg_off;
    int n_seg;
    const float* cand;                         // [n_seg, n_cand, dim] candidate vectors
    int n_cand;
    const float* closest;                      // [n_rows] or null (first centre)
    float* out_d;                              // [n_cand, n_rows] min(closest, d)
    double* pot;                               // [n_seg, n_cand]
};

__global__ void __launch_bounds__(256) sqdist_cand_kernel(const CandParams p) {
    extern __shared__ __align__(16) double s_y[];              // [n_cand][dim] + norms
    const int g = blockIdx.y;
    const int D = p.dim, NC = p.n_cand;
    double* s_yy = s_y + (size_t)NC * D;
    __shared__ double s_pot[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NC * D; i += blockDim.x) {
        const int j = i / D, d = i - j * D;
        s_y[i] = (double)p.cand[((size_t)g * NC + j) * D + d];
    }
    __syncthreads();
    if (warp < NC) {
        double s = 0.0;
        for (int d = lane; d < D; d += 32) s += s_y[warp * D + d] * s_y[warp * D + d];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
        if (lane == 0) s_yy[warp] = s;
    }
    __syncthreads();
    const int64_t r0 = p.seg_off[g], r1 = p.seg_off[g + 1];
    const int64_t n_rows = p.seg_off[p.n_seg];
    double pot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t r = r0 + (int64_t)blockIdx.x * 8 + warp; r < r1; r += (int64_t)gridDim.x * 8) {
        const float* __restrict__ xr = p.x + r * D;
        double xx = 0.0, dot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int d = lane; d < D; d += 32) {
            const double v = (double)__ldg(xr + d);
            xx += v * v;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < NC) dot[j] += v * s_y[j * D + d];
        }
        for (int o = 16; o > 0; o >>= 1) {
            xx += __shfl_xor_sync(kFullMask, xx, o);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < NC) dot[j] += __shfl_xor_sync(kFullMask, dot[j], o);
        }
        if (lane == 0) {
            const float cl = p.closest ? p.closest[r] : FLT_MAX;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < NC) {
                    float d32 = (float)(-2.0 * dot[j] + s_yy[j] + xx);
                    d32 = fminf(fmaxf(d32, 0.f), cl);
                    p.out_d[(size_t)j * n_rows + r] = d32;
                    pot[j] += (double)d32;
                }
        }
    }
    if (lane == 0)
        for (int j = 0; j < 8; ++j) s_pot[warp][j] = pot[j];
    __syncthreads();
    if (threadIdx.x < NC) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_pot[w][threadIdx.x];
        atomicAdd(&p.pot[(size_t)g * NC + threadIdx.x], t);
    }
}

// ---- mean-centring of every segment (KMeans.fit: `X -= X.mean(axis=0)`, sklearn/cluster/_kmeans.py:1487-1497) and the
// variance behind sklearn's tolerance (`_tolerance`, :283-293).  Two streaming passes, float64 accumulation, block
// partials combined in block order -> bit-reproducible.
constexpr int kCsThreads = 160;
constexpr int kCsBlocks = 64;                  // row blocks per segment

// partial[g, b, d] = sum over the rows of block b of segment g of x[r, d] (float64)
template <bool VEC4>
__global__ void __launch_bounds__(kCsThreads) seg_colsum_kernel(const float* __restrict__ x, int dim, const int64_t* __restrict__ seg_off,
                                                                double* __restrict__ partial) {
    const int g = blockIdx.y, b = blockIdx.x;
    const int64_t s0 = seg_off[g], s1 = seg_off[g + 1];
    const int64_t per = (s1 - s0 + kCsBlocks - 1) / kCsBlocks;
    const int64_t r0 = min(s1, s0 + per * b), r1 = min(s1, r0 + per);
    double* __restrict__ out = partial + ((size_t)g * kCsBlocks + b) * dim;
    if (VEC4) {
        for (int c = threadIdx.x; c < dim / 4; c += kCsThreads) {
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            const float* __restrict__ col = x + c * 4;
            int64_t r = r0;
            for (; r + 8 <= r1; r += 8) {
                float4 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = __ldg(reinterpret_cast<const float4*>(col + (r + q) * dim));
#pragma unroll
                for (int q = 0; q < 8; ++q) { a0 += v[q].x; a1 += v[q].y; a2 += v[q].z; a3 += v[q].w; }
            }
            for (; r < r1; ++r) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(col + r * dim));
                a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
            }
            out[c * 4] = a0; out[c * 4 + 1] = a1; out[c * 4 + 2] = a2; out[c * 4 + 3] = a3;
        }
    } else {
        for (int c = threadIdx.x; c < dim; c += kCsThreads) {
            double a = 0;
            for (int64_t r = r0; r < r1; ++r) a += __ldg(x + r * dim + c);
            out[c] = a;
        }
    }
}

// sums[g, d] = sum_b partial[g, b, d] in increasing b
__global__ void seg_colsum_reduce_kernel(const double* __restrict__ partial, int dim, double* __restrict__ sums) {
    const int g = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    double t = 0;
    for (int b = 0; b < kCsBlocks; ++b) t += partial[((size_t)g * kCsBlocks + b) * dim + d];
    sums[(size_t)g * dim + d] = t;
}

// out[r, d] = x[r, d] - mean[g, d];  sq_partial[g, b] = sum over the block of out^2 (float64)
template <bool VEC4>
__global__ void __launch_bounds__(kCsThreads) seg_center_kernel(const float* __restrict__ x, int dim, const int64_t* __restrict__ seg_off,
                                                                const float* __restrict__ mean, float* __restrict__ out,
                                                                double* __restrict__ sq_partial) {
    __shared__ double s_red[kCsThreads];
    const int g = blockIdx.y, b = blockIdx.x;
    const int64_t s0 = seg_off[g], s1 = seg_off[g + 1];
    const int64_t per = (s1 - s0 + kCsBlocks - 1) / kCsBlocks;
    const int64_t r0 = min(s1, s0 + per * b), r1 = min(s1, r0 + per);
    const float* __restrict__ mg = mean + (size_t)g * dim;
    double sq = 0;
    if (VEC4) {
        for (int c = threadIdx.x; c < dim / 4; c += kCsThreads) {
            const float4 m = *reinterpret_cast<const float4*>(mg + c * 4);
            int64_t r = r0;
            for (; r + 4 <= r1; r += 4) {
                float4 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const float4*>(x + (r + q) * dim + c * 4));
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v[q].x -= m.x; v[q].y -= m.y; v[q].z -= m.z; v[q].w -= m.w;
                    sq += (double)v[q].x * v[q].x; sq += (double)v[q].y * v[q].y;
                    sq += (double)v[q].z * v[q].z; sq += (double)v[q].w * v[q].w;
                    *reinterpret_cast<float4*>(out + (r + q) * dim + c * 4) = v[q];
                }
            }
            for (; r < r1; ++r) {
                float4 v = __ldg(reinterpret_cast<const float4*>(x + r * dim + c * 4));
                v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
                sq += (double)v.x * v.x; sq += (double)v.y * v.y; sq += (double)v.z * v.z; sq += (double)v.w * v.w;
                *reinterpret_cast<float4*>(out + r * dim + c * 4) = v;
            }
        }
    } else {
        for (int c = threadIdx.x; c < dim; c += kCsThreads) {
            const float m = mg[c];
            for (int64_t r = r0; r < r1; ++r) {
                const float v = __ldg(x + r * dim + c) - m;
                sq += (double)v * v;
                out[r * dim + c] = v;
            }
        }
    }
    s_red[threadIdx.x] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < kCsThreads; ++i) t += s_red[i];
        sq_partial[(size_t)g * kCsBlocks + b] = t;
    }
}

__global__ void seg_sq_reduce_kernel(const double* __restrict__ sq_partial, double* __restrict__ sq) {
    const int g = blockIdx.x;
    if (threadIdx.x == 0) {
        double t = 0;
        for (int b = 0; b < kCsBlocks; ++b) t += sq_partial[(size_t)g * kCsBlocks + b];
        sq[g] = t;
    }
}

}  // namespace oodb200

using namespace oodb200;

extern "C" int64_t oodb200_kmeans_smem_bytes(int k, int dim) {
    const int dp = (dim + 3) & ~3;
    return (int64_t)sizeof(float) * ((size_t)2 * k * dp + (size_t)kKmWarps * dp + 2 * k);
}

extern "C" int oodb200_kmeans_step_f32(const float* x, int dim, int n_seg, int k, const int32_t* seg_k, const float* cent,
                                       const int32_t* block_seg, const int64_t* block_row0, const int64_t* block_row1,
                                       int n_blocks, const int32_t* active, int32_t* labels, float* psums, float* pcounts,
                                       int32_t* n_changed, int update, void* stream) {
    OODB200_REQUIRE(dim > 0 && dim <= 1024 && k > 0 && n_seg >= 0 && n_blocks >= 0, "kmeans_step: bad size (dim <= 1024)");
    if (n_blocks == 0) return OODB200_OK;
    OODB200_REQUIRE(x && seg_k && cent && block_seg && block_row0 && block_row1 && labels && n_changed, "kmeans_step: null pointer");
    OODB200_REQUIRE(!update || (psums && pcounts), "kmeans_step: update needs the partial buffers");
    const int64_t smem = oodb200_kmeans_smem_bytes(k, dim);
    OODB200_REQUIRE(smem <= 220 * 1024, "kmeans_step: k*dim = %d*%d does not fit shared memory (%lld B)", k, dim, (long long)smem);
    KmParams p = {x, dim, (dim + 3) & ~3, k, seg_k, cent, block_seg, block_row0, block_row1, active, labels, psums, pcounts,
                  n_changed, update};
    const int nj = (dim + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (k <= 16 && nj <= 6 && dim % 4 == 0 && ((uintptr_t)x & 15) == 0) {   // TMA-fed register-tiled fast path
        const size_t fsmem = sizeof(float) * ((size_t)16 * dim + (size_t)2 * kFastRows * dim + 16);
#define OODB200_KMF_LAUNCH(NJ)                                                                                        \
    case NJ: {                                                                                                         \
        cudaError_t e = cudaFuncSetAttribute(kmeans_step_fast_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)fsmem);                                                              \
        if (e != cudaSuccess) { set_error("kmeans_step: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }        \
        kmeans_step_fast_kernel<NJ><<<n_blocks, kKmThreads, fsmem, st>>>(p);                                           \
    } break;
        switch (nj) {
            OODB200_KMF_LAUNCH(1) OODB200_KMF_LAUNCH(2) OODB200_KMF_LAUNCH(3) OODB200_KMF_LAUNCH(4)
            OODB200_KMF_LAUNCH(5) OODB200_KMF_LAUNCH(6)
        }
#undef OODB200_KMF_LAUNCH
        return check_launch("kmeans_step");
    }
#define OODB200_KM_LAUNCH(NJ)                                                                                          \
    case NJ: {                                                                                                         \
        if (smem > 48 * 1024) {                                                                                        \
            cudaError_t e = cudaFuncSetAttribute(kmeans_step_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                                 (int)smem);                                                           \
            if (e != cudaSuccess) { set_error("kmeans_step: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }    \
        }                                                                                                              \
        kmeans_step_kernel<NJ><<<n_blocks, kKmThreads, smem, st>>>(p);                                                 \
    } break;
    switch (nj) {
        OODB200_KM_LAUNCH(1) OODB200_KM_LAUNCH(2) OODB200_KM_LAUNCH(3) OODB200_KM_LAUNCH(4)
        OODB200_KM_LAUNCH(5) OODB200_KM_LAUNCH(6) OODB200_KM_LAUNCH(7) OODB200_KM_LAUNCH(8)
        default: set_error("kmeans_step: dim %d", dim); return OODB200_ERR_INVALID;
    }
#undef OODB200_KM_LAUNCH
    return check_launch("kmeans_step");
}

extern "C" int oodb200_kmeans_reduce_f32(const float* in, const int32_t* first, int n_groups, int64_t elems, float* out,
                                         void* stream) {
    OODB200_REQUIRE(n_groups >= 0 && elems >= 0, "kmeans_reduce: negative size");
    if (n_groups == 0 || elems == 0) return OODB200_OK;
    OODB200_REQUIRE(in && first && out, "kmeans_reduce: null pointer");
    OODB200_REQUIRE(n_groups <= 65535, "kmeans_reduce: too many groups");
    long long gx = (elems + 255) / 256;
    if (gx > 1024) gx = 1024;
    kmeans_reduce_kernel<<<dim3((unsigned)gx, (unsigned)n_groups), 256, 0, (cudaStream_t)stream>>>(in, first, n_groups, elems, out);
    return check_launch("kmeans_reduce");
}

extern "C" int oodb200_kmeans_reduce_step_f32(const float* psums, const float* pcounts, const int32_t* first, int n_groups,
                                              int64_t elems_s, int64_t elems_c, float* out_s, float* out_c, int32_t* n_changed,
                                              float* chg_f, void* stream) {
    OODB200_REQUIRE(n_groups >= 0 && elems_s >= 0 && elems_c >= 0, "kmeans_reduce_step: negative size");
    if (n_groups == 0) return OODB200_OK;
    OODB200_REQUIRE(psums && pcounts && first && out_s && out_c && n_changed && chg_f, "kmeans_reduce_step: null pointer");
    OODB200_REQUIRE(n_groups <= 65535, "kmeans_reduce_step: too many groups");
    long long gx = (elems_s + elems_c + 255) / 256;
    if (gx > 1024) gx = 1024;
    if (gx < 1) gx = 1;
    kmeans_reduce_step_kernel<<<dim3((unsigned)gx, (unsigned)n_groups), 256, 0, (cudaStream_t)stream>>>(psums, pcounts, first, elems_s,
                                                                                                       elems_c, out_s, out_c, n_changed, chg_f);
    return check_launch("kmeans_reduce_step");
}

extern "C" int oodb200_kmeans_update_f32(const float* sums, const float* counts, const float* cent_old, const int32_t* seg_k,
                                         const int32_t* active, int n_seg, int k, int dim, float* cent_new, float* shift_sq,
                                         int32_t* n_empty, void* stream) {
    OODB200_REQUIRE(n_seg >= 0 && k > 0 && dim > 0 && k <= 4096, "kmeans_update: bad size");
    if (n_seg == 0) return OODB200_OK;
    OODB200_REQUIRE(sums && counts && cent_old && seg_k && cent_new && shift_sq && n_empty, "kmeans_update: null pointer");
    const size_t smem = sizeof(float) * (2 * (size_t)k + (size_t)k * dim / 4 + 4);
    OODB200_REQUIRE(smem <= 200 * 1024, "kmeans_update: k * dim too large");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kmeans_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("kmeans_update: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    }
    kmeans_update_kernel<<<n_seg, kUpdThreads, smem, (cudaStream_t)stream>>>(
        sums, counts, nullptr, 0, 0, 0, nullptr, 0, 0u, cent_old, seg_k, active, k, dim, cent_new, shift_sq, n_empty, nullptr, nullptr);
    return check_launch("kmeans_update");
}

extern "C" int oodb200_kmeans_update_peers_f32(const float* const* peer_bufs, int n_peers, int64_t counts_off, int64_t chg_off,
                                               uint32_t* const* peer_flags, int my_rank, uint32_t epoch, const float* cent_old, const int32_t* seg_k, const int32_t* active, int n_seg,
                                               int k, int dim, float* cent_new, float* shift_sq, int32_t* n_empty,
                                               float* cnts_out, float* chg_out, void* stream) {
    OODB200_REQUIRE(n_seg >= 0 && k > 0 && dim > 0 && k <= 4096, "kmeans_update_peers: bad size");
    OODB200_REQUIRE(n_peers >= 1 && n_peers <= 64, "kmeans_update_peers: %d peers", n_peers);
    OODB200_REQUIRE(!peer_flags || (my_rank >= 0 && my_rank < n_peers), "kmeans_update_peers: rank %d of %d", my_rank, n_peers);
    if (n_seg == 0) return OODB200_OK;
    OODB200_REQUIRE(peer_bufs && cent_old && seg_k && cent_new && shift_sq && n_empty && cnts_out && chg_out,
                    "kmeans_update_peers: null pointer");
    const size_t smem = sizeof(float) * (2 * (size_t)k + (size_t)k * dim / 4 + 4);
    OODB200_REQUIRE(smem <= 200 * 1024, "kmeans_update_peers: k * dim too large");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kmeans_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("kmeans_update_peers: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    }
    kmeans_update_kernel<<<n_seg, kUpdThreads, smem, (cudaStream_t)stream>>>(
        nullptr, nullptr, peer_bufs, n_peers, counts_off, chg_off, peer_flags, my_rank, epoch, cent_old, seg_k, active, k, dim,
        cent_new, shift_sq, n_empty, cnts_out, chg_out);
    return check_launch("kmeans_update_peers");
}

extern "C" int oodb200_kmeans_converge_f32(const int32_t* n_changed_i, const float* n_changed_f, const float* shift,
                                           const int32_t* n_empty, const double* tol_abs, const float* cnts, int n_seg, int k,
                                           int32_t* active, int32_t* state, float* counts, int32_t* any_active, void* stream) {
    OODB200_REQUIRE(n_seg >= 0 && k > 0, "kmeans_converge: bad size");
    OODB200_REQUIRE(any_active, "kmeans_converge: null pointer");
    OODB200_REQUIRE(n_seg == 0 || ((n_changed_i || n_changed_f) && shift && n_empty && tol_abs && cnts && active && state && counts),
                    "kmeans_converge: null pointer");
    kmeans_converge_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(n_changed_i, n_changed_f, shift, n_empty, tol_abs, cnts, n_seg, k,
                                                                 active, state, counts, any_active);
    return check_launch("kmeans_converge");
}

extern "C" int oodb200_sqdist_cand_f32(const float* x, int dim, const int64_t* seg_off, int n_seg, int64_t max_seg_rows,
                                       const float* cand, int n_cand, const float* closest, float* out_d, double* pot,
                                       void* stream) {
    OODB200_REQUIRE(dim > 0 && n_seg >= 0 && n_cand >= 1 && n_cand <= 8, "sqdist_cand: bad size (n_cand <= 8)");
    if (n_seg == 0 || max_seg_rows == 0) return OODB200_OK;
    OODB200_REQUIRE(x && seg_off && cand && out_d && pot, "sqdist_cand: null pointer");
    OODB200_REQUIRE(n_seg <= 65535, "sqdist_cand: too many segments");
    const size_t smem = sizeof(double) * ((size_t)n_cand * dim + 8);
    OODB200_REQUIRE(smem <= 200 * 1024, "sqdist_cand: dim too large");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sqdist_cand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("sqdist_cand: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    }
    CandParams p = {x, dim, seg_off, n_seg, cand, n_cand, closest, out_d, pot};
    long long gx = (max_seg_rows + 63) / 64;
    if (gx > 592) gx = 592;
    if (gx < 1) gx = 1;
    sqdist_cand_kernel<<<dim3((unsigned)gx, (unsigned)n_seg), 256, smem, (cudaStream_t)stream>>>(p);
    return check_launch("sqdist_cand");
}

extern "C" int64_t oodb200_segment_scratch_doubles(int n_seg, int dim) { return (int64_t)n_seg * kCsBlocks * (dim > 1 ? dim : 1); }

extern "C" int oodb200_segment_colsum_f64(const float* x, int dim, const int64_t* seg_off, int n_seg, double* scratch,
                                          double* sums, void* stream) {
    OODB200_REQUIRE(dim > 0 && n_seg >= 0 && n_seg <= 65535, "segment_colsum: bad size");
    if (n_seg == 0) return OODB200_OK;
    OODB200_REQUIRE(x && seg_off && scratch && sums, "segment_colsum: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(kCsBlocks, (unsigned)n_seg);
    if (dim % 4 == 0 && ((uintptr_t)x & 15) == 0) seg_colsum_kernel<true><<<grid, kCsThreads, 0, st>>>(x, dim, seg_off, scratch);
    else seg_colsum_kernel<false><<<grid, kCsThreads, 0, st>>>(x, dim, seg_off, scratch);
    int rc = check_launch("segment_colsum");
    if (rc) return rc;
    seg_colsum_reduce_kernel<<<dim3((unsigned)((dim + 127) / 128), (unsigned)n_seg), 128, 0, st>>>(scratch, dim, sums);
    return check_launch("segment_colsum_reduce");
}

extern "C" int oodb200_segment_center_f32(const float* x, int dim, const int64_t* seg_off, int n_seg, const float* mean,
                                          float* out, double* scratch, double* sq, void* stream) {
    OODB200_REQUIRE(dim > 0 && n_seg >= 0 && n_seg <= 65535, "segment_center: bad size");
    if (n_seg == 0) return OODB200_OK;
    OODB200_REQUIRE(x && seg_off && mean && out && scratch && sq, "segment_center: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(kCsBlocks, (unsigned)n_seg);
    if (dim % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)mean & 15) == 0)
        seg_center_kernel<true><<<grid, kCsThreads, 0, st>>>(x, dim, seg_off, mean, out, scratch);
    else
        seg_center_kernel<false><<<grid, kCsThreads, 0, st>>>(x, dim, seg_off, mean, out, scratch);
    int rc = check_launch("segment_center");
    if (rc) return rc;
    seg_sq_reduce_kernel<<<n_seg, 32, 0, st>>>(scratch, sq);
    return check_launch("segment_sq_reduce");
}
