// Fit-stage kernels, sm_100a:
//   K2 standalone  vec_score   -- normalise + distance-min of already pooled vectors, segmented by (class, stride)
//                                 (/root/reference/ood_utils.py:2000-2036, :2404-2430)
//   K5             radix_hist  -- one pass of an exact radix select over float32 score bit patterns
//                                 (np.percentile(scores, q, method='lower'), /root/reference/ood_utils.py:613,626)
// Both are bandwidth-bound streams over N rows / N scores; see DESIGN.md for the byte accounting.
#include "common.cuh"

#include <float.h>

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
    const size_t smem = sizeof(float) * 2 * (size_t)p.d_pad * kVecWarps;
    OODB200_REQUIRE(smem <= 200 * 1024, "vec_score: dim %d too large", dim);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(vec_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("vec_score: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    }
    long long grid = (n_rows + kVecWarps - 1) / kVecWarps;
    if (grid > 148LL * 32) grid = 148LL * 32;
    vec_score_kernel<<<(int)grid, kVecThreads, smem, (cudaStream_t)stream>>>(p);
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
