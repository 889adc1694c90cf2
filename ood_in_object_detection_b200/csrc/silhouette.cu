// Per-cluster distance sums behind the silhouette score of the k-search, sm_100a.
//
// Replaces the O(n^2 D) part of sklearn.metrics.silhouette_score as the reference calls it while it searches the number
// of k-means clusters of one (class, stride) segment (/root/reference/cluster_utils.py:203-302, :277
// `silhouette_score(feature_maps_one_run, cluster_labels, metric=metric)`; sklearn 1.9 metrics/cluster/_unsupervised.py
// `silhouette_samples` -> `pairwise_distances_chunked` + `_silhouette_reduce`):
//
//     S[i][c] = sum over the rows j of the segment with labels[j] == c of d(x_i, x_j),   d in {l1, l2, cosine}
//
// The n x n distance matrix never exists: a CTA owns 64 rows i, walks the 64-column tiles j it is assigned, keeps the
// 64 x 64 block of partial distances in registers (4 x 4 per thread, operands staged transposed in shared memory, k-steps
// of 16 columns of D) and folds the finished block into a [64][kc] float64 table in shared memory, which it adds to S at
// the end.  d(i, i) = 0 exactly (sklearn zeroes the diagonal); cosine takes rows already normalised like sklearn's
// `normalize` and clips 1 - x.y to [0, 2].  Bound by the FP32 pipe: 2 (l1, l2) or 1 (cosine) instructions per element
// pair; the l2 / cosine cross-terms are a dense contraction that can move to tcgen05 like the Lloyd step did.
//
// The k-search scores 13 labelings of the SAME rows: `pair_dist_matrix` therefore stores the distances once (upper-triangle
// tiles computed, both halves written: every metric here is exactly symmetric in float32), and `matrix_cluster_sums` folds
// a labeling in one pass over the matrix (thread = row i, block = one cluster: the member rows j are added in ascending j,
// float64, no atomics: bit-reproducible).  The per-k cost drops from O(n^2 D) flops to n^2 * 4 bytes of HBM traffic.
#include "common.cuh"

namespace oodb200 {

constexpr int kPT = 64;                      // tile of rows / columns
constexpr int kPK = 16;                      // k-step over D
constexpr int kPThreads = 256;               // 16 x 16 threads, 4 x 4 pairs each

// STORE = false: fold every finished block into S by label.  STORE = true: gridDim = (tiles, tiles); blocks below the
// diagonal exit, the others write their block to M[i][j] and M[j][i] (labels / kc / S unused).
template <int METRIC, bool STORE>
__global__ void __launch_bounds__(kPThreads) pair_cluster_sums_kernel(const float* __restrict__ x, int n, int d, int64_t ld,
                                                                      const int32_t* __restrict__ labels, int kc,
                                                                      double* __restrict__ S, float* __restrict__ M, int64_t ldm) {
    __shared__ __align__(16) float As[kPK][kPT + 4];
    __shared__ __align__(16) float Bs[kPK][kPT + 4];
    __shared__ int s_lab[kPT];
    extern __shared__ double s_sum[];        // [kPT][kc]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.x * kPT;
    if (STORE && blockIdx.y < blockIdx.x) return;
    if (!STORE)
        for (int e = tid; e < kPT * kc; e += kPThreads) s_sum[e] = 0.0;
    const int lr = tid >> 2, lc = (tid & 3) * 4;          // loader: row of the tile, 4 consecutive columns of the k-step
    const int n_tiles = (n + kPT - 1) / kPT;
    for (int jt = blockIdx.y; jt < n_tiles; jt += STORE ? n_tiles : gridDim.y) {
        const int j0 = jt * kPT;
        __syncthreads();                                   // previous tile's labels / sums are consumed
        if (!STORE && tid < kPT) s_lab[tid] = j0 + tid < n ? labels[j0 + tid] : -1;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        for (int k0 = 0; k0 < d; k0 += kPK) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            const int ia = i0 + lr, jb = j0 + lr, kcol = k0 + lc;
            if ((d & 3) == 0 && (ld & 3) == 0) {           // rows are 16-byte aligned: one 128-bit load
                if (ia < n && kcol < d) a = __ldg(reinterpret_cast<const float4*>(x + (int64_t)ia * ld + kcol));
                if (jb < n && kcol < d) b = __ldg(reinterpret_cast<const float4*>(x + (int64_t)jb * ld + kcol));
            } else {
                float ta[4] = {0.f, 0.f, 0.f, 0.f}, tb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (ia < n && kcol + q < d) ta[q] = __ldg(x + (int64_t)ia * ld + kcol + q);
                    if (jb < n && kcol + q < d) tb[q] = __ldg(x + (int64_t)jb * ld + kcol + q);
                }
                a = make_float4(ta[0], ta[1], ta[2], ta[3]);
                b = make_float4(tb[0], tb[1], tb[2], tb[3]);
            }
            __syncthreads();                               // the previous k-step has been read
            As[lc + 0][lr] = a.x; As[lc + 1][lr] = a.y; As[lc + 2][lr] = a.z; As[lc + 3][lr] = a.w;
            Bs[lc + 0][lr] = b.x; Bs[lc + 1][lr] = b.y; Bs[lc + 2][lr] = b.z; Bs[lc + 3][lr] = b.w;
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kPK; ++kk) {
                const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (METRIC == OODB200_METRIC_L1) {
                            acc[r][c] += fabsf(ar[r] - br[c]);
                        } else if (METRIC == OODB200_METRIC_L2) {
                            const float e = ar[r] - br[c];
                            acc[r][c] = fmaf(e, e, acc[r][c]);
                        } else {
                            acc[r][c] = fmaf(ar[r], br[c], acc[r][c]);
                        }
                    }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + ty * 4 + r;
            if (i >= n) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int jl = tx * 4 + c, j = j0 + jl;
                if (j >= n) continue;
                float v = acc[r][c];
                if (METRIC == OODB200_METRIC_L2) v = sqrtf(fmaxf(v, 0.f));
                if (METRIC == OODB200_METRIC_COS) v = fminf(fmaxf(1.0f - v, 0.f), 2.f);
                if (i == j) v = 0.f;                       // pairwise_distances(X) zeroes the diagonal
                if (STORE) {
                    M[(int64_t)i * ldm + j] = v;
                    M[(int64_t)j * ldm + i] = v;
                } else {
                    const int lab = s_lab[jl];
                    if (lab >= 0 && lab < kc) atomicAdd(&s_sum[(ty * 4 + r) * kc + lab], (double)v);
                }
            }
        }
    }
    if (STORE) return;
    __syncthreads();
    for (int e = tid; e < kPT * kc; e += kPThreads) {
        const int i = i0 + e / kc;
        if (i < n) {
            if (gridDim.y == 1) S[(int64_t)i * kc + e % kc] = s_sum[e];
            else atomicAdd(&S[(int64_t)i * kc + e % kc], s_sum[e]);
        }
    }
}

// S[i][c] = sum over the members j of cluster c (ascending j: `order` is a stable sort of the rows by label) of M[j][i]
// (= M[i][j]); thread = row i, blockIdx.y = cluster.  Rows of M are read once overall, coalesced.
__global__ void __launch_bounds__(256) matrix_cluster_sums_kernel(const float* __restrict__ M, int n, int64_t ldm,
                                                                  const int32_t* __restrict__ order,
                                                                  const int64_t* __restrict__ member_off, int kc,
                                                                  double* __restrict__ S) {
    const int i = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
    if (i >= n) return;
    const int64_t t0 = member_off[c], t1 = member_off[c + 1];
    double acc = 0.0;
    int64_t t = t0;
    for (; t + 4 <= t1; t += 4) {                          // 4 independent loads in flight, added in order
        const float v0 = __ldg(M + (int64_t)order[t] * ldm + i), v1 = __ldg(M + (int64_t)order[t + 1] * ldm + i);
        const float v2 = __ldg(M + (int64_t)order[t + 2] * ldm + i), v3 = __ldg(M + (int64_t)order[t + 3] * ldm + i);
        acc += (double)v0; acc += (double)v1; acc += (double)v2; acc += (double)v3;
    }
    for (; t < t1; ++t) acc += (double)__ldg(M + (int64_t)order[t] * ldm + i);
    S[(int64_t)i * kc + c] = acc;
}

}  // namespace oodb200

using namespace oodb200;

extern "C" int oodb200_pair_cluster_sums_f32(const float* x, int n, int d, int64_t ld, const int32_t* labels, int kc,
                                             int metric, double* sums, void* stream) {
    OODB200_REQUIRE(n >= 0 && d > 0 && ld >= d && kc > 0, "pair_cluster_sums: bad size");
    OODB200_REQUIRE(metric >= 0 && metric < OODB200_N_METRICS, "pair_cluster_sums: metric %d", metric);
    OODB200_REQUIRE((size_t)kPT * kc * sizeof(double) <= 40 * 1024, "pair_cluster_sums: at most %d clusters", 40 * 1024 / 8 / kPT);
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(x && labels && sums, "pair_cluster_sums: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int tiles = (n + kPT - 1) / kPT;
    int split = 1;                                         // column tiles are dealt to gridDim.y CTAs per row tile when
    while (tiles * split < 2 * 148 && split < tiles) split *= 2;   // the row tiles alone cannot fill the GPU
    if (split > 1) {
        cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)n * kc, st);
        if (e != cudaSuccess) { set_error("pair_cluster_sums: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    }
    const dim3 grid(tiles, split);
    const size_t smem = sizeof(double) * kPT * kc;
    if (metric == OODB200_METRIC_L1) pair_cluster_sums_kernel<OODB200_METRIC_L1, false><<<grid, kPThreads, smem, st>>>(x, n, d, ld, labels, kc, sums, nullptr, 0);
    else if (metric == OODB200_METRIC_L2) pair_cluster_sums_kernel<OODB200_METRIC_L2, false><<<grid, kPThreads, smem, st>>>(x, n, d, ld, labels, kc, sums, nullptr, 0);
    else pair_cluster_sums_kernel<OODB200_METRIC_COS, false><<<grid, kPThreads, smem, st>>>(x, n, d, ld, labels, kc, sums, nullptr, 0);
    return check_launch("pair_cluster_sums");
}

extern "C" int oodb200_pair_dist_matrix_f32(const float* x, int n, int d, int64_t ld, int metric, float* dist, int64_t ld_dist,
                                            void* stream) {
    OODB200_REQUIRE(n >= 0 && d > 0 && ld >= d && ld_dist >= n, "pair_dist_matrix: bad size");
    OODB200_REQUIRE(metric >= 0 && metric < OODB200_N_METRICS, "pair_dist_matrix: metric %d", metric);
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(x && dist, "pair_dist_matrix: null pointer");
    const int tiles = (n + kPT - 1) / kPT;
    OODB200_REQUIRE(tiles <= 65535, "pair_dist_matrix: at most %d rows", 65535 * kPT);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(tiles, tiles);
    if (metric == OODB200_METRIC_L1) pair_cluster_sums_kernel<OODB200_METRIC_L1, true><<<grid, kPThreads, 0, st>>>(x, n, d, ld, nullptr, 0, nullptr, dist, ld_dist);
    else if (metric == OODB200_METRIC_L2) pair_cluster_sums_kernel<OODB200_METRIC_L2, true><<<grid, kPThreads, 0, st>>>(x, n, d, ld, nullptr, 0, nullptr, dist, ld_dist);
    else pair_cluster_sums_kernel<OODB200_METRIC_COS, true><<<grid, kPThreads, 0, st>>>(x, n, d, ld, nullptr, 0, nullptr, dist, ld_dist);
    return check_launch("pair_dist_matrix");
}

extern "C" int oodb200_matrix_cluster_sums_f32(const float* dist, int n, int64_t ld_dist, const int32_t* order,
                                               const int64_t* member_off, int kc, double* sums, void* stream) {
    OODB200_REQUIRE(n >= 0 && ld_dist >= n && kc > 0 && kc <= 65535, "matrix_cluster_sums: bad size");
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(dist && order && member_off && sums, "matrix_cluster_sums: null pointer");
    const dim3 grid((n + 255) / 256, kc);
    matrix_cluster_sums_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dist, n, ld_dist, order, member_off, kc, sums);
    return check_launch("matrix_cluster_sums");
}
