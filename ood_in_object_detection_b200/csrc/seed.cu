// K4b: k-means++ seeding on the device, every (class, stride) segment in lock-step, sm_100a.
//
// Replaces the per-centre loop of sklearn's `_kmeans_plusplus` (sklearn/cluster/_kmeans.py:180-278), which
// `KMeans(n_clusters=k, random_state=10).fit_predict(X)` (/root/reference/cluster_utils.py:62-73) runs once per
// (class, stride).  Per new centre:
//
//   seed_scan      `np.searchsorted(np.cumsum(closest_dist_sq), rand_vals)` (:252-257).  np.cumsum on float32 is a
//                  sequential float32 sum: a single dependent FADD chain per segment (4 cycles per value: 0.5 ms for the
//                  200 000 rows of a C3 segment, and it does not shrink with the number of ranks).  The kernel (one CTA per
//                  segment) evaluates that chain in PARALLEL and bit for bit: inside one binade the running sum is an
//                  integer number of ulps and every addend contributes a fixed integer increment (its value rounded to the
//                  ulp; an exact tie goes to the even total, which makes a thread's result a function of the parity it
//                  starts from), so 512 threads each sum the increments of 128 values for both parities and a prefix scan
//                  over the composition of those parity maps gives the exact sum at every 128-value boundary; where the sum
//                  leaves the binade (~20 times per segment) one block is redone with the float chain itself and the pass
//                  restarts behind it (details at the code).  `OODB200_SEED_SCAN=serial` selects the single-thread chain
//                  (staged through shared memory, look-ahead loads pinned ahead of the FADDs).  Both record the running sum
//                  at the end of every 128-value chunk.
//                  The search is a binary search over the chunk sums followed by a re-scan of ONE chunk from its
//                  exact start value, which reproduces the sequential sums bit for bit.  rand_vals = u * pot is
//                  float64 (u ~ RandomState.uniform drawn on the host up front: the stream does not depend on the
//                  data); `cum[i] >= v` with cum float32 and v float64 is the same as `cum[i] >= float32_roundup(v)`.
//   sqdist_cand4   `_euclidean_distances(X[candidate_ids], X, squared=True)` + `np.minimum(closest, .)` (:260-265):
//                  float64 expansion cast to float32 (like `_euclidean_distances_upcast`), 4 rows per warp in
//                  registers so every candidate value read from shared memory feeds 4 rows, ONE transposing
//                  butterfly per 4 rows, per-block float64 partial potentials (no atomics: bit-reproducible).
//   seed_pots      fixed-order sum of the block partials -> candidate potentials [n_seg, n_cand] (:268).
//   seed_pick      best candidate = first minimum of the float32 potentials (:271-276); the new closest distances,
//                  the new centre, the new potential.
// Segments are described in GLOBAL row order by `pieces` so that N ranks run the identical scan on an all-gathered
// copy of the closest distances (rank-major = row order).
#include "common.cuh"

#include <float.h>

namespace oodb200 {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kScanThreads = 512;      // parallel evaluation: one 128-value block per thread and pass; serial: 15 staging warps + the chain
constexpr int kScanBatch = 4096;       // values staged per buffer
constexpr int kScanChunk = 128;        // running sum recorded every kScanChunk values
constexpr int kMaxPieces = 16;         // ranks
constexpr int kScanPitch = 65;         // floats per block row of a warp's tile in the parallel evaluation (64 values + 1)

struct ScanParams {
    const float* closest_all;          // all-gathered closest distances
    const int64_t* piece_off;          // [n_seg, n_pieces] element offset of the piece in closest_all
    const int64_t* piece_cnt;          // [n_seg, n_pieces] rows of the piece
    int n_pieces;
    const double* uniform;             // [n_seg, n_trials] u in [0, 1)
    const float* pot;                  // [n_seg] current potential (float32)
    const int32_t* seg_trials;         // [n_seg] trials used by the segment (<= n_trials)
    const int32_t* seg_on;             // [n_seg] 0 -> segment skipped this round
    int n_trials;
    float* chunk_sum;                  // [n_seg, max_chunks] scratch
    int64_t max_chunks;
    int64_t* cand_id;                  // [n_seg, n_trials] global row index within the segment
    int serial;                        // 1: the single-thread FADD chain (A/B and fallback), 0: the parallel exact evaluation
};

// 32 consecutive floats from shared memory / the sequential float32 sum over them (program order pinned: see seed_scan_kernel)
__device__ __forceinline__ void scan_ld32(float (&v)[32], uint32_t addr) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[4 * q]), "=f"(v[4 * q + 1]), "=f"(v[4 * q + 2]), "=f"(v[4 * q + 3])
                     : "r"(addr + 16u * q)
                     : "memory");
}
__device__ __forceinline__ float scan_chain32(float run, const float (&v)[32]) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        asm volatile("add.rn.f32 %0, %0, %1;\n add.rn.f32 %0, %0, %2;\n add.rn.f32 %0, %0, %3;\n add.rn.f32 %0, %0, %4;\n"
                     " add.rn.f32 %0, %0, %5;\n add.rn.f32 %0, %0, %6;\n add.rn.f32 %0, %0, %7;\n add.rn.f32 %0, %0, %8;"
                     : "+f"(run)
                     : "f"(v[8 * q]), "f"(v[8 * q + 1]), "f"(v[8 * q + 2]), "f"(v[8 * q + 3]), "f"(v[8 * q + 4]), "f"(v[8 * q + 5]),
                       "f"(v[8 * q + 6]), "f"(v[8 * q + 7]));
    return run;
}

__global__ void __launch_bounds__(kScanThreads) seed_scan_kernel(const ScanParams p) {
    __shared__ __align__(16) float s_buf[2][kScanBatch + 32];   // + one group: the look-ahead load of the chain needs no bounds test
    extern __shared__ float s_tile[];                  // [warps][32][kScanPitch], parallel evaluation only
    __shared__ int64_t s_start[kMaxPieces + 1];        // global index of the first row of every piece
    __shared__ int64_t s_off[kMaxPieces];
    const int g = blockIdx.x, tid = threadIdx.x;
    if (p.seg_on && !p.seg_on[g]) return;
    if (tid == 0) {
        int64_t acc = 0;
        for (int r = 0; r < p.n_pieces; ++r) {
            s_start[r] = acc;
            s_off[r] = p.piece_off[(size_t)g * p.n_pieces + r];
            acc += p.piece_cnt[(size_t)g * p.n_pieces + r];
        }
        s_start[p.n_pieces] = acc;
    }
    __syncthreads();
    const int np = p.n_pieces;
    const int64_t n = s_start[np];
    if (n == 0) return;
    auto value = [&](int64_t i) -> float {             // closest distance of global row i of this segment
        int r = 0;
        while (r + 1 < np && i >= s_start[r + 1]) ++r;
        return __ldg(p.closest_all + s_off[r] + (i - s_start[r]));
    };
    auto stage = [&](int64_t i0, int b) {
        for (int j = tid; j < kScanBatch; j += kScanThreads) {
            const int64_t i = i0 + j;
            s_buf[b][j] = i < n ? value(i) : 0.f;
        }
    };
    float* __restrict__ cs = p.chunk_sum + (size_t)g * p.max_chunks;
    const int64_t n_blocks = (n + kScanChunk - 1) / kScanChunk;
    if (!p.serial) {
        // ---- the sequential float32 sum at every 128-value boundary, evaluated in parallel and bit for bit ----
        // While the running sum s stays inside one binade [2^e, 2^(e+1)) it is an integer multiple S of u = 2^(e-23), and
        // RN(s + v) = (S + rn(v / u)) * u: the increment rn(v / u) does not depend on s, except that an exact tie (fraction
        // 0.5) rounds to the side that makes S even.  So a thread can sum the increments of its 128 values on its own, once
        // for each parity its start value can have (T0, T1); the threads' results are combined by a prefix scan over the
        // composition "parity in -> (increment, parity out)"; every sum the float chain would have produced at a 128-value
        // boundary is then S0 + prefix, exactly.  The assumption fails where the sum leaves the binade (at most ~40 times per
        // segment, plus the start where s is tiny): the first block whose end reaches 2^24 units is redone with the plain float
        // chain from its exact start value, and the pass restarts behind it with the new unit.  The number of blocks per pass
        // grows with the blocks already summed (the next binade ends about where the sum has doubled).
        __shared__ float s_run;
        __shared__ long long s_blk, s_w0[kScanThreads / 32], s_w1[kScanThreads / 32];
        __shared__ int s_cross;
        const int lane = tid & 31, warp = tid >> 5;
        if (tid == 0) { s_run = 0.f; s_blk = 0; }
        __syncthreads();
        // the plain float chain over one block: all threads stage its values (one round trip), one thread adds them in order
        auto chain_block = [&](float s0, int64_t blk, int owner) {      // called by every thread; results through s_run / s_blk / cs
            const int64_t i0 = blk * kScanChunk;
            if (tid < kScanChunk) s_buf[0][tid] = i0 + tid < n ? value(i0 + tid) : 0.f;   // + 0.0f leaves a sum >= 0 unchanged
            __syncthreads();
            if (tid == owner) {
                const float4* __restrict__ v4 = reinterpret_cast<const float4*>(s_buf[0]);
#pragma unroll 8
                for (int q = 0; q < kScanChunk / 4; ++q) {
                    const float4 v = v4[q];
                    s0 = __fadd_rn(s0, v.x); s0 = __fadd_rn(s0, v.y); s0 = __fadd_rn(s0, v.z); s0 = __fadd_rn(s0, v.w);
                }
                cs[blk] = s0;
                s_run = s0;
                s_blk = blk + 1;
            }
        };
        for (;;) {
            const int64_t blk0 = s_blk;
            const float s = s_run;
            if (blk0 >= n_blocks) break;
            const int eb = (__float_as_int(s) >> 23) & 0xff;           // biased exponent (s >= 0)
            if (eb < 27 || eb > 250) {                                   // zero / tiny / huge: no unit to count in, take the block serially
                __syncthreads();
                chain_block(s, blk0, 0);
                __syncthreads();
                continue;
            }
            const float scale = __int_as_float((277 - eb) << 23);       // 2^(23 - e): v * scale = v / u, exact
            const float u = __int_as_float((eb - 23) << 23);            // 2^(e - 23)
            const long long Sint = (long long)(s * scale);              // in [2^23, 2^24)
            long long lim = 2 * blk0 + 4;                                // blocks this pass (see above)
            if (lim > kScanThreads) lim = kScanThreads;
            const int nb = (int)min((long long)(n_blocks - blk0), lim);
            unsigned t0 = 0, t1 = 0;                                     // increments for start parity 0 / 1 (<= 128 * 2^24 < 2^32)
            if (warp * 32 < nb) {                                        // warp-uniform: this warp's 32 blocks = 4096 consecutive values
                // The warp loads them with coalesced requests into its own shared-memory tile (a lane reading its own block
                // straight from global memory touches 32 different lines per request) and every lane then walks its block there
                // (row pitch 65 floats: conflict-free).  Two halves of 64 values per block keep the tile at 8.3 KB per warp.
                float* __restrict__ tile = s_tile + (size_t)warp * (32 * kScanPitch);
                const int64_t w0 = (blk0 + warp * 32) * kScanChunk;
                int pr = 0;                                              // the piece (rank) the warp's range starts in
                while (pr + 1 < np && w0 >= s_start[pr + 1]) ++pr;
                const bool inside = w0 + 32 * kScanChunk <= s_start[pr + 1];       // ... and ends in: one contiguous source
                const float* __restrict__ src = p.closest_all + s_off[pr] + (w0 - s_start[pr]);
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    if (inside) {                                        // the common case, free of branches: the loads batch
#pragma unroll 1
                        for (int e0 = 0; e0 < 32 * 64; e0 += 32 * 32) {
                            float vv[32];
#pragma unroll
                            for (int q = 0; q < 32; ++q) {               // element e: block e / 64, value h * 64 + e % 64
                                const int e = e0 + q * 32 + lane;
                                vv[q] = __ldg(src + (e >> 6) * kScanChunk + h * 64 + (e & 63));
                            }
#pragma unroll
                            for (int q = 0; q < 32; ++q) {
                                const int e = e0 + q * 32 + lane;
                                tile[(e >> 6) * kScanPitch + (e & 63)] = vv[q];
                            }
                        }
                    } else {
                        for (int e = lane; e < 32 * 64; e += 32) {
                            const int b = e >> 6, j = e & 63;
                            const int64_t i = w0 + b * kScanChunk + h * 64 + j;
                            tile[b * kScanPitch + j] = i < n ? value(i) : 0.f;
                        }
                    }
                    __syncwarp();
                    if (tid < nb) {
                        const float* __restrict__ mine = tile + lane * kScanPitch;
#pragma unroll 8
                        for (int j = 0; j < 64; ++j) {
                            float x = mine[j] * scale;
                            if (!(x < 16777216.f)) x = 16777216.f;      // leaves the binade anyway (also inf / nan)
                            const float fl = floorf(x), fr = x - fl;    // exact
                            const unsigned qi = (unsigned)fl;
                            if (fr == 0.5f) {                            // tie: to the even total
                                t0 += qi + ((t0 + qi) & 1u);
                                t1 += qi + ((1u + t1 + qi) & 1u);
                            } else {
                                const unsigned up = fr > 0.5f ? 1u : 0u;
                                t0 += qi + up;
                                t1 += qi + up;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            // inclusive scan of the composition inside the warp: (b then a)(p) = b_p + a_{p ^ (b_p & 1)}
            long long a0 = t0, a1 = t1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long b0 = __shfl_up_sync(kFull, a0, o), b1 = __shfl_up_sync(kFull, a1, o);
                if (lane >= o) {
                    const long long n0 = b0 + ((b0 & 1) ? a1 : a0), n1 = b1 + ((b1 & 1) ? a0 : a1);
                    a0 = n0;
                    a1 = n1;
                }
            }
            const long long up0 = __shfl_up_sync(kFull, a0, 1), up1 = __shfl_up_sync(kFull, a1, 1);
            if (lane == 31) { s_w0[warp] = a0; s_w1[warp] = a1; }
            if (tid == 0) s_cross = nb;
            __syncthreads();
            long long acc = 0;
            int par = (int)(Sint & 1);
            for (int w = 0; w < warp; ++w) {
                const long long t = par ? s_w1[w] : s_w0[w];
                acc += t;
                par ^= (int)(t & 1);
            }
            const long long incl = acc + (par ? a1 : a0);               // units added up to and including this thread's block
            const long long excl = acc + (lane ? (par ? up1 : up0) : 0);
            const bool crossing = tid < nb && Sint + incl >= 16777216ll;
            if (crossing) atomicMin(&s_cross, tid);
            __syncthreads();
            const int tc = s_cross;
            if (tid < nb && tid < tc) cs[blk0 + tid] = (float)(Sint + incl) * u;   // < 2^24 units: exact
            if (tc < nb) {
                chain_block((float)(Sint + excl) * u, blk0 + tc, tc);     // only thread tc's start value is used
            } else if (tid == nb - 1) {
                s_run = (float)(Sint + incl) * u;
                s_blk = blk0 + nb;
            }
            __syncthreads();
        }
    }
    const int64_t n_batches = p.serial ? (n + kScanBatch - 1) / kScanBatch : 0;
    if (p.serial) stage(0, 0);
    __syncthreads();
    float run = 0.f;                                   // thread 0: the sequential float32 sum
    for (int64_t b = 0; b < n_batches; ++b) {
        const int cur = (int)(b & 1);
        if (tid == 0) {
            const int64_t i0 = b * kScanBatch;
            const int valid = (int)min((int64_t)kScanBatch, n - i0);
            const int n_chunks = (valid + kScanChunk - 1) / kScanChunk;
            // The FADD chain (4 cycles per value) must never wait for shared memory: two register sets of 32 values, the loads
            // of one set issued before the chain over the other (volatile asm keeps that order through the compiler).
            const uint32_t base = (uint32_t)__cvta_generic_to_shared(s_buf[cur]);
            const int n_groups = n_chunks * (kScanChunk / 32);                 // even; padding values are +0.0f: x + 0 == x for x >= 0
            float va[32], vb[32];
            scan_ld32(va, base);
            for (int gq = 0; gq < n_groups; gq += 2) {
                scan_ld32(vb, base + (uint32_t)(gq + 1) * 128u);
                run = scan_chain32(run, va);
                scan_ld32(va, base + (uint32_t)(gq + 2) * 128u);               // beyond the last group: the pad, never added
                run = scan_chain32(run, vb);
                if ((gq & 3) == 2) cs[b * (kScanBatch / kScanChunk) + (gq >> 2)] = run;
            }
        } else if (tid >= 32 && b + 1 < n_batches) {
            // warps 1.. stage the next batch meanwhile; the other lanes of warp 0 stay idle: a divergent warp would
            // interleave their global loads with the FADD chain of lane 0
            const int64_t i0 = (b + 1) * kScanBatch;
            constexpr int kStagers = kScanThreads - 32;
            for (int j0 = tid - 32; j0 < kScanBatch; j0 += 4 * kStagers) {     // 4 independent loads in flight per thread
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u * kStagers;
                    v[u] = (j < kScanBatch && i0 + j < n) ? value(i0 + j) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u * kStagers;
                    if (j < kScanBatch) s_buf[cur ^ 1][j] = v[u];
                }
            }
        }
        __syncthreads();
    }
    // ---- searchsorted(cumsum, u * pot), side='left', then clip to n - 1
    const int64_t n_chunks_tot = (n + kScanChunk - 1) / kScanChunk;
    const int trials = p.seg_trials ? p.seg_trials[g] : p.n_trials;
    __shared__ int64_t s_lo[32];                       // per trial: first chunk whose end sum reaches the threshold
    __shared__ float s_thr[32];
    if (tid < trials) {
        const double v = p.uniform[(size_t)g * p.n_trials + tid] * (double)p.pot[g];
        const float thr = __double2float_ru(v);        // cum >= v  <=>  cum >= thr for float32 cum
        int64_t lo = 0, hi = n_chunks_tot;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (cs[mid] >= thr) hi = mid; else lo = mid + 1;
        }
        s_lo[tid] = lo;
        s_thr[tid] = thr;
    }
    __syncthreads();
    // the chunk of every trial is staged by all threads (independent loads), then re-scanned from its exact start value
    for (int e = tid; e < trials * kScanChunk; e += kScanThreads) {
        const int t = e / kScanChunk;
        const int64_t i = s_lo[t] * kScanChunk + (e - t * kScanChunk);
        s_buf[0][e] = (s_lo[t] < n_chunks_tot && i < n) ? value(i) : 0.f;
    }
    __syncthreads();
    if (tid < p.n_trials) {
        int64_t id = 0;
        if (tid < trials) {
            const int64_t lo = s_lo[tid];
            const float thr = s_thr[tid];
            if (lo >= n_chunks_tot) {
                id = n - 1;                            // beyond the total: np.clip(ids, None, n - 1)
            } else {
                float s = lo > 0 ? cs[lo - 1] : 0.f;
                const int64_t i0 = lo * kScanChunk;
                const int cnt = (int)min((int64_t)kScanChunk, n - i0);
                int hit = cnt - 1;
                bool found = false;
                const float* __restrict__ vals = s_buf[0] + tid * kScanChunk;
#pragma unroll 8
                for (int q = 0; q < kScanChunk; ++q) {                         // no early exit: the loads pipeline
                    s = __fadd_rn(s, vals[q]);
                    if (!found && q < cnt && s >= thr) { hit = q; found = true; }
                }
                id = i0 + hit;
            }
        }
        p.cand_id[(size_t)g * p.n_trials + tid] = id;
    }
    __syncthreads();
    if (tid == 0 && trials < p.n_trials) {             // padding slots repeat the first candidate (ignored by pick)
        const int64_t id0 = p.cand_id[(size_t)g * p.n_trials];
        for (int t = trials; t < p.n_trials; ++t) p.cand_id[(size_t)g * p.n_trials + t] = id0;
    }
}

// ---- candidate distances, 4 rows per warp -----------------------------------------------------------------------
struct Cand4Params {
    const float* x;
    int dim;
    const int64_t* seg_off;            // [n_seg + 1] local row offsets
    int n_seg;
    const float* cand;                 // [n_seg, n_cand, dim]
    int n_cand;                        // <= 4
    const float* closest;              // [n_rows] or null
    float* out_d;                      // [n_cand, n_rows]
    double* pot_part;                  // [n_seg, gridDim.x, 4]
};

__device__ __forceinline__ double shfl_xor_f64(double v, int o) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(kFull, lo, o);
    hi = __shfl_xor_sync(kFull, hi, o);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double transpose_reduce_f64(double (&v)[32], int lane) {   // lane l ends with the total of v[l]
    int o = 16;
#pragma unroll
    for (int n = 32; n > 1; n >>= 1, o >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const double send = up ? v[i] : v[i + n / 2];
            const double keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + shfl_xor_f64(send, o);
        }
    }
    return v[0];
}

__device__ __forceinline__ uint32_t sd_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sd_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sd_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sd_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sd_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sd_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sd_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(sd_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void sd_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(sd_smem_u32(bar)), "r"(parity)
                     : "memory");
    }
}

constexpr int kC4Rows = 32;                                     // rows per TMA chunk: 8 warps x 4 rows

// One CTA owns a contiguous range of 32-row chunks of its segment.  The rows arrive by 1-D TMA bulk copies into a
// double-buffered landing zone (one copy = one chunk = 72 KB at D = 576), so 72-144 KB are in flight per SM while the
// warps convert and multiply the previous chunk from shared memory; registers only hold the 32 float64 accumulators.
template <int NJ>
__global__ void __launch_bounds__(256, 1) sqdist_cand4_kernel(const Cand4Params p) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int g = blockIdx.y;
    const int D = p.dim, NC = p.n_cand;
    float* s_stage = reinterpret_cast<float*>(s_raw);                              // [2][32][D]
    double* s_y = reinterpret_cast<double*>(s_raw + sizeof(float) * 2 * kC4Rows * D);   // [4][D]
    double* s_yy = s_y + (size_t)4 * D;
    __shared__ __align__(8) uint64_t s_full[2];
    __shared__ double s_pot[8][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t r0 = p.seg_off[g], r1 = p.seg_off[g + 1];
    const int64_t n_rows = p.seg_off[p.n_seg];
    const int64_t seg_chunks = (r1 - r0 + kC4Rows - 1) / kC4Rows;
    const int64_t per_block = (seg_chunks + gridDim.x - 1) / gridDim.x;
    const int64_t c_first = (int64_t)blockIdx.x * per_block;
    const int n_chunks = (int)max((int64_t)0, min(per_block, seg_chunks - c_first));
    const int64_t b0 = r0 + c_first * kC4Rows;                                     // first row of this block
    auto issue = [&](int c) {
        const int64_t ra = b0 + (int64_t)c * kC4Rows;
        const uint32_t bytes = (uint32_t)(min((int64_t)kC4Rows, r1 - ra) * D * 4);
        sd_mbar_expect_tx(&s_full[c & 1], bytes);
        sd_bulk_g2s(s_stage + (size_t)(c & 1) * kC4Rows * D, p.x + ra * D, bytes, &s_full[c & 1]);
    };
    if (tid == 0) {
        sd_mbar_init(&s_full[0], 1);
        sd_mbar_init(&s_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (n_chunks > 0) issue(0);
        if (n_chunks > 1) issue(1);
    }
    for (int i = tid; i < 4 * D; i += blockDim.x) {
        const int j = i / D, d = i - j * D;
        s_y[i] = j < NC ? (double)p.cand[((size_t)g * NC + j) * D + d] : 0.0;
    }
    __syncthreads();
    if (warp < 4) {
        double s = 0.0;
        for (int d = lane; d < D; d += 32) s += s_y[warp * D + d] * s_y[warp * D + d];
        for (int o = 16; o > 0; o >>= 1) s += shfl_xor_f64(s, o);
        if (lane == 0) s_yy[warp] = s;
    }
    __syncthreads();
    const int q = lane & 7, rr = lane >> 3;                    // after the butterfly: lane = row rr, quantity q
    double pot = 0.0;
    for (int c = 0; c < n_chunks; ++c) {
        const float* __restrict__ st = s_stage + (size_t)(c & 1) * kC4Rows * D + (size_t)warp * 4 * D;
        const int64_t base = b0 + (int64_t)c * kC4Rows + warp * 4;
        sd_mbar_wait(&s_full[c & 1], (uint32_t)(c >> 1) & 1u);
        double acc[32];                                        // [row][8]: 4 candidate dots, ||x||^2, 3 unused
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int d = lane * 4 + 128 * j;
            if (d < D) {
                double2 y01[4], y23[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    y01[cc] = *reinterpret_cast<const double2*>(s_y + (size_t)cc * D + d);
                    y23[cc] = *reinterpret_cast<const double2*>(s_y + (size_t)cc * D + d + 2);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    // rows past the end of the segment were not copied: stale data, results discarded below
                    const float4 xv = *reinterpret_cast<const float4*>(st + (size_t)r * D + d);
                    const double a = (double)xv.x, b = (double)xv.y, c2 = (double)xv.z, e = (double)xv.w;
                    double xx = acc[r * 8 + 4];
                    xx += a * a; xx += b * b; xx += c2 * c2; xx += e * e;
                    acc[r * 8 + 4] = xx;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        double s = acc[r * 8 + cc];
                        s += a * y01[cc].x; s += b * y01[cc].y; s += c2 * y23[cc].x; s += e * y23[cc].y;
                        acc[r * 8 + cc] = s;
                    }
                }
            }
        }
        const double tot = transpose_reduce_f64(acc, lane);    // lane: row rr, quantity q
        int xlo = __double2loint(tot), xhi = __double2hiint(tot);
        xlo = __shfl_sync(kFull, xlo, (lane & ~7) | 4);
        xhi = __shfl_sync(kFull, xhi, (lane & ~7) | 4);
        const double xnorm = __hiloint2double(xhi, xlo);
        const int64_t r = base + rr;
        if (q < NC && r < r1) {
            const float cl = p.closest ? p.closest[r] : FLT_MAX;
            float d32 = (float)(-2.0 * tot + s_yy[q] + xnorm);
            d32 = fminf(fmaxf(d32, 0.f), cl);
            p.out_d[(size_t)q * n_rows + r] = d32;
            pot += (double)d32;
        }
        __syncthreads();                                       // the buffer is free again
        if (tid == 0 && c + 2 < n_chunks) issue(c + 2);
    }
    // block partial in a fixed order: rows of a lane group, then warps
    pot += shfl_xor_f64(pot, 8);
    pot += shfl_xor_f64(pot, 16);
    if (lane < 4) s_pot[warp][lane] = pot;
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_pot[w][threadIdx.x];
        p.pot_part[((size_t)g * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = t;
    }
}

// generic shapes (dim % 4 != 0, unaligned x): one row per warp, same arithmetic and the same partial layout
__global__ void __launch_bounds__(256) sqdist_cand1_kernel(const Cand4Params p) {
    extern __shared__ __align__(16) double s_y[];
    const int g = blockIdx.y;
    const int D = p.dim, NC = p.n_cand;
    double* s_yy = s_y + (size_t)4 * D;
    __shared__ double s_pot[8][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = p.seg_off[g], r1 = p.seg_off[g + 1];
    const int64_t n_rows = p.seg_off[p.n_seg];
    for (int i = threadIdx.x; i < 4 * D; i += blockDim.x) {
        const int j = i / D, d = i - j * D;
        s_y[i] = j < NC ? (double)p.cand[((size_t)g * NC + j) * D + d] : 0.0;
    }
    __syncthreads();
    if (warp < 4) {
        double s = 0.0;
        for (int d = lane; d < D; d += 32) s += s_y[warp * D + d] * s_y[warp * D + d];
        for (int o = 16; o > 0; o >>= 1) s += shfl_xor_f64(s, o);
        if (lane == 0) s_yy[warp] = s;
    }
    __syncthreads();
    double pot[4] = {0, 0, 0, 0};
    for (int64_t r = r0 + (int64_t)blockIdx.x * 8 + warp; r < r1; r += (int64_t)gridDim.x * 8) {
        const float* __restrict__ xr = p.x + r * D;
        double xx = 0.0, dot[4] = {0, 0, 0, 0};
        for (int d = lane; d < D; d += 32) {
            const double v = (double)__ldg(xr + d);
            xx += v * v;
#pragma unroll
            for (int j = 0; j < 4; ++j) dot[j] += v * s_y[j * D + d];
        }
        for (int o = 16; o > 0; o >>= 1) {
            xx += shfl_xor_f64(xx, o);
#pragma unroll
            for (int j = 0; j < 4; ++j) dot[j] += shfl_xor_f64(dot[j], o);
        }
        if (lane == 0) {
            const float cl = p.closest ? p.closest[r] : FLT_MAX;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < NC) {
                    float d32 = (float)(-2.0 * dot[j] + s_yy[j] + xx);
                    d32 = fminf(fmaxf(d32, 0.f), cl);
                    p.out_d[(size_t)j * n_rows + r] = d32;
                    pot[j] += (double)d32;
                }
        }
    }
    if (lane == 0)
        for (int j = 0; j < 4; ++j) s_pot[warp][j] = pot[j];
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_pot[w][threadIdx.x];
        p.pot_part[((size_t)g * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = t;
    }
}

// pots[g, j] = sum_b pot_part[g, b, j] in increasing b
__global__ void seed_pots_kernel(const double* __restrict__ part, int n_blocks, int n_cand, double* __restrict__ pots) {
    const int g = blockIdx.x, j = threadIdx.x;
    if (j >= n_cand) return;
    double t = 0.0;
    for (int b = 0; b < n_blocks; ++b) t += part[((size_t)g * n_blocks + b) * 4 + j];
    pots[(size_t)g * n_cand + j] = t;
}

struct PickParams {
    const double* pots;                // [n_seg, n_cand] (summed over ranks)
    const int32_t* seg_trials;
    const int32_t* seg_on;
    int n_seg, n_cand;
    const int64_t* seg_off;            // local rows
    const float* newd;                 // [n_cand, n_rows]
    const float* cand_vec;             // [n_seg, n_cand, dim]
    int dim;
    float* closest;                    // [n_rows] updated in place
    float* pot;                        // [n_seg]
    float* cent_out;                   // centre slot of this round: cent + c * dim, segment stride cent_stride
    int64_t cent_stride;
    int32_t* best_out;                 // [n_seg]
};

__device__ __forceinline__ int pick_best(const PickParams& p, int g) {
    const int t = p.seg_trials ? p.seg_trials[g] : p.n_cand;
    int best = 0;
    float bp = (float)p.pots[(size_t)g * p.n_cand];
    for (int j = 1; j < t; ++j) {                      // np.argmin over float32 potentials: first minimum
        const float v = (float)p.pots[(size_t)g * p.n_cand + j];
        if (v < bp) { bp = v; best = j; }
    }
    return best;
}

__global__ void __launch_bounds__(256) seed_pick_kernel(const PickParams p) {
    const int g = blockIdx.y;
    if (p.seg_on && !p.seg_on[g]) return;
    const int best = pick_best(p, g);
    const int64_t r0 = p.seg_off[g], r1 = p.seg_off[g + 1];
    const int64_t n_rows = p.seg_off[p.n_seg];
    const float* __restrict__ src = p.newd + (size_t)best * n_rows;
    for (int64_t r = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += (int64_t)gridDim.x * blockDim.x)
        p.closest[r] = src[r];
    if (blockIdx.x == 0) {
        for (int d = threadIdx.x; d < p.dim; d += blockDim.x)
            p.cent_out[(size_t)g * p.cent_stride + d] = p.cand_vec[((size_t)g * p.n_cand + best) * p.dim + d];
        if (threadIdx.x == 0) {
            p.pot[g] = (float)p.pots[(size_t)g * p.n_cand + best];
            if (p.best_out) p.best_out[g] = best;
        }
    }
}

// vec[g, j, :] = x[local row of global id] if this rank owns the row, else 0 (ranks are summed afterwards)
__global__ void seed_gather_kernel(const float* __restrict__ x, int dim, const int64_t* __restrict__ cand_id, int n_cand,
                                   const int64_t* __restrict__ seg_off, const int64_t* __restrict__ shard_first,
                                   float* __restrict__ vec) {
    const int g = blockIdx.x, j = blockIdx.y;
    const int64_t id = cand_id[(size_t)g * n_cand + j] - shard_first[g];
    const int64_t cnt = seg_off[g + 1] - seg_off[g];
    const bool mine = id >= 0 && id < cnt;
    float* __restrict__ o = vec + ((size_t)g * n_cand + j) * dim;
    const float* __restrict__ s = x + (size_t)(seg_off[g] + (mine ? id : 0)) * dim;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) o[d] = mine ? s[d] : 0.f;
}

}  // namespace oodb200

using namespace oodb200;

extern "C" int oodb200_seed_grid(int64_t max_seg_rows) {
    long long gx = (max_seg_rows + 255) / 256;
    if (gx > 148) gx = 148;
    if (gx < 1) gx = 1;
    return (int)gx;
}

extern "C" int oodb200_seed_sqdist_f32(const float* x, int dim, const int64_t* seg_off, int n_seg, int64_t max_seg_rows,
                                       const float* cand, int n_cand, const float* closest, float* out_d,
                                       double* pot_part, double* pots, void* stream) {
    OODB200_REQUIRE(dim > 0 && n_seg >= 0 && n_cand >= 1 && n_cand <= 4, "seed_sqdist: bad size (n_cand <= 4)");
    if (n_seg == 0) return OODB200_OK;
    OODB200_REQUIRE(x && seg_off && cand && out_d && pot_part && pots, "seed_sqdist: null pointer");
    OODB200_REQUIRE(n_seg <= 65535, "seed_sqdist: too many segments");
    const size_t smem = sizeof(double) * ((size_t)4 * dim + 4);
    OODB200_REQUIRE(smem <= 200 * 1024, "seed_sqdist: dim too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = oodb200_seed_grid(max_seg_rows);
    Cand4Params p = {x, dim, seg_off, n_seg, cand, n_cand, closest, out_d, pot_part};
    const int nj = (dim + 127) / 128;
    const dim3 grid((unsigned)gx, (unsigned)n_seg);
    cudaError_t e = cudaSuccess;
    const size_t fsmem = sizeof(float) * 2 * kC4Rows * (size_t)dim + smem;
    if (dim % 4 == 0 && nj <= 6 && ((uintptr_t)x & 15) == 0 && fsmem <= 220 * 1024) {
#define OODB200_C4_LAUNCH(NJ)                                                                                         \
    case NJ:                                                                                                          \
        e = cudaFuncSetAttribute(sqdist_cand4_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);   \
        if (e == cudaSuccess) sqdist_cand4_kernel<NJ><<<grid, 256, fsmem, st>>>(p);                                   \
        break;
        switch (nj) {
            OODB200_C4_LAUNCH(1) OODB200_C4_LAUNCH(2) OODB200_C4_LAUNCH(3) OODB200_C4_LAUNCH(4)
            OODB200_C4_LAUNCH(5) OODB200_C4_LAUNCH(6)
        }
#undef OODB200_C4_LAUNCH
    } else {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(sqdist_cand1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) sqdist_cand1_kernel<<<grid, 256, smem, st>>>(p);
    }
    if (e != cudaSuccess) { set_error("seed_sqdist: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    int rc = check_launch("seed_sqdist");
    if (rc) return rc;
    seed_pots_kernel<<<n_seg, 32, 0, st>>>(pot_part, gx, n_cand, pots);
    return check_launch("seed_pots");
}

extern "C" int oodb200_seed_scan_f32(const float* closest_all, const int64_t* piece_off, const int64_t* piece_cnt,
                                     int n_seg, int n_pieces, const double* uniform, const float* pot,
                                     const int32_t* seg_trials, const int32_t* seg_on, int n_trials, float* chunk_sum,
                                     int64_t max_chunks, int64_t* cand_id, void* stream) {
    OODB200_REQUIRE(n_seg >= 0 && n_pieces >= 1 && n_pieces <= kMaxPieces, "seed_scan: 1..%d pieces per segment", kMaxPieces);
    OODB200_REQUIRE(n_trials >= 1 && n_trials <= 32, "seed_scan: 1..32 trials");
    if (n_seg == 0) return OODB200_OK;
    OODB200_REQUIRE(closest_all && piece_off && piece_cnt && uniform && pot && chunk_sum && cand_id, "seed_scan: null pointer");
    static int serial = -1;                                         // OODB200_SEED_SCAN=serial: the single-thread chain (A/B runs)
    if (serial < 0) { const char* env = getenv("OODB200_SEED_SCAN"); serial = (env && env[0] == 's') ? 1 : 0; }
    ScanParams p = {closest_all, piece_off, piece_cnt, n_pieces, uniform, pot, seg_trials, seg_on, n_trials, chunk_sum,
                    max_chunks, cand_id, serial};
    const size_t tile_bytes = serial ? 0 : sizeof(float) * (kScanThreads / 32) * 32 * kScanPitch;
    if (tile_bytes && cudaFuncSetAttribute(seed_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes) != cudaSuccess) {
        set_error("seed_scan: cannot reserve %zu bytes of shared memory", tile_bytes);
        return OODB200_ERR_CUDA;
    }
    seed_scan_kernel<<<n_seg, kScanThreads, tile_bytes, (cudaStream_t)stream>>>(p);
    return check_launch("seed_scan");
}

extern "C" int oodb200_seed_gather_f32(const float* x, int dim, const int64_t* cand_id, int n_seg, int n_cand,
                                       const int64_t* seg_off, const int64_t* shard_first, float* vec, void* stream) {
    OODB200_REQUIRE(dim > 0 && n_seg >= 0 && n_cand >= 1 && n_cand <= 65535, "seed_gather: bad size");
    if (n_seg == 0) return OODB200_OK;
    OODB200_REQUIRE(x && cand_id && seg_off && shard_first && vec, "seed_gather: null pointer");
    seed_gather_kernel<<<dim3((unsigned)n_seg, (unsigned)n_cand), 128, 0, (cudaStream_t)stream>>>(x, dim, cand_id, n_cand, seg_off,
                                                                                                   shard_first, vec);
    return check_launch("seed_gather");
}

extern "C" int oodb200_seed_pick_f32(const double* pots, const int32_t* seg_trials, const int32_t* seg_on, int n_seg,
                                     int n_cand, const int64_t* seg_off, int64_t max_seg_rows, const float* newd,
                                     const float* cand_vec, int dim, float* closest, float* pot, float* cent_out,
                                     int64_t cent_stride, int32_t* best_out, void* stream) {
    OODB200_REQUIRE(n_seg >= 0 && n_cand >= 1 && dim > 0, "seed_pick: bad size");
    if (n_seg == 0) return OODB200_OK;
    OODB200_REQUIRE(pots && seg_off && newd && cand_vec && closest && pot && cent_out, "seed_pick: null pointer");
    PickParams p = {pots, seg_trials, seg_on, n_seg, n_cand, seg_off, newd, cand_vec, dim, closest, pot, cent_out, cent_stride,
                    best_out};
    long long gx = (max_seg_rows + 1023) / 1024;
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    seed_pick_kernel<<<dim3((unsigned)gx, (unsigned)n_seg), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("seed_pick");
}
