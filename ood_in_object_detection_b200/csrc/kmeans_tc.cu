// K4 on the tensor pipe: Lloyd assignment with tcgen05 (kind::tf32, split-float "3xTF32") + register partial sums, sm_100a.
//
// Same contract as kmeans_step_fast_kernel (kmeans.cu): replaces sklearn's `lloyd_iter_chunked_dense` behind
// `KMeans(n_clusters=k, random_state=10).fit_predict(X)` (/root/reference/cluster_utils.py:62-73).  The FP32 kernel
// needs 2*K*D FFMA-flops per row (18.4 kflop at K = 16, D = 576) and runs at ~20 % of the HBM roofline because the
// FMA pipe, not memory, bounds it.  Here the x.c cross-term is a [128 rows x D] x [D x 16] contraction on the tensor
// core, and the CUDA cores only split, compare and accumulate.  One CTA per block of rows, 128-row tiles:
//
//   warp 9 (1 thread)   TMA producer: 2-D tensor-map copies (box 128 rows x 32 floats = one k-block, SWIZZLE_128B) into a
//                       ring (5 stages at D = 576).
//   warps 0-3           split: x_lo = tf32(x - trunc_tf32(x)) of the stage into the lo ring (generic -> async proxy
//                       fence).  The raw float32 stage is itself the "hi" operand: kind::tf32 ignores the low 13 bits.
//                       After the last k-block the same warps are the epilogue: tcgen05.ld (lane = row, warp = lane
//                       quadrant), d_k = ||c_k||^2 - 2 x.c_k, first-minimum label, changed count, per-cluster row masks.
//   warp 10 (1 thread)  MMA issuer: per k-block 4 k-steps x { x_raw.[c_hi | c_lo] (N = 32), x_lo.c_hi (N = 16) }, M = 128,
//                       accumulators in TMEM (2 x 32 columns, double buffered across tiles); tcgen05.commit frees the
//                       ring stage / publishes the tile.
//   warps 4-8           M-step of the PREVIOUS tile, overlapped with the streaming of the next one: thread = one 16-byte
//                       column chunk; the epilogue sorts the tile's rows by (label, row); the rows are re-read from L2 in
//                       that order (coalesced LDG.128 through a register ring; the tile was fetched microseconds ago) and
//                       every RUN of one cluster is added in registers between one load and one store of the cluster's
//                       slot in a thread-private [16][D] shared-memory accumulator -> block partials are bit-reproducible
//                       and identical in order to the FP32 kernel's.
// Centroids are pre-split once per iteration (kmeans_tc_prep_kernel) into the exact shared-memory image (hi and lo,
// K-major SWIZZLE_128B, 32 rows per k-block) and arrive with ONE bulk copy per CTA.
// Accuracy: x = x_hi + x_lo and c = c_hi + c_lo with 11-bit pieces; the dropped x_lo.c_lo term and the rounding of the
// lo pieces are ~2^-22 relative per product, i.e. float32-level, accumulated in FP32 in TMEM.
// History (profiles/r1_summary.md): a resident-tile variant (32-row half-tiles kept in shared memory for the M-step,
// M = 64) was correct but slower than the FP32 kernel: ~50 cycles per tcgen05.mma regardless of N, so only 128 useful
// rows per instruction amortise the issue cost, and a 2-stage lo ring made every k-block a ~550-cycle handshake.
#include "common.cuh"

#include <cuda.h>
#include <float.h>
#include <limits.h>
#include <stdlib.h>

namespace oodb200 {

constexpr int kTcRows = 128;           // rows per tile = MMA M
constexpr int kTcThreads = 352;
constexpr int kTcMaxKb = 20;           // D <= 640 (one 16-byte column chunk per M-step thread)
constexpr int kTcMaxXs = 7;            // raw ring: up to 6 x 16 KB (as many as shared memory allows: HBM latency)
constexpr int kTcMaxLs = 4;             // lo ring: up to 4 x 16 KB
constexpr int kNSplit = 4, kWAcc0 = 4, kNAcc = 5, kWTma = 9, kWMma = 10;
constexpr int kTcCols = 64;            // TMEM columns: 2 tiles x (16 hi + 16 lo)

struct TcParams {
    int dim, kb;                       // kb = dim / 32
    int k;                             // table stride (<= 16)
    const int32_t* seg_k;
    const float* bimg;                 // [n_seg][kb][32 rows (16 hi, 16 lo) x 32] centroid image, swizzled
    const float* csn;                  // [n_seg][16] ||c||^2
    const int32_t* block_seg;
    const int64_t* block_row0;
    const int64_t* block_row1;
    const int32_t* active;
    int32_t* labels;
    float* psums;
    float* pcounts;
    int32_t* n_changed;
    int update;
    int debug;                         // tuning only (OODB200_TC_DEBUG): 1 no split work, 2 no MMA, 4 no accumulation
    const float* x;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// Spin with a wall-clock bound: a protocol error traps (the launch fails) instead of hanging the GPU.
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    uint64_t t0 = 0;
    for (uint32_t it = 0;; ++it) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(tc_smem_u32(bar)), "r"(parity)
                     : "memory");
        if (ok) return;
        if ((it & 1023u) == 1023u) {
            uint64_t t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tc_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(tc_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
                 "l"(map), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint32_t tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
// shared-memory matrix descriptor, K-major, SWIZZLE_32B: a row is ONE k-step (8 floats = 32 B), 8-row groups of 256 B
// (SBO), LBO = 1 (unused), version 1.  With 128-byte rows every tcgen05.mma fetched the whole 128-byte row of all M rows
// to use 32 bytes of it (~170 cycles per instruction at N <= 32); 32-byte rows make the operand fetch 4x smaller.
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}
// instruction descriptor: D = F32, A = B = TF32, K-major both, M = 128 (row i -> TMEM lane i), N = 32 (c_hi | c_lo) or 16 (c_hi)
constexpr uint32_t kTcIdescN32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kTcIdescN16 = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One lane of a CONVERGED warp.  The single-thread instructions (TMA, tcgen05.mma, tcgen05.commit) take their operands
// from uniform registers: under `if (lane == 0)` the compiler wraps every one of them in an ELECT / BRA.U.ANY loop
// (~40 issue slots per MMA, measured); with warp-uniform control flow + elect they are issued directly.
__device__ __forceinline__ bool tc_elect() {
    uint32_t pred;
    asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

template <int XS_, int LS_>
__global__ void __launch_bounds__(kTcThreads, 1) kmeans_step_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p) {
    extern __shared__ unsigned char tc_dyn[];
    __shared__ __align__(8) uint64_t s_afull[kTcMaxXs], s_done[kTcMaxXs], s_lofull[kTcMaxLs];   // s_done[n % XS]: the MMAs of k-block n have completed
    __shared__ __align__(8) uint64_t s_accfull[2], s_labready[2], s_mdone[2], s_bfull;
    __shared__ unsigned s_mask[2][4][16];                   // [tile parity][32-row group][cluster] (counts)
    __shared__ unsigned char s_order[2][kTcRows];           // rows of the tile sorted by (label, row)
    __shared__ int s_start[2][17];                          // first sorted position of every cluster; [16] = valid rows
    __shared__ float s_csn[16];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int g = p.block_seg[b];
    if (p.active && !p.active[g]) return;
    const int D = p.dim, KB = p.kb;
    const int Kg = p.seg_k[g];
    const int64_t r0 = p.block_row0[b], r1 = p.block_row1[b];
    const int n_tiles = (int)((r1 - r0 + kTcRows - 1) / kTcRows);
    constexpr uint32_t XS = XS_, LS = LS_;                  // raw ring / lo ring depth
    const uint32_t x_base = (tc_smem_u32(tc_dyn) + 1023u) & ~1023u;          // [XS][4 k-steps][128 rows][32 B]
    const uint32_t lo_base = x_base + XS * 16384u;                             // [LS][4 k-steps][128 rows][32 B]
    const uint32_t b_base = lo_base + LS * 16384u;                    // [KB][4 k-steps][32 rows: c_hi, c_lo][32 B]
    const uint32_t acc_base = b_base + (uint32_t)KB * 4096u;                    // [16][D] float32 partial sums

    if (tid == 0) {
        for (int s = 0; s < kTcMaxXs; ++s) {
            tc_mbar_init(&s_afull[s], 1);
            tc_mbar_init(&s_done[s], 1);
        }
        for (int s = 0; s < kTcMaxLs; ++s) tc_mbar_init(&s_lofull[s], kNSplit);
        for (int e = 0; e < 2; ++e) {
            tc_mbar_init(&s_accfull[e], 1);
            tc_mbar_init(&s_labready[e], kNSplit);
            tc_mbar_init(&s_mdone[e], kNAcc);
        }
        tc_mbar_init(&s_bfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(kTcCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 16) s_csn[tid] = tid < Kg ? p.csn[(size_t)g * 16 + tid] : FLT_MAX;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == kWTma) {                                                        // whole warp in the loop, one elected lane issues
        const uint32_t b_bytes = (uint32_t)KB * 4096u;
        if (tc_elect()) {
            tc_mbar_expect_tx(&s_bfull, b_bytes);
            tc_bulk_g2s(b_base, p.bimg + (size_t)g * KB * 1024, b_bytes, &s_bfull);
        }
        uint64_t pol_last;                                                      // the tile is read again by the M-step: keep it in L2
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
        uint32_t n = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int row = (int)(r0 + (int64_t)t * kTcRows);
            for (int kb = 0; kb < KB; ++kb, ++n) {
                const uint32_t s = n % XS, u = n / XS;
                if (u >= 1) tc_mbar_wait(&s_done[s], (u - 1) & 1u);           // MMAs of k-block n - XS done: the stage is free
                if (tc_elect()) {
                    tc_mbar_expect_tx(&s_afull[s], 16384u);
                    tc_tma_3d(x_base + s * 16384u, &tmap, 0, row, kb * 4, &s_afull[s], pol_last);
                }
            }
        }
        __syncwarp();
    } else if (warp == kWMma) {
        {
            tc_mbar_wait(&s_bfull, 0);
            uint32_t n = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const uint32_t acc = tmem + (uint32_t)(t & 1) * 32u;
                for (int kb = 0; kb < KB; ++kb, ++n) {
                    const uint32_t s = n % XS, u = n / XS, sl = n % LS, ul = n / LS;
                    const uint32_t a_raw = x_base + s * 16384u, a_lo = lo_base + sl * 16384u;
                    const uint32_t b_img = b_base + (uint32_t)kb * 4096u;       // rows 0-15 c_hi, rows 16-31 c_lo
                    tc_mbar_wait(&s_afull[s], u & 1u);                          // the raw products do not wait for the split
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (!(p.debug & 2) && tc_elect())
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)                              // UMMA_K = 8 floats: k-step k4 = the stage's k4-th [128 rows][32 B] block
                        tc_mma(acc, tc_desc(a_raw + k4 * 4096), tc_desc(b_img + k4 * 1024), kTcIdescN32, (kb | k4) != 0);   // cols 0-15 += x.c_hi, 16-31 += x.c_lo
                    tc_mbar_wait(&s_lofull[sl], ul & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (tc_elect()) {
                        if (!(p.debug & 2))
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc_mma(acc, tc_desc(a_lo + k4 * 4096), tc_desc(b_img + k4 * 1024), kTcIdescN16, 1u);           // cols 0-15 += x_lo.c_hi
                        tc_commit(&s_done[s]);                                  // ONE commit per k-block frees the raw and the lo stage
                        if (kb == KB - 1) tc_commit(&s_accfull[t & 1]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < kNSplit) {                                                // split, then epilogue of rows 32*warp ..
        int changed = 0;
        uint32_t n = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int e = t & 1, j = t >> 1;
            for (int kb = 0; kb < KB; ++kb, ++n) {
                const uint32_t s = n % XS, u = n / XS, sl = n % LS, ul = n / LS;
                tc_mbar_wait(&s_afull[s], u & 1u);
                if (n >= LS) tc_mbar_wait(&s_done[(n - LS) % XS], ((n - LS) / XS) & 1u);   // MMAs that read lo[sl] are done
                if (p.debug & 1) { __syncwarp(); if (lane == 0) tc_mbar_arrive(&s_lofull[sl]); continue; }
                const uint32_t src = x_base + s * 16384u + (uint32_t)warp * 4096u;
                const uint32_t dst = lo_base + sl * 16384u + (uint32_t)warp * 4096u;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t off = (uint32_t)((half * 4 + q) * 32 + lane) * 16u;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[q].x), "=f"(v[q].y), "=f"(v[q].z), "=f"(v[q].w) : "r"(src + off));
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t off = (uint32_t)((half * 4 + q) * 32 + lane) * 16u;
                        uint32_t o[4];
                        const float el[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float hi = __uint_as_float(__float_as_uint(el[c]) & 0xffffe000u);   // what kind::tf32 reads
                            // round-to-nearest (ties away) to 10 mantissa bits = cvt.rna.tf32 without the conversion pipe
                            o[c] = (__float_as_uint(el[c] - hi) + 0x1000u) & 0xffffe000u;
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + off), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&s_lofull[sl]);
            }
            // ---- epilogue of tile t: TMEM lanes 32*warp .. +31 = rows of this warp
            tc_mbar_wait(&s_accfull[e], (uint32_t)j & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t d[16], f[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)e * 32u;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
                  "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                : "r"(taddr));
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(f[0]), "=r"(f[1]), "=r"(f[2]), "=r"(f[3]), "=r"(f[4]), "=r"(f[5]), "=r"(f[6]), "=r"(f[7]), "=r"(f[8]),
                  "=r"(f[9]), "=r"(f[10]), "=r"(f[11]), "=r"(f[12]), "=r"(f[13]), "=r"(f[14]), "=r"(f[15])
                : "r"(taddr + 16u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float best = FLT_MAX;
            int lab = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float dot = __uint_as_float(d[k]) + __uint_as_float(f[k]);                    // (x.c_hi + x_lo.c_hi) + x.c_lo
                const float pd = k < Kg ? fmaf(-2.0f, dot, s_csn[k]) : FLT_MAX;                      // gemm(alpha=-2, beta=1) on ||c||^2
                if (pd < best) { best = pd; lab = k; }                                              // strict <: first minimum
            }
            const int64_t row = r0 + (int64_t)t * kTcRows + warp * 32 + lane;
            const bool valid = row < r1;
            if (valid) {
                if (p.labels[row] != lab) ++changed;
                p.labels[row] = lab;
            }
            if (j >= 1) tc_mbar_wait(&s_mdone[e], (uint32_t)(j - 1) & 1u);      // the M-step of tile t-2 has consumed these masks
            // rows of the tile sorted by (label, row): per-cluster masks of the 4 warps -> exclusive starts, then every
            // row's position = start[label] + rows with the same label before it
            unsigned mym = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned m = __ballot_sync(0xffffffffu, valid && lab == k);
                if (lane == k) s_mask[e][warp][k] = m;
                if (lab == k) mym = m;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");                     // the 4 epilogue warps
            int cntk = 0;
            if (lane < 16) cntk = __popc(s_mask[e][0][lane]) + __popc(s_mask[e][1][lane]) + __popc(s_mask[e][2][lane]) + __popc(s_mask[e][3][lane]);
            int incl = cntk;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const int startk = incl - cntk;                                     // lane k < 16: first position of cluster k
            const int my_start = __shfl_sync(0xffffffffu, startk, lab & 15);
            if (valid) {
                int pos = my_start + __popc(mym & ((1u << lane) - 1u));
                for (int w = 0; w < warp; ++w) pos += __popc(s_mask[e][w][lab]);
                s_order[e][pos] = (unsigned char)(warp * 32 + lane);
            }
            if (warp == 0 && lane < 16) s_start[e][lane] = startk;
            if (warp == 0 && lane == 15) s_start[e][16] = incl;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&s_labready[e]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
        if (lane == 0 && changed) atomicAdd(&p.n_changed[g], changed);
    } else {                                                                    // warps 4-8: M-step of the finished tile
        const int c = tid - kWAcc0 * 32;                                        // 0..159: one 16-byte column chunk each
        const bool have = c < D / 4;
        // per-cluster sums of this thread's 4 columns: thread-private [16][D] in shared memory (the cluster is a run-time
        // index); a whole RUN of rows of one cluster is added in registers between one load and one store of its slot
        float4* __restrict__ accp = reinterpret_cast<float4*>(tc_dyn + (acc_base - tc_smem_u32(tc_dyn))) + c;
        const int row_f4 = D / 4;
        if (have)
#pragma unroll
            for (int k = 0; k < 16; ++k) accp[k * row_f4] = make_float4(0.f, 0.f, 0.f, 0.f);
        float cnt = 0.f;
        uint64_t pol_first;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
        const bool accumulate = p.update && !(p.debug & 4);
        for (int t = 0; t < n_tiles; ++t) {
            const int e = t & 1, j = t >> 1;
            tc_mbar_wait(&s_labready[e], (uint32_t)j & 1u);
            if (accumulate) {
                const float* __restrict__ xt = p.x + (size_t)(r0 + (int64_t)t * kTcRows) * D + c * 4;
                const int n_valid = s_start[e][16];
                if (c < 16) cnt += (float)(s_start[e][c + 1] - s_start[e][c]);
                // the tile's rows in (label, row) order through a register ring of 4 batches x 4 rows (coalesced LDG.128
                // from L2: the tile was fetched microseconds ago); rows of a cluster arrive in increasing order
                auto load4 = [&](int bi, float4 (&v)[4]) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = bi * 4 + q;
                        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (have && i < n_valid) {                              // last use of the line: first to leave L2
                            const int r = s_order[e][i];
                            asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                                         : "=f"(v[q].x), "=f"(v[q].y), "=f"(v[q].z), "=f"(v[q].w)
                                         : "l"(xt + (size_t)r * D), "l"(pol_first));
                        }
                    }
                };
                int cur_k = -1, next_start = 0;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                auto consume4 = [&](int bi, const float4 (&v)[4]) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = bi * 4 + q;
                        if (i < n_valid) {
                            while (i >= next_start) {                           // next cluster (warp-uniform): swap the slot
                                if (cur_k >= 0 && have) accp[cur_k * row_f4] = a;
                                ++cur_k;
                                next_start = s_start[e][cur_k + 1];
                                if (have) a = accp[cur_k * row_f4];
                            }
                            a.x += v[q].x; a.y += v[q].y; a.z += v[q].z; a.w += v[q].w;
                        }
                    }
                };
                float4 v0[4], v1[4], v2[4], v3[4];
                load4(0, v0);
                load4(1, v1);
                load4(2, v2);
#pragma unroll 1
                for (int bi = 0; bi * 4 < n_valid; bi += 4) {
                    load4(bi + 3, v3);
                    consume4(bi, v0);
                    load4(bi + 4, v0);
                    consume4(bi + 1, v1);
                    load4(bi + 5, v1);
                    consume4(bi + 2, v2);
                    load4(bi + 6, v2);
                    consume4(bi + 3, v3);
                }
                if (cur_k >= 0 && have) accp[cur_k * row_f4] = a;
            }
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&s_mdone[e]);
        }
        if (p.update) {
            float* __restrict__ ps = p.psums + (size_t)b * p.k * D;
            if (have)
                for (int k = 0; k < p.k; ++k) *reinterpret_cast<float4*>(ps + (size_t)k * D + c * 4) = accp[k * row_f4];
            if (c < p.k) p.pcounts[(size_t)b * p.k + c] = cnt;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcCols) : "memory");
    }
}

// Centroids of every segment -> the shared-memory image the MMA reads (hi and lo pieces, K-major SWIZZLE_128B, 16 rows
// per 32-float k-block, rows >= seg_k zero) + ||c||^2 with the FP32 kernel's summation order.
__global__ void __launch_bounds__(256) kmeans_tc_prep_kernel(const float* __restrict__ cent, const int32_t* __restrict__ seg_k,
                                                             const int32_t* __restrict__ active, int k, int dim,
                                                             float* __restrict__ bimg, float* __restrict__ csn) {
    const int g = blockIdx.x;
    if (active && !active[g]) return;
    const int Kg = seg_k[g], KB = dim / 32;
    const float* __restrict__ cg = cent + (size_t)g * k * dim;
    uint32_t* __restrict__ img = reinterpret_cast<uint32_t*>(bimg) + (size_t)g * KB * 1024;
    for (int idx = threadIdx.x; idx < 16 * dim; idx += blockDim.x) {
        const int n = idx / dim, d = idx - n * dim;
        const float c = n < Kg ? cg[(size_t)n * dim + d] : 0.f;
        const uint32_t hi = tf32_rna(c);
        const uint32_t lo = tf32_rna(c - __uint_as_float(hi));
        const int kb = d >> 5, kc = d & 31;
        // k-block image: 4 k-steps x [32 rows (0-15 hi pieces, 16-31 lo pieces)][32 B], 8-row groups of 256 B,
        // 32-byte swizzle (16-byte chunk index ^ bit 2 of the row); offsets in floats
        const int k4 = kc >> 3, ch = (kc >> 2) & 1;
        const int off = kb * 1024 + k4 * 256 + (n >> 3) * 64 + (n & 7) * 8 + ((ch ^ ((n >> 2) & 1)) << 2) + (kc & 3);
        img[off] = hi;
        img[off + 128] = lo;                                       // rows 16-31 of the same k-step block
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int kk = warp; kk < 16; kk += 8) {
        float s = 0.f;
        if (kk < Kg)
            for (int d = lane; d < dim; d += 32) s = fmaf(cg[(size_t)kk * dim + d], cg[(size_t)kk * dim + d], s);
        s = warp_sum(s);
        if (lane == 0) csn[(size_t)g * 16 + kk] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K2b: nearest-centroid distance of already pooled vectors on the tensor pipe, K <= 64 (BASELINE config C5: K = 64 clusters
// per class).  Replaces `pairwise_distances(cluster, activations, metric).min(axis=0)` (/root/reference/ood_utils.py:2422-2430)
// for 'l2' (sklearn `euclidean_distances`: XX - 2 X.Y^T + YY combined in float64, cast to float32, sqrt) and 'cosine'.
// Same pipeline as the Lloyd step (128-row tiles, 32-byte-swizzle operands, split-float x), but the centroid image of
// K = 64 (hi | lo = 128 rows x D) does not fit shared memory: its k-block travels with the rows' k-block through the
// same ring stage (it is L2 resident), N = 128 for the raw products and 64 for the lo products, 2 x 128 TMEM columns.
constexpr int kVtThreads = 192;        // warps 0-3 split + epilogue, warp 4 TMA, warp 5 MMA
constexpr int kVtStages = 4, kVtLo = 2;
constexpr int kVtCols = 256;
constexpr uint32_t kVtIdescN128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kVtIdescN64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

struct VtParams {
    int dim, kb, n_seg, metric;        // metric: OODB200_METRIC_L2 or OODB200_METRIC_COS
    int64_t n_rows;
    const int32_t* cent_k;
    const float* bimg;                 // [n_seg][kb][4 k-steps][128 rows: 64 hi, 64 lo][8 floats], swizzled
    const double* cc;                  // [n_seg][64] ||c||^2 (l2)
    const double* xx;                  // [n_rows] ||x||^2
    const int32_t* block_seg;
    const int64_t* block_row0;
    const int64_t* block_row1;
    float* dist;                       // [3][n_rows]
    int32_t* argmin;
    const double* thr;                 // [3][n_seg] or null
    uint8_t* decision;
};

__global__ void __launch_bounds__(kVtThreads, 1) vec_score_tc_kernel(const __grid_constant__ CUtensorMap tmap, const VtParams p) {
    extern __shared__ unsigned char tc_dyn[];
    __shared__ __align__(8) uint64_t s_afull[kVtStages], s_done[kVtStages], s_lofull[kVtLo], s_accfull[2];
    __shared__ uint32_t s_tmem;
    constexpr uint32_t XS = kVtStages, LS = kVtLo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int g = p.block_seg[b];
    const int KB = p.kb;
    const int Kg = p.cent_k[g];
    const int64_t r0 = p.block_row0[b], r1 = p.block_row1[b];
    const int n_tiles = (int)((r1 - r0 + kTcRows - 1) / kTcRows);
    const uint32_t x_base = (tc_smem_u32(tc_dyn) + 1023u) & ~1023u;          // [XS][4 k-steps][128 rows][32 B]
    const uint32_t b_base = x_base + XS * 16384u;                              // [XS][4 k-steps][128 rows: c_hi, c_lo][32 B]
    const uint32_t lo_base = b_base + XS * 16384u;                             // [LS][4 k-steps][128 rows][32 B]
    if (tid == 0) {
        for (int s = 0; s < kVtStages; ++s) { tc_mbar_init(&s_afull[s], 1); tc_mbar_init(&s_done[s], 1); }
        for (int s = 0; s < kVtLo; ++s) tc_mbar_init(&s_lofull[s], 4);
        tc_mbar_init(&s_accfull[0], 1);
        tc_mbar_init(&s_accfull[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(kVtCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const bool work = Kg > 0;                                                   // a segment without centroids: 1000 / -1 below

    if (warp == 4) {                                                            // TMA producer
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        uint32_t n = 0;
        for (int t = 0; work && t < n_tiles; ++t) {
            const int row = (int)(r0 + (int64_t)t * kTcRows);
            for (int kb = 0; kb < KB; ++kb, ++n) {
                const uint32_t s = n % XS, u = n / XS;
                if (u >= 1) tc_mbar_wait(&s_done[s], (u - 1) & 1u);
                if (tc_elect()) {
                    tc_mbar_expect_tx(&s_afull[s], 32768u);
                    tc_tma_3d(x_base + s * 16384u, &tmap, 0, row, kb * 4, &s_afull[s], pol);
                    tc_bulk_g2s(b_base + s * 16384u, p.bimg + ((size_t)g * KB + kb) * 4096, 16384u, &s_afull[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {                                                     // MMA issuer
        uint32_t n = 0;
        for (int t = 0; work && t < n_tiles; ++t) {
            const uint32_t acc = tmem + (uint32_t)(t & 1) * 128u;
            for (int kb = 0; kb < KB; ++kb, ++n) {
                const uint32_t s = n % XS, u = n / XS, sl = n % LS, ul = n / LS;
                const uint32_t a_raw = x_base + s * 16384u, a_lo = lo_base + sl * 16384u, b_img = b_base + s * 16384u;
                tc_mbar_wait(&s_afull[s], u & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tc_elect())
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)                              // cols 0-63 += x.c_hi, cols 64-127 += x.c_lo
                        tc_mma(acc, tc_desc(a_raw + k4 * 4096), tc_desc(b_img + k4 * 4096), kVtIdescN128, (kb | k4) != 0);
                tc_mbar_wait(&s_lofull[sl], ul & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tc_elect()) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)                              // cols 0-63 += x_lo.c_hi
                        tc_mma(acc, tc_desc(a_lo + k4 * 4096), tc_desc(b_img + k4 * 4096), kVtIdescN64, 1u);
                    tc_commit(&s_done[s]);
                    if (kb == KB - 1) tc_commit(&s_accfull[t & 1]);
                }
                __syncwarp();
            }
        }
    } else {                                                                    // warps 0-3: split, then epilogue of rows 32*warp ..
        uint32_t n = 0;
        const int slot = p.metric;
        const double thr = p.thr ? p.thr[(size_t)slot * p.n_seg + g] : 0.0;
        for (int t = 0; t < n_tiles; ++t) {
            const int e = t & 1, j = t >> 1;
            for (int kb = 0; work && kb < KB; ++kb, ++n) {
                const uint32_t s = n % XS, u = n / XS, sl = n % LS;
                tc_mbar_wait(&s_afull[s], u & 1u);
                if (n >= LS) tc_mbar_wait(&s_done[(n - LS) % XS], ((n - LS) / XS) & 1u);
                const uint32_t src = x_base + s * 16384u + (uint32_t)warp * 4096u;
                const uint32_t dst = lo_base + sl * 16384u + (uint32_t)warp * 4096u;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t off = (uint32_t)((half * 4 + q) * 32 + lane) * 16u;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[q].x), "=f"(v[q].y), "=f"(v[q].z), "=f"(v[q].w) : "r"(src + off));
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t off = (uint32_t)((half * 4 + q) * 32 + lane) * 16u;
                        uint32_t o[4];
                        const float el[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float hi = __uint_as_float(__float_as_uint(el[c]) & 0xffffe000u);
                            o[c] = (__float_as_uint(el[c] - hi) + 0x1000u) & 0xffffe000u;
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + off), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&s_lofull[sl]);
            }
            // ---- epilogue: lane = row
            const int64_t row = r0 + (int64_t)t * kTcRows + warp * 32 + lane;
            const bool valid = row < r1;
            float best = FLT_MAX;
            int barg = -1;
            if (work) {
                tc_mbar_wait(&s_accfull[e], (uint32_t)j & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const double xx = valid ? p.xx[row] : 1.0;
                const float inv_norm = (float)(1.0 / sqrt(xx > 0.0 ? xx : 1.0));
#pragma unroll 1
                for (int c16 = 0; c16 < 4; ++c16) {
                    uint32_t d[16], f[16];
                    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)e * 128u + (uint32_t)c16 * 16u;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
                          "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                        : "r"(taddr));
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(f[0]), "=r"(f[1]), "=r"(f[2]), "=r"(f[3]), "=r"(f[4]), "=r"(f[5]), "=r"(f[6]), "=r"(f[7]), "=r"(f[8]),
                          "=r"(f[9]), "=r"(f[10]), "=r"(f[11]), "=r"(f[12]), "=r"(f[13]), "=r"(f[14]), "=r"(f[15])
                        : "r"(taddr + 64u));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int k = c16 * 16 + i;
                        if (k < Kg) {
                            const float dot = __uint_as_float(d[i]) + __uint_as_float(f[i]);
                            float val;
                            if (p.metric == OODB200_METRIC_L2)                  // XX - 2 X.Y + YY in float64, cast, clamp (sklearn)
                                val = fmaxf((float)(xx - 2.0 * (double)dot + p.cc[(size_t)g * 64 + k]), 0.f);
                            else                                                // 1 - cos, clipped to [0, 2] (sklearn cosine_distances)
                                val = fminf(fmaxf(1.0f - dot * inv_norm, 0.f), 2.f);
                            if (val < best) { best = val; barg = k; }           // strict <: first minimum
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            if (valid) {
                const size_t o = (size_t)slot * p.n_rows + row;
                const float dv = work ? (p.metric == OODB200_METRIC_L2 ? sqrtf(best) : best) : 1000.f;   // no cluster: ood_utils.py:2159-2164
                p.dist[o] = dv;
                p.argmin[o] = barg;
                if (p.decision) p.decision[o] = (thr == thr && (double)dv < thr) ? 1 : 0;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kVtCols) : "memory");
    }
}

// centroids of every segment -> streamed image [kb][4 k-steps][128 rows: 64 hi pieces, 64 lo pieces][8 floats] + ||c||^2
__global__ void __launch_bounds__(256) vec_score_tc_prep_kernel(const float* __restrict__ cent, const int64_t* __restrict__ cent_row_off,
                                                                const int32_t* __restrict__ cent_k, int dim, float* __restrict__ bimg,
                                                                double* __restrict__ cc) {
    const int g = blockIdx.x;
    const int Kg = cent_k[g], KB = dim / 32;
    const float* __restrict__ cg = cent + (size_t)cent_row_off[g] * dim;
    uint32_t* __restrict__ img = reinterpret_cast<uint32_t*>(bimg) + (size_t)g * KB * 4096;
    for (int idx = threadIdx.x; idx < 64 * dim; idx += blockDim.x) {
        const int n = idx / dim, d = idx - n * dim;
        const float c = n < Kg ? cg[(size_t)n * dim + d] : 0.f;
        const uint32_t hi = tf32_rna(c);
        const uint32_t lo = tf32_rna(c - __uint_as_float(hi));
        const int kb = d >> 5, kc = d & 31, k4 = kc >> 3, ch = (kc >> 2) & 1;
        const int off = kb * 4096 + k4 * 1024 + (n >> 3) * 64 + (n & 7) * 8 + ((ch ^ ((n >> 2) & 1)) << 2) + (kc & 3);
        img[off] = hi;
        img[off + 512] = lo;                                       // rows 64-127 of the same k-step block
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int kk = warp; kk < 64; kk += 8) {
        double s = 0.0;
        if (kk < Kg)
            for (int d = lane; d < dim; d += 32) s += (double)cg[(size_t)kk * dim + d] * (double)cg[(size_t)kk * dim + d];
        for (int o = 16; o > 0; o >>= 1) {
            int lo = __double2loint(s), hi = __double2hiint(s);
            lo = __shfl_xor_sync(0xffffffffu, lo, o);
            hi = __shfl_xor_sync(0xffffffffu, hi, o);
            s += __hiloint2double(hi, lo);
        }
        if (lane == 0) cc[(size_t)g * 64 + kk] = s;
    }
}

// xx[r] = sum_d x[r, d]^2 in float64, one warp per row
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ x, int dim, int64_t n_rows, double* __restrict__ xx) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_rows; r += (int64_t)gridDim.x * 8) {
        double s = 0.0;
        for (int d = lane; d < dim; d += 32) { const double v = (double)__ldg(x + r * dim + d); s += v * v; }
        for (int o = 16; o > 0; o >>= 1) {
            int lo = __double2loint(s), hi = __double2hiint(s);
            lo = __shfl_xor_sync(0xffffffffu, lo, o);
            hi = __shfl_xor_sync(0xffffffffu, hi, o);
            s += __hiloint2double(hi, lo);
        }
        if (lane == 0) xx[r] = s;
    }
}

typedef CUresult (*TcEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TcEncodeFn tc_encode_fn() {
    static TcEncodeFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (TcEncodeFn)f;
    }
    return fn;
}

static int tc_lo_stages() {
    const char* e = getenv("OODB200_TC_LS");
    const int v = e ? atoi(e) : 2;
    return v < 2 ? 2 : (v > 3 ? 3 : v);
}
static int tc_x_stages(int dim) {          // as many raw stages as fit next to the lo ring, the centroid image and the sums
    const long fixed = (long)(dim / 32) * 4096 + (long)tc_lo_stages() * 16384 + 64L * dim + 1024 + 1024;
    long xs = (232448 - fixed) / 16384;
    const char* e = getenv("OODB200_TC_XS");
    if (e && atoi(e) >= 2 && atoi(e) < xs) xs = atoi(e);
    return (int)(xs > kTcMaxXs ? kTcMaxXs : xs);   // the lo ring is never deeper than the raw ring (s_done indexing)
}
static size_t tc_smem_bytes(int dim) {
    return (size_t)(dim / 32) * 4096 + (size_t)(tc_x_stages(dim) + tc_lo_stages()) * 16384 + (size_t)64 * dim + 1024;
}

}  // namespace oodb200

using namespace oodb200;

extern "C" int64_t oodb200_kmeans_tc_workspace_bytes(int n_seg, int k, int dim) {
    if (k < 1 || k > 16 || dim < 128 || dim % 32 != 0 || dim / 32 > kTcMaxKb || n_seg < 1) return 0;
    return (int64_t)n_seg * ((int64_t)(dim / 32) * 4096 + 64);
}

extern "C" int oodb200_kmeans_step_tc_f32(const float* x, int64_t n_rows, int dim, int n_seg, int k, const int32_t* seg_k,
                                          const float* cent, const int32_t* block_seg, const int64_t* block_row0,
                                          const int64_t* block_row1, int n_blocks, const int32_t* active, int32_t* labels,
                                          float* psums, float* pcounts, int32_t* n_changed, int update, void* workspace,
                                          void* stream) {
    OODB200_REQUIRE(oodb200_kmeans_tc_workspace_bytes(n_seg > 0 ? n_seg : 1, k, dim) > 0,
                    "kmeans_step_tc: needs k <= 16 and dim %% 32 == 0, 128 <= dim <= 640 (k = %d, dim = %d)", k, dim);
    OODB200_REQUIRE(update == 0 || update == 1, "kmeans_step_tc: update must be 0 or 1");
    OODB200_REQUIRE(n_rows > 0 && n_rows < INT_MAX && n_blocks >= 0, "kmeans_step_tc: bad size");
    if (n_blocks == 0) return OODB200_OK;
    OODB200_REQUIRE(x && seg_k && cent && block_seg && block_row0 && block_row1 && labels && n_changed && workspace,
                    "kmeans_step_tc: null pointer");
    OODB200_REQUIRE(!update || (psums && pcounts), "kmeans_step_tc: update needs the partial buffers");
    OODB200_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "kmeans_step_tc: x / workspace must be 16-byte aligned");
    TcEncodeFn enc = tc_encode_fn();
    OODB200_REQUIRE(enc != nullptr, "kmeans_step_tc: cuTensorMapEncodeTiled is not available from this driver");
    const int KB = dim / 32;
    CUtensorMap tmap;
    // X [n_rows, dim] seen as (8 floats of a k-step, row, k-step): one copy = the 4 k-steps of a k-block of a 128-row
    // tile, landing as [k-step][row][32 B] with the 32-byte swizzle the MMA descriptors name
    const cuuint64_t gdim[3] = {8, (cuuint64_t)n_rows, (cuuint64_t)dim / 8};
    const cuuint64_t gstride[2] = {(cuuint64_t)dim * 4, 32};
    const cuuint32_t box[3] = {8, (cuuint32_t)kTcRows, 4};
    const cuuint32_t estride[3] = {1, 1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), gdim, gstride, box, estride,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("kmeans_step_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr); return OODB200_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    float* bimg = reinterpret_cast<float*>(workspace);
    float* csn = bimg + (size_t)n_seg * KB * 1024;
    kmeans_tc_prep_kernel<<<n_seg, 256, 0, st>>>(cent, seg_k, active, k, dim, bimg, csn);
    int rc = check_launch("kmeans_tc_prep");
    if (rc) return rc;
    const char* dbg = getenv("OODB200_TC_DEBUG");
    TcParams p = {dim, KB, k, seg_k, bimg, csn, block_seg, block_row0, block_row1, active, labels, psums, pcounts, n_changed, update,
                  dbg ? atoi(dbg) : 0, x};
    const size_t smem = tc_smem_bytes(dim);
    const int xs = tc_x_stages(dim), ls = tc_lo_stages();
    cudaError_t e = cudaErrorInvalidValue;
#define OODB200_TC_LAUNCH(XS, LS)                                                                                          \
    if (xs == XS && ls == LS) {                                                                                            \
        e = cudaFuncSetAttribute(kmeans_step_tc_kernel<XS, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e == cudaSuccess) kmeans_step_tc_kernel<XS, LS><<<n_blocks, kTcThreads, smem, st>>>(tmap, p);                  \
    }
    OODB200_TC_LAUNCH(3, 2) OODB200_TC_LAUNCH(4, 2) OODB200_TC_LAUNCH(5, 2) OODB200_TC_LAUNCH(6, 2) OODB200_TC_LAUNCH(7, 2)
    OODB200_TC_LAUNCH(3, 3) OODB200_TC_LAUNCH(4, 3) OODB200_TC_LAUNCH(5, 3) OODB200_TC_LAUNCH(6, 3)
#undef OODB200_TC_LAUNCH
    if (e != cudaSuccess) { set_error("kmeans_step_tc: %s (xs %d, ls %d)", cudaGetErrorString(e), xs, ls); return OODB200_ERR_CUDA; }
    return check_launch("kmeans_step_tc");
}

extern "C" int64_t oodb200_vec_score_tc_workspace_bytes(int n_seg, int64_t n_rows, int dim) {
    if (dim < 128 || dim % 32 != 0 || dim > 2048 || n_seg < 1 || n_rows < 0) return 0;
    return (int64_t)n_seg * ((int64_t)(dim / 32) * 16384 + 64 * 8) + n_rows * 8 + 256;
}

extern "C" int oodb200_vec_score_tc_f32(const float* x, int64_t n_rows, int dim, int n_seg, int metric, const float* cent,
                                        const int64_t* cent_row_off, const int32_t* cent_k, int max_k, const int32_t* block_seg,
                                        const int64_t* block_row0, const int64_t* block_row1, int n_blocks, float* dist,
                                        int32_t* argmin, const double* thr, uint8_t* decision, void* workspace, void* stream) {
    OODB200_REQUIRE(oodb200_vec_score_tc_workspace_bytes(n_seg > 0 ? n_seg : 1, n_rows, dim) > 0,
                    "vec_score_tc: needs dim %% 32 == 0, 128 <= dim <= 2048 (dim = %d)", dim);
    OODB200_REQUIRE(metric == OODB200_METRIC_L2 || metric == OODB200_METRIC_COS, "vec_score_tc: metric must be l2 or cosine");
    OODB200_REQUIRE(max_k >= 0 && max_k <= 64, "vec_score_tc: at most 64 centroids per segment (max_k = %d)", max_k);
    OODB200_REQUIRE(n_rows >= 0 && n_rows < INT_MAX && n_blocks >= 0, "vec_score_tc: bad size");
    if (n_blocks == 0 || n_rows == 0) return OODB200_OK;
    OODB200_REQUIRE(x && cent && cent_row_off && cent_k && block_seg && block_row0 && block_row1 && dist && argmin && workspace,
                    "vec_score_tc: null pointer");
    OODB200_REQUIRE(!decision || thr, "vec_score_tc: decision needs thresholds");
    OODB200_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)workspace & 255) == 0, "vec_score_tc: x must be 16-byte, workspace 256-byte aligned");
    TcEncodeFn enc = tc_encode_fn();
    OODB200_REQUIRE(enc != nullptr, "vec_score_tc: cuTensorMapEncodeTiled is not available from this driver");
    const int KB = dim / 32;
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {8, (cuuint64_t)n_rows, (cuuint64_t)dim / 8};
    const cuuint64_t gstride[2] = {(cuuint64_t)dim * 4, 32};
    const cuuint32_t box[3] = {8, (cuuint32_t)kTcRows, 4};
    const cuuint32_t estride[3] = {1, 1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), gdim, gstride, box, estride,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("vec_score_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr); return OODB200_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    float* bimg = reinterpret_cast<float*>(workspace);
    double* cc = reinterpret_cast<double*>(bimg + (size_t)n_seg * KB * 4096);
    double* xx = cc + (size_t)n_seg * 64;
    vec_score_tc_prep_kernel<<<n_seg, 256, 0, st>>>(cent, cent_row_off, cent_k, dim, bimg, cc);
    int rc = check_launch("vec_score_tc_prep");
    if (rc) return rc;
    long long gx = (n_rows + 7) / 8;
    if (gx > 148LL * 16) gx = 148LL * 16;
    row_sqnorm_kernel<<<(int)gx, 256, 0, st>>>(x, dim, n_rows, xx);
    rc = check_launch("row_sqnorm");
    if (rc) return rc;
    const size_t smem = (size_t)(2 * kVtStages + kVtLo) * 16384 + 1024;
    cudaError_t e = cudaFuncSetAttribute(vec_score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("vec_score_tc: %s", cudaGetErrorString(e)); return OODB200_ERR_CUDA; }
    VtParams p = {dim, KB, n_seg, metric, n_rows, cent_k, bimg, cc, xx, block_seg, block_row0, block_row1, dist, argmin, thr, decision};
    vec_score_tc_kernel<<<n_blocks, kVtThreads, smem, st>>>(tmap, p);
    return check_launch("vec_score_tc");
}
