// K3 logit-method scoring and K6 fusion rules, sm_100a.
//
// Replaces /root/reference/ood_utils.py:1195-1257 (decision + INDness per box) and :1388-1443
// (MSP / Energy / ODIN / Sigmoid scorers, torch CPU float32), plus FusionMethod.fuse_ood_decisions
// (:2906-2940) and TripleFusionMethod.fuse_ood_decisions (:3282-3301).
// One warp per detection: the [NC] logit row is read once (coalesced) and every requested method is
// evaluated from it -- max, exp-sum (softmax / logsumexp), temperature-scaled variants, sigmoid.
// Bandwidth-bound: 4*NC+4 bytes in, 9 bytes out per method and box.
#include "common.cuh"

#include <float.h>

namespace oodb200 {

constexpr int kLogitThreads = 256;

struct LogitParams {
    const float* logits;
    const int32_t* cls;
    int n, nc, method_mask;
    float t_energy, t_odin;
    const double* thr;
    const double* smin;
    const double* smax;
    int clip;
    float* scores;
    float* indness;
    uint8_t* decision;
    int32_t* sigmoid_mismatch;
};

// LogitsMethod.compute_indness (ood_utils.py:1224-1257), python-float (float64) arithmetic
__device__ __forceinline__ double indness_of(double score, double t, double mn, double mx, int clip) {
    double a = 0.0, b = 0.0;
    if (score > t) {
        a = 1.0 / (mx - t);
        b = -t / (mx - t);
    } else if (score < t) {
        a = -1.0 / (mn - t);
        b = t / (mn - t);
    }
    double v = a * score + b;
    if (clip) v = fmax(-1.0, fmin(v, 1.0));
    return v;
}

__global__ void __launch_bounds__(kLogitThreads) logit_kernel(const LogitParams p) {
    const int lane = threadIdx.x & 31;
    const int box = blockIdx.x * (kLogitThreads / 32) + (threadIdx.x >> 5);
    if (box >= p.n) return;
    const float* __restrict__ z = p.logits + (size_t)box * p.nc;
    const int cls = p.cls[box];
    const bool cls_ok = cls >= 0 && cls < p.nc;
    const bool want_odin = p.method_mask >> OODB200_LOGIT_ODIN & 1;
    const bool want_energy = p.method_mask >> OODB200_LOGIT_ENERGY & 1;

    // pass 1: maxima (raw, /T_energy, /T_odin); NC is 20 or 80 -> the row sits in L1 for pass 2
    float m = -FLT_MAX, me = -FLT_MAX, mo = -FLT_MAX;
    int am = -1;
    for (int j = lane; j < p.nc; j += 32) {
        const float v = __ldg(z + j);
        if (v > m) { m = v; am = j; }
        if (want_energy) me = fmaxf(me, __fdiv_rn(v, p.t_energy));
        if (want_odin) mo = fmaxf(mo, __fdiv_rn(v, p.t_odin));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {               // arg-max with first-index tie-break (numpy argmax)
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa >= 0 && (am < 0 || oa < am))) { m = om; am = oa; }
    }
    me = warp_max(me);
    mo = warp_max(mo);
    // pass 2: exp sums
    float se = 0.f, ss = 0.f, so = 0.f;
    for (int j = lane; j < p.nc; j += 32) {
        const float v = __ldg(z + j);
        ss += expf(v - m);
        if (want_energy) se += expf(__fdiv_rn(v, p.t_energy) - me);
        if (want_odin) so += expf(__fdiv_rn(v, p.t_odin) - mo);
    }
    ss = warp_sum(ss);
    se = warp_sum(se);
    so = warp_sum(so);
    if (lane != 0) return;

    const float zc = cls_ok ? __ldg(z + cls) : 0.f;
    float sc[OODB200_N_LOGIT];
    sc[OODB200_LOGIT_MSP] = cls_ok ? __fdiv_rn(expf(zc - m), ss) : 0.f;                                  // :1394-1397
    sc[OODB200_LOGIT_ENERGY] = want_energy ? p.t_energy * (me + logf(se)) : 0.f;                          // :1409-1412
    sc[OODB200_LOGIT_ODIN] = (want_odin && cls_ok) ? __fdiv_rn(expf(__fdiv_rn(zc, p.t_odin) - mo), so) : 0.f;  // :1424-1427
    // :1436-1443; with OODB200_LOGIT_FLAG_POST_SIGMOID the inputs already went through the detector's sigmoid
    // (use_values_before_sigmoid=False, :1438-1439): the score is the input value itself
    sc[OODB200_LOGIT_SIGMOID] = !cls_ok ? 0.f : ((p.method_mask & OODB200_LOGIT_FLAG_POST_SIGMOID) ? zc : __fdiv_rn(1.0f, 1.0f + expf(-zc)));
    sc[OODB200_LOGIT_MAXLOGIT] = m;                                                                      // no reference (Q7)
    if ((p.method_mask >> OODB200_LOGIT_SIGMOID & 1) && p.sigmoid_mismatch && am != cls) atomicAdd(p.sigmoid_mismatch, 1);
#pragma unroll
    for (int k = 0; k < OODB200_N_LOGIT; ++k) {
        if (!(p.method_mask >> k & 1)) continue;
        const size_t o = (size_t)k * p.n + box;
        p.scores[o] = sc[k];
        if (p.thr) {
            const double t = cls_ok ? p.thr[(size_t)k * p.nc + cls] : 0.0;
            if (p.decision) p.decision[o] = ((double)sc[k] < t) ? 0 : 1;                                  // :1203-1206
            if (p.indness && p.smin && p.smax) {
                const double mn = cls_ok ? p.smin[(size_t)k * p.nc + cls] : 0.0;
                const double mx = cls_ok ? p.smax[(size_t)k * p.nc + cls] : 0.0;
                p.indness[o] = (float)indness_of((double)sc[k], t, mn, mx, p.clip);
            }
        }
    }
}

__global__ void fuse_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                               const uint8_t* __restrict__ c, int n, int strategy, uint8_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = a[i], y = b[i];
    int r;
    if (strategy == OODB200_FUSE_AND) r = max(x, y);
    else if (strategy == OODB200_FUSE_OR) r = min(x, y);
    else r = (x + y + (int)c[i]) >= 2;
    out[i] = (uint8_t)r;
}

__global__ void fuse_score_kernel(const float* __restrict__ s1, const float* __restrict__ s2, int n,
                                  uint8_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ((double)s1[i] + (double)s2[i] > 0.0) ? 1 : 0;
}

static thread_local char g_err[512] = "";
char* error_buffer() { return g_err; }

}  // namespace oodb200

using namespace oodb200;

extern "C" int oodb200_abi_version(void) { return OODB200_ABI_VERSION; }
extern "C" const char* oodb200_last_error(void) { return error_buffer(); }

extern "C" int oodb200_logit_score_f32(const float* logits, const int32_t* cls, int n, int nc, int method_mask,
                                       float t_energy, float t_odin, const double* thr, const double* smin,
                                       const double* smax, int clip_indness, float* scores, float* indness,
                                       uint8_t* decision, int32_t* sigmoid_mismatch, void* stream) {
    OODB200_REQUIRE(n >= 0 && nc > 0, "logit_score: bad size");
    OODB200_REQUIRE((method_mask & ((1 << OODB200_N_LOGIT) - 1)) != 0 &&
                    (method_mask & ~(((1 << OODB200_N_LOGIT) - 1) | OODB200_LOGIT_FLAG_POST_SIGMOID)) == 0,
                    "logit_score: method_mask %d", method_mask);
    OODB200_REQUIRE(t_energy != 0.f && t_odin != 0.f, "logit_score: zero temperature");
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(logits && cls && scores, "logit_score: null pointer");
    LogitParams p = {logits, cls, n, nc, method_mask, t_energy, t_odin, thr, smin, smax, clip_indness,
                     scores, indness, decision, sigmoid_mismatch};
    const int per_block = kLogitThreads / 32;
    logit_kernel<<<(n + per_block - 1) / per_block, kLogitThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch("logit_score");
}

extern "C" int oodb200_fuse_u8(const uint8_t* a, const uint8_t* b, const uint8_t* c, int n, int strategy,
                               uint8_t* out, void* stream) {
    OODB200_REQUIRE(n >= 0, "fuse: negative n");
    OODB200_REQUIRE(strategy >= OODB200_FUSE_AND && strategy <= OODB200_FUSE_MAJORITY, "fuse: strategy %d", strategy);
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(a && b && out, "fuse: null pointer");
    OODB200_REQUIRE(strategy != OODB200_FUSE_MAJORITY || c, "fuse: majority vote needs three inputs");
    fuse_u8_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, b, c, n, strategy, out);
    return check_launch("fuse_u8");
}

extern "C" int oodb200_fuse_score_f32(const float* s1, const float* s2, int n, uint8_t* out, void* stream) {
    OODB200_REQUIRE(n >= 0, "fuse_score: negative n");
    if (n == 0) return OODB200_OK;
    OODB200_REQUIRE(s1 && s2 && out, "fuse_score: null pointer");
    fuse_score_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(s1, s2, n, out);
    return check_launch("fuse_score");
}
