// NMS with the OoD payload, sm_100a: the step immediately before the scoring hot path.
//
// Replaces the default path of /root/reference/ultralytics/utils/ops.py:348-530 (`non_max_suppression_old`): per image
//   candidates = anchors whose best class confidence exceeds conf_thres (:412, :463-466)
//   xywh -> xyxy (:455-456, xywh2xyxy :645-649), confidence = best class, descending sort (:478-482, at most max_nms)
//   class-aware greedy NMS (boxes offset by cls * max_wh, torchvision nms: IoU in float32, strict '>') (:485-489), max_det (:490)
//   the payload rows the reference gathers with the same indices: extra_item (raw class logits per anchor) and strides.
// The reference does this with ~10 small launches + torchvision.nms per IMAGE in a python loop; here ONE CTA per image does
// all of it: a coalesced pass over the [4 + nc, A] prediction slab (lanes over anchors), order-preserving compaction of the
// candidates, their descending rank by counting (stable in anchor order), the greedy suppression with the CTA's threads
// over the remaining boxes, and the gather of the kept rows and their payload.
#include "common.cuh"

#include <float.h>

namespace oodb200 {

constexpr int kNmsThreads = 256;

struct NmsParams {
    const float* pred;          // [bs, 4 + nc, A]
    const float* extra;         // [bs, ne, A] or null
    const float* strides;       // [A] or null
    int bs, nc, ne, A;
    float conf_thres, iou_thres, max_wh;
    int max_det, max_nms;
    // scratch, per image slabs of A entries
    int32_t* cand;              // anchor of every candidate (anchor order)
    float* conf;                // its confidence
    int32_t* cls;               // its class
    int32_t* order;             // candidate index at every sorted position
    float4* box;                // offset boxes in sorted order
    uint8_t* supp;
    // outputs
    float* det;                 // [bs, max_det, 6]
    float* out_extra;           // [bs, max_det, ne]
    float* out_strides;         // [bs, max_det]
    int32_t* out_anchor;        // [bs, max_det] anchor index of every kept detection
    int32_t* count;             // [bs]
};

__global__ void __launch_bounds__(kNmsThreads) nms_kernel(const NmsParams p) {
    __shared__ int s_scan[kNmsThreads / 32];
    __shared__ int s_base, s_kept[1024];
    const int xi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = p.A, nc = p.nc;
    const float* __restrict__ pr = p.pred + (size_t)xi * (4 + nc) * A;
    int32_t* cand = p.cand + (size_t)xi * A;
    float* conf = p.conf + (size_t)xi * A;
    int32_t* cls = p.cls + (size_t)xi * A;
    int32_t* order = p.order + (size_t)xi * A;
    float4* box = p.box + (size_t)xi * A;
    uint8_t* supp = p.supp + (size_t)xi * A;
    if (tid == 0) s_base = 0;
    __syncthreads();
    // ---- candidates: best class confidence > conf_thres, compacted in anchor order
    for (int a0 = 0; a0 < A; a0 += kNmsThreads) {
        const int a = a0 + tid;
        float best = -FLT_MAX;
        int bj = 0;
        if (a < A)
            for (int c = 0; c < nc; ++c) {             // lanes over anchors: coalesced rows of the slab; first maximum like torch.max
                const float v = __ldg(pr + (size_t)(4 + c) * A + a);
                if (v > best) { best = v; bj = c; }
            }
        const bool take = a < A && best > p.conf_thres;
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (lane == 0) s_scan[warp] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_scan[w];
        if (take) {
            const int i = off + __popc(m & ((1u << lane) - 1u));
            cand[i] = a;
            conf[i] = best;
            cls[i] = bj;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < kNmsThreads / 32; ++w) t += s_scan[w]; s_base += t; }
        __syncthreads();
    }
    const int m = s_base;
    if (m == 0) { if (tid == 0) p.count[xi] = 0; return; }
    // ---- descending order by confidence (ties: anchor order), by counting; offset boxes in sorted order
    for (int i = tid; i < m; i += kNmsThreads) {
        const float ci = conf[i];
        int r = 0;
        for (int j = 0; j < m; ++j) {
            const float cj = conf[j];
            r += (cj > ci) || (cj == ci && j < i);
        }
        order[r] = i;
    }
    __syncthreads();
    const int n = min(m, p.max_nms);
    for (int r = tid; r < n; r += kNmsThreads) {
        const int i = order[r], a = cand[i];
        const float cx = __ldg(pr + a), cy = __ldg(pr + (size_t)A + a);
        const float hw = __fdiv_rn(__ldg(pr + (size_t)2 * A + a), 2.f), hh = __fdiv_rn(__ldg(pr + (size_t)3 * A + a), 2.f);
        const float c = __fmul_rn((float)cls[i], p.max_wh);                        // class offset: boxes of different classes never overlap
        box[r] = make_float4(__fadd_rn(__fsub_rn(cx, hw), c), __fadd_rn(__fsub_rn(cy, hh), c), __fadd_rn(__fadd_rn(cx, hw), c),
                             __fadd_rn(__fadd_rn(cy, hh), c));
        supp[r] = 0;
    }
    __syncthreads();
    // ---- greedy suppression in score order (torchvision nms: float32 IoU, strict '>')
    int kept = 0;
    const int cap = min(p.max_det, 1024);
    for (int i = 0; i < n && kept < cap; ++i) {
        if (supp[i]) continue;                         // block-uniform: the flags only change between barriers
        if (tid == 0) s_kept[kept] = i;
        ++kept;
        const float4 bi = box[i];
        const float ai = __fmul_rn(bi.z - bi.x, bi.w - bi.y);
        for (int j = i + 1 + tid; j < n; j += kNmsThreads) {
            if (supp[j]) continue;
            const float4 bj = box[j];
            const float w = fmaxf(0.f, fminf(bi.z, bj.z) - fmaxf(bi.x, bj.x)), h = fmaxf(0.f, fminf(bi.w, bj.w) - fmaxf(bi.y, bj.y));
            const float inter = __fmul_rn(w, h);
            const float aj = __fmul_rn(bj.z - bj.x, bj.w - bj.y);
            if (__fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter)) > p.iou_thres) supp[j] = 1;
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) p.count[xi] = kept;
    // ---- kept rows (un-offset boxes, confidence, class) and their payload
    for (int r = tid; r < kept; r += kNmsThreads) {
        const int i = order[s_kept[r]], a = cand[i];
        const float cx = __ldg(pr + a), cy = __ldg(pr + (size_t)A + a);
        const float hw = __fdiv_rn(__ldg(pr + (size_t)2 * A + a), 2.f), hh = __fdiv_rn(__ldg(pr + (size_t)3 * A + a), 2.f);
        float* d = p.det + ((size_t)xi * p.max_det + r) * 6;
        d[0] = __fsub_rn(cx, hw); d[1] = __fsub_rn(cy, hh); d[2] = __fadd_rn(cx, hw); d[3] = __fadd_rn(cy, hh);
        d[4] = conf[i]; d[5] = (float)cls[i];
        p.out_anchor[(size_t)xi * p.max_det + r] = a;
        if (p.strides) p.out_strides[(size_t)xi * p.max_det + r] = __ldg(p.strides + a);
    }
    if (p.extra)
        for (int e = tid; e < kept * p.ne; e += kNmsThreads) {
            const int r = e / p.ne, c = e - r * p.ne;
            const int a = cand[order[s_kept[r]]];
            p.out_extra[((size_t)xi * p.max_det + r) * p.ne + c] = __ldg(p.extra + ((size_t)xi * p.ne + c) * A + a);
        }
}

}  // namespace oodb200

using namespace oodb200;

extern "C" int64_t oodb200_nms_workspace_bytes(int bs, int n_anchors) {
    if (bs < 0 || n_anchors < 0) return -1;
    return (int64_t)bs * n_anchors * (4 + 4 + 4 + 4 + 16 + 1) + 256;
}

extern "C" int oodb200_nms_payload_f32(const float* prediction, const float* extra_item, const float* strides, int bs, int nc,
                                       int n_extra, int n_anchors, float conf_thres, float iou_thres, float max_wh, int max_det,
                                       int max_nms, float* det, float* out_extra, float* out_strides, int32_t* out_anchor,
                                       int32_t* count, void* workspace, int64_t workspace_bytes, void* stream) {
    OODB200_REQUIRE(bs >= 0 && nc > 0 && n_anchors >= 0 && n_extra >= 0, "nms: bad size");
    OODB200_REQUIRE(max_det > 0 && max_det <= 1024 && max_nms > 0, "nms: max_det must be in 1..1024, max_nms positive");
    if (bs == 0) return OODB200_OK;
    OODB200_REQUIRE(prediction && det && out_anchor && count, "nms: null pointer");
    OODB200_REQUIRE(!extra_item || out_extra, "nms: extra_item needs out_extra");
    OODB200_REQUIRE(!strides || out_strides, "nms: strides needs out_strides");
    OODB200_REQUIRE(workspace && workspace_bytes >= oodb200_nms_workspace_bytes(bs, n_anchors), "nms: workspace too small");
    OODB200_REQUIRE(((uintptr_t)workspace & 15) == 0, "nms: workspace must be 16-byte aligned");
    const size_t na = (size_t)bs * n_anchors;
    char* ws = (char*)workspace;
    NmsParams p = {};
    p.pred = prediction; p.extra = extra_item; p.strides = strides;
    p.bs = bs; p.nc = nc; p.ne = n_extra; p.A = n_anchors;
    p.conf_thres = conf_thres; p.iou_thres = iou_thres; p.max_wh = max_wh; p.max_det = max_det; p.max_nms = max_nms;
    p.box = (float4*)ws; ws += 16 * na;
    p.cand = (int32_t*)ws; ws += 4 * na;
    p.conf = (float*)ws; ws += 4 * na;
    p.cls = (int32_t*)ws; ws += 4 * na;
    p.order = (int32_t*)ws; ws += 4 * na;
    p.supp = (uint8_t*)ws;
    p.det = det; p.out_extra = out_extra; p.out_strides = out_strides; p.out_anchor = out_anchor; p.count = count;
    nms_kernel<<<bs, kNmsThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch("nms");
}
