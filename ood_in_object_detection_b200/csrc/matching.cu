// Fit-data collection: which predictions are valid in-distribution samples, sm_100a.
//
// Replaces /root/reference/ood_utils.py:233-292 (`OODMethod.match_predicted_boxes_to_targets`): per image
//   score[p, g] = box_iou(pred_p, gt_g) * (cls_p == cls_g)          torchvision box_iou in float32 + the O(P*G) python mask loop (:251-257)
//   (rows, cols) = scipy.optimize.linear_sum_assignment(score, maximize=True)                                              (:283)
//   valid_preds = [i for i, c in enumerate(cols) if score[i, c] > iou_threshold]                                    (:288-291, quirk Q8)
// One CTA per image.  The assignment is scipy's own algorithm (rectangular_lsap.cpp: shortest augmenting paths, Crouse 2016)
// with its scan order and tie rules, because IoU x mask matrices are mostly zeros and the result depends on them: the frontier
// is scanned in the order of a `remaining` list initialised in reverse column order; among the minimal reduced costs the LAST
// unassigned column of the scan wins, else the FIRST column.  The scan over the remaining columns runs across the CTA's
// threads (every column's relaxation is independent), the selection is a deterministic two-key reduction, and the dual
// updates / augmentation are short serial steps.  All arithmetic in float64 like scipy's.
// Sizes: max(P, G) <= 1024 per image (YOLO's max_det is 300); the state lives in shared memory.
#include "common.cuh"

#include <float.h>
#include <math.h>

namespace oodb200 {

constexpr int kMatchThreads = 128;
constexpr int kMatchMax = 1024;                        // max(P, G) per image

struct MatchParams {
    const float* pred;          // [n, 4] xyxy
    const int32_t* pred_cls;    // [n]
    const int32_t* pred_start;  // [n_img + 1]
    const float* gt;            // [m, 4] xyxy
    const int32_t* gt_cls;      // [m]
    const int32_t* gt_start;    // [n_img + 1]
    const int64_t* score_off;   // [n_img + 1] element offset of every image's [P, G] score matrix
    float thr;
    int compat;                 // quirk Q8: test score[position, col] instead of score[row, col]
    float* score;               // [sum P*G]
    int32_t* row_ind;           // [n] per image: the first min(P, G) entries are used, rest -1
    int32_t* col_ind;           // [n]
    uint8_t* valid;             // [n] 1 = valid prediction (index as the reference stores it in valid_preds)
    int32_t* status;            // [1] set to 1 when an image exceeds kMatchMax or a score is not finite
};

__global__ void __launch_bounds__(kMatchThreads) match_kernel(const MatchParams p) {
    __shared__ double s_u[kMatchMax], s_v[kMatchMax], s_spc[kMatchMax];
    __shared__ int s_path[kMatchMax], s_col4row[kMatchMax], s_row4col[kMatchMax], s_rem[kMatchMax];
    __shared__ unsigned char s_SR[kMatchMax], s_SC[kMatchMax];
    __shared__ double s_rmin[kMatchThreads / 32];
    __shared__ int s_rlast[kMatchThreads / 32], s_rfirst[kMatchThreads / 32];
    __shared__ int s_sel[4];                           // index chosen, i (current row), sink, num_remaining
    __shared__ double s_minval;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p0 = p.pred_start[img], P = p.pred_start[img + 1] - p0;
    const int g0 = p.gt_start[img], G = p.gt_start[img + 1] - g0;
    for (int i = tid; i < P; i += kMatchThreads) { p.row_ind[p0 + i] = -1; p.col_ind[p0 + i] = -1; p.valid[p0 + i] = 0; }
    if (P == 0 || G == 0) return;
    if (P > kMatchMax || G > kMatchMax) { if (tid == 0) *p.status = 1; return; }
    float* __restrict__ score = p.score + p.score_off[img];
    // ---- score = IoU x same-class mask (float32, torchvision box_iou arithmetic)
    for (int e = tid; e < P * G; e += kMatchThreads) {
        const int i = e / G, j = e - i * G;
        const float4 a = *reinterpret_cast<const float4*>(p.pred + 4 * (size_t)(p0 + i));
        const float4 b = *reinterpret_cast<const float4*>(p.gt + 4 * (size_t)(g0 + j));
        const float area_a = __fmul_rn(a.z - a.x, a.w - a.y), area_b = __fmul_rn(b.z - b.x, b.w - b.y);
        const float w = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.f), h = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.f);
        const float inter = __fmul_rn(w, h);
        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
        const float s = __fmul_rn(iou, p.pred_cls[p0 + i] == p.gt_cls[g0 + j] ? 1.f : 0.f);
        if (!isfinite(s)) *p.status = 1;
        score[e] = s;
    }
    __syncthreads();
    // ---- scipy's solver on cost = -score; rows = the smaller side (transposed when G < P)
    const bool tr = G < P;
    const int nr = tr ? G : P, nc = tr ? P : G;
    auto cost = [&](int i, int j) -> double { return -(double)(tr ? score[j * G + i] : score[i * G + j]); };
    for (int i = tid; i < nr; i += kMatchThreads) { s_u[i] = 0.0; s_col4row[i] = -1; }
    for (int j = tid; j < nc; j += kMatchThreads) { s_v[j] = 0.0; s_row4col[j] = -1; s_path[j] = -1; }
    __syncthreads();
    for (int cur = 0; cur < nr; ++cur) {
        for (int j = tid; j < nc; j += kMatchThreads) { s_rem[j] = nc - j - 1; s_spc[j] = INFINITY; s_SC[j] = 0; }
        for (int i = tid; i < nr; i += kMatchThreads) s_SR[i] = 0;
        if (tid == 0) { s_sel[1] = cur; s_sel[2] = -1; s_sel[3] = nc; s_minval = 0.0; }
        __syncthreads();
        while (s_sel[2] == -1) {
            const int i = s_sel[1], nrem = s_sel[3];
            const double min_val = s_minval, ui = s_u[i];
            // relax the remaining columns from row i; per thread: min value, last unassigned / first position among its minima
            double lo = INFINITY;
            int last_un = -1, first = INT_MAX;
            for (int it = tid; it < nrem; it += kMatchThreads) {
                const int j = s_rem[it];
                const double r = min_val + cost(i, j) - ui - s_v[j];
                double sj = s_spc[j];
                if (r < sj) { s_path[j] = i; s_spc[j] = r; sj = r; }
                if (sj < lo) { lo = sj; last_un = s_row4col[j] == -1 ? it : -1; first = it; }
                else if (sj == lo) { if (s_row4col[j] == -1) last_un = it; }           // `it` increases: last unassigned so far
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double olo = __shfl_xor_sync(0xffffffffu, lo, o);
                const int olast = __shfl_xor_sync(0xffffffffu, last_un, o), ofirst = __shfl_xor_sync(0xffffffffu, first, o);
                if (olo < lo) { lo = olo; last_un = olast; first = ofirst; }
                else if (olo == lo) { last_un = max(last_un, olast); first = min(first, ofirst); }
            }
            if (lane == 0) { s_rmin[warp] = lo; s_rlast[warp] = last_un; s_rfirst[warp] = first; }
            if (tid == 0) s_SR[i] = 1;
            __syncthreads();
            if (tid == 0) {
                double m = s_rmin[0];
                int lu = s_rlast[0], fi = s_rfirst[0];
                for (int w = 1; w < kMatchThreads / 32; ++w) {
                    if (s_rmin[w] < m) { m = s_rmin[w]; lu = s_rlast[w]; fi = s_rfirst[w]; }
                    else if (s_rmin[w] == m) { lu = max(lu, s_rlast[w]); fi = min(fi, s_rfirst[w]); }
                }
                // sequential rule of the scan: among the minimal entries the LAST unassigned column wins, else the FIRST entry
                const int index = lu >= 0 ? lu : fi;
                s_minval = m;
                const int j = s_rem[index];
                if (s_row4col[j] == -1) s_sel[2] = j; else s_sel[1] = s_row4col[j];
                s_SC[j] = 1;
                s_rem[index] = s_rem[nrem - 1];
                s_sel[3] = nrem - 1;
            }
            __syncthreads();
        }
        // dual update and augmentation
        const double min_val = s_minval;
        for (int i = tid; i < nr; i += kMatchThreads)
            if (i == cur) s_u[i] += min_val;
            else if (s_SR[i]) s_u[i] += min_val - s_spc[s_col4row[i]];
        for (int j = tid; j < nc; j += kMatchThreads)
            if (s_SC[j]) s_v[j] -= min_val - s_spc[j];
        __syncthreads();
        if (tid == 0) {
            int j = s_sel[2];
            for (;;) {
                const int i = s_path[j];
                s_row4col[j] = i;
                const int nj = s_col4row[i];
                s_col4row[i] = j;
                j = nj;
                if (i == cur) break;
            }
        }
        __syncthreads();
    }
    // ---- (row_ind, col_ind) with rows ascending; valid predictions (ood_utils.py:288-291)
    for (int i = tid; i < nr; i += kMatchThreads) {
        int pos = i, row = i, col = s_col4row[i];
        if (tr) {                                      // solver rows are ground-truth boxes: order the pairs by prediction index
            row = s_col4row[i];
            col = i;
            pos = 0;
            for (int e = 0; e < nr; ++e) pos += s_col4row[e] < row;
        }
        p.row_ind[p0 + pos] = row;
        p.col_ind[p0 + pos] = col;
        const int rr = p.compat ? pos : row;           // Q8: the reference indexes with the position in the assignment
        if (score[rr * G + col] > p.thr) p.valid[p0 + rr] = 1;
    }
}

}  // namespace oodb200

using namespace oodb200;

extern "C" int oodb200_match_boxes_f32(const float* pred_xyxy, const int32_t* pred_cls, const int32_t* pred_start,
                                       const float* gt_xyxy, const int32_t* gt_cls, const int32_t* gt_start,
                                       const int64_t* score_off, int n_img, float iou_threshold, int compat,
                                       float* score, int32_t* row_ind, int32_t* col_ind, uint8_t* valid, int32_t* status,
                                       void* stream) {
    OODB200_REQUIRE(n_img >= 0, "match_boxes: negative n_img");
    if (n_img == 0) return OODB200_OK;
    OODB200_REQUIRE(pred_xyxy && pred_cls && pred_start && gt_xyxy && gt_cls && gt_start && score_off && score && row_ind &&
                    col_ind && valid && status, "match_boxes: null pointer");
    MatchParams p = {pred_xyxy, pred_cls, pred_start, gt_xyxy, gt_cls, gt_start, score_off, iou_threshold, compat, score,
                     row_ind, col_ind, valid, status};
    match_kernel<<<n_img, kMatchThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch("match_boxes");
}
