/* oodb200.h -- C ABI of the B200-native OoD-scoring hot path (liboodb200.so).
 *
 * The reference (aitor-martinez-seras/OoD_in_Object_Detection) is pure Python and has no
 * FFI; the boundary it offers is the `OODMethod` class surface of `ood_utils.py`.  The
 * entry points below are what a binding for that path would call: each one replaces the
 * *inside* of the reference function cited next to it.  The Python classes in
 * `ood_in_object_detection_b200/ood_utils.py` (same names / arguments / error behaviour as the
 * reference's) call these through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - every function only enqueues work on `stream` (a cudaStream_t passed as void*), never
 *     synchronises, never allocates, and returns 0 or a negative OODB200_ERR_* code;
 *     `oodb200_last_error()` gives the message for the calling thread;
 *   - no torch types anywhere: plain pointers and sizes.
 *   - 1 = in-distribution, 0 = out-of-distribution, as in /root/reference/ood_utils.py:148-158.
 */
#ifndef OODB200_H_
#define OODB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OODB200_ABI_VERSION 2

#define OODB200_OK 0
#define OODB200_ERR_INVALID (-1)  /* bad argument (null pointer, size out of range, ...) */
#define OODB200_ERR_CUDA (-2)     /* CUDA runtime reported an error at launch            */

/* metric slots / bit masks (reference: `metric` of the DistanceMethod subclasses,
 * /root/reference/ood_utils.py:2574-2595) */
#define OODB200_METRIC_L1 0
#define OODB200_METRIC_L2 1
#define OODB200_METRIC_COS 2
#define OODB200_N_METRICS 3

/* logit-method slots (reference: /root/reference/ood_utils.py:1388-1443; MaxLogit has no
 * counterpart there) */
#define OODB200_LOGIT_MSP 0
#define OODB200_LOGIT_ENERGY 1
#define OODB200_LOGIT_ODIN 2
#define OODB200_LOGIT_SIGMOID 3
#define OODB200_LOGIT_MAXLOGIT 4
#define OODB200_N_LOGIT 5
/* flag bit OR-ed into method_mask: the inputs are post-sigmoid probabilities, the Sigmoid slot returns logits[cls] as is
 * (`use_values_before_sigmoid=False`, /root/reference/ood_utils.py:1438-1439) */
#define OODB200_LOGIT_FLAG_POST_SIGMOID 0x100

/* fusion strategies (reference: FusionMethod.fuse_ood_decisions :2906-2940,
 * TripleFusionMethod.fuse_ood_decisions :3282-3301) */
#define OODB200_FUSE_AND 0      /* elementwise max  */
#define OODB200_FUSE_OR 1       /* elementwise min  */
#define OODB200_FUSE_MAJORITY 2 /* a+b+c >= 2       */

int oodb200_abi_version(void);
const char* oodb200_last_error(void);

/* ---- K1: per-detection RoIAlign pooling --------------------------------------------------
 * Replaces `extract_roi_aligned_features_from_correct_stride`
 * (/root/reference/ultralytics/models/yolo/detect/predict.py:13-90), i.e. torchvision
 * roi_align(output_size=(1,1), sampling_ratio=-1, aligned=False) of every box on the map of its
 * own stride.
 *   map_ptrs   [n_img*3] device array of device pointers; map_ptrs[img*3+s] -> float32 [C_s,H_s,W_s]
 *              contiguous (the per-image CHW tensors `Results.extra_item[0]` holds, no copy)
 *   map_chw    host int32[9]: C,H,W of stride 0,1,2
 *   scale      host float[3]: spatial_scale per stride = (float)(W_s / img_w)  (predict.py:68)
 *   boxes      [n,4] xyxy float32 in input-image pixels; img_idx/stride_idx [n] int32 (boxes of one image
 *              are contiguous); img_start [n_img+1] int32 prefix of boxes per image
 *   out        [n, out_ld] float32; row i gets C_{stride_idx[i]} values (rest untouched)
 * A box whose stride_idx is outside {0,1,2} is skipped (the reference never pools it either).
 * Maps that are 16-byte aligned with W_s % 4 == 0 (every torch-allocated YOLO map) take the 128-bit gather; anything
 * else falls back to a scalar gather with the same results.  At most 2^24 - 1 boxes per call.
 */
int oodb200_roi_pool_f32(const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                         const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                         const int32_t* img_start, int n,
                         float* out, int out_ld, void* workspace, int64_t workspace_bytes, void* stream);

/* Scratch the two pooling entry points need (device memory, 256-byte aligned, owned by the caller; its
 * contents need not be preserved between calls): Q1 plan, per-box window geometry and RoIAlign weights,
 * the item work list, the (stride, class) ordering of the score kernel and the pooled vectors [n, round4(Cmax)].
 * map_chw as in roi_pool (host); nc = number of classes of the centroid table (0 for roi_pool).
 * Returns -1 when map_chw is NULL. */
int64_t oodb200_fmap_workspace_bytes(int n, int nc, const int32_t* map_chw);

/* ---- K1+K2 fused: pool -> L2-normalise -> distance to the class/stride centroids -> min ->
 *      threshold.  Replaces the per-image/per-stride/per-box loop of
 * `DistanceMethod.compute_ood_decision_on_results` + `_compute_ood_decision_for_one_result_...`
 * (/root/reference/ood_utils.py:2038-2180), `activations_transformation` (:2404-2409) and
 * `compute_distance` (:2422-2430).
 *   cls          [n] predicted class of every box (`res.boxes.cls`)
 *   img_start    [n_img+1] int32 prefix of boxes per image
 *   compat_q1    1 = the reference's behaviour (ood_utils.py:2152-2154): the class used for
 *                the centroid / threshold lookup is the one of the box with the same IN-STRIDE index, and
 *                results are written stride-major within each image.  0: class of the box itself, box order.
 *   metric_mask  OR of (1<<OODB200_METRIC_*): every requested metric is scored in the same pass
 *   normalize    1 = L2-normalise the pooled vector first (vanilla FMap methods); 0 = score as is
 *   cent         packed float32 centroids; (stride s, class c) has cent_k[s*nc+c] rows of C_s floats
 *                starting at element cent_off[s*nc+c]; cent_unit = same layout, rows L2-normalised
 *                (needed for cosine only, may be NULL otherwise)
 *   thr          [3 metrics][3*nc] float64, NaN = "no threshold" ([] / 0 / 0.0 in the reference,
 *                ood_utils.py:2173) -> OoD
 *   dist/argmin/decision   [3][n] (slot = metric id); only requested slots are written.
 *                No cluster: dist 1000, argmin -1 (ood_utils.py:2159-2164).
 *   pooled       optional [n, pooled_ld] raw pooled vectors (NULL to skip), rows in OUTPUT order
 *   cls_used_out / out_index_out   optional [n] int32: class used and output slot of every box
 *   workspace    see oodb200_fmap_workspace_bytes
 */
int oodb200_fmap_score_f32(const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                           const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                           const int32_t* cls, const int32_t* img_start, int compat_q1, int n,
                           int metric_mask, int normalize,
                           const float* cent, const float* cent_unit, const int64_t* cent_off, const int32_t* cent_k,
                           int nc, const double* thr,
                           float* dist, int32_t* argmin, uint8_t* decision,
                           float* pooled, int pooled_ld, int32_t* cls_used_out, int32_t* out_index_out,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* ---- K1 / K1+K2 on CHANNELS-LAST maps (what a detector run in torch.channels_last hands out): same arguments and
 * results as oodb200_roi_pool_f32 / oodb200_fmap_score_f32, but map_ptrs[i*3+s] points at a [H_s, W_s, C_s] array
 * (element (c, y, x) at [(y * W_s + x) * C_s + c]).  A window cell is then C_s contiguous floats: every fetched
 * 128-byte line is used in full and pooling needs no cross-lane reduction. */
int oodb200_roi_pool_nhwc_f32(const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                         const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                         const int32_t* img_start, int n,
                         float* out, int out_ld, void* workspace, int64_t workspace_bytes, void* stream);
int oodb200_fmap_score_nhwc_f32(const float* const* map_ptrs, const int32_t* map_chw, const float* scale, int n_img,
                           const float* boxes, const int32_t* img_idx, const int32_t* stride_idx,
                           const int32_t* cls, const int32_t* img_start, int compat_q1, int n,
                           int metric_mask, int normalize,
                           const float* cent, const float* cent_unit, const int64_t* cent_off, const int32_t* cent_k,
                           int nc, const double* thr,
                           float* dist, int32_t* argmin, uint8_t* decision,
                           float* pooled, int pooled_ld, int32_t* cls_used_out, int32_t* out_index_out,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* ---- Q1 plan (standalone; oodb200_fmap_score_f32 does the same internally when img_start != NULL):
 * the reference looks the class up with the in-stride index and emits decisions
 * stride-major (/root/reference/ood_utils.py:2152-2154).  For box b of an image
 * (local index b, stride s, j = number of earlier boxes of the same stride):
 *   cls_used[b] = cls[img_start + j]      out_index[b] = img_start + #boxes(stride < s) + j
 *   img_start  [n_img+1] int32 prefix of boxes per image
 */
int oodb200_q1_plan_i32(const int32_t* img_start, const int32_t* stride_idx, const int32_t* cls, int n_img,
                        int32_t* cls_used, int32_t* out_index, void* stream);

/* ---- fit-data collection: which predictions are valid in-distribution samples.  Replaces
 * `OODMethod.match_predicted_boxes_to_targets` (/root/reference/ood_utils.py:233-292): per image the IoU (torchvision
 * `box_iou`, float32) x same-class mask of every prediction against every ground-truth box (:251-257),
 * `scipy.optimize.linear_sum_assignment(score, maximize=True)` (:283; scipy's shortest-augmenting-path algorithm with its
 * scan order and tie rules, float64) and the walk over the assignment (:288-291).  One CTA per image; max(P, G) <= 1024.
 *   pred_xyxy [n, 4] / pred_cls [n] / pred_start [n_img+1]    predictions, images back to back
 *   gt_xyxy [m, 4] / gt_cls [m] / gt_start [n_img+1]          ground truth in absolute pixels (create_targets_dict, :201-231)
 *   score_off [n_img+1] int64   element offset of every image's [P, G] row-major score matrix inside `score`
 *   compat    1 = the reference's quirk Q8: the walk tests score[POSITION in the assignment, col] and records the position
 *             (identical to the assigned row when P <= G); 0 = tests the assigned (row, col) pair
 *   row_ind / col_ind [n]   the assignment of every image, rows ascending, first min(P, G) entries, rest -1
 *   valid [n] u8            1 at the indices the reference appends to `valid_preds`
 *   status [1] int32        must be 0 on entry; set to 1 when an image exceeds 1024 boxes or a score is not finite
 */
int oodb200_match_boxes_f32(const float* pred_xyxy, const int32_t* pred_cls, const int32_t* pred_start,
                            const float* gt_xyxy, const int32_t* gt_cls, const int32_t* gt_start,
                            const int64_t* score_off, int n_img, float iou_threshold, int compat,
                            float* score, int32_t* row_ind, int32_t* col_ind, uint8_t* valid, int32_t* status,
                            void* stream);

/* ---- NMS with the OoD payload.  Replaces the default path of
 * `non_max_suppression_old` (/root/reference/ultralytics/utils/ops.py:348-530): candidates by best class confidence
 * (:412, :463-466), xywh -> xyxy (:455-456), descending confidence order (:478-482), class-aware greedy NMS with boxes offset
 * by cls * max_wh (torchvision nms semantics: float32 IoU, strict '>'; :485-489), max_det (:490), and the gather of the
 * per-anchor payload rows `extra_item` (raw class logits) and `strides` with the same indices (:433-436, :479-481, :503-506).
 * One CTA per image.
 *   prediction [bs, 4 + nc, A] float32 (cx, cy, w, h, class confidences); extra_item [bs, n_extra, A] or NULL; strides [A] or NULL
 *   det [bs, max_det, 6] (xyxy, confidence, class), out_extra [bs, max_det, n_extra], out_strides [bs, max_det],
 *   out_anchor [bs, max_det] anchor index of every kept detection, count [bs] detections kept per image (max_det <= 1024)
 *   workspace: oodb200_nms_workspace_bytes(bs, A) bytes, 16-byte aligned
 */
int64_t oodb200_nms_workspace_bytes(int bs, int n_anchors);
int oodb200_nms_payload_f32(const float* prediction, const float* extra_item, const float* strides, int bs, int nc,
                            int n_extra, int n_anchors, float conf_thres, float iou_thres, float max_wh, int max_det,
                            int max_nms, float* det, float* out_extra, float* out_strides, int32_t* out_anchor,
                            int32_t* count, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- K3: logit methods.  Replaces `LogitsMethod.compute_ood_decision_on_results` /
 * `compute_INDness_scores_on_results` (/root/reference/ood_utils.py:1195-1257) and the scorers
 * (:1388-1443).  One pass computes every method in method_mask.
 *   logits [n, nc] float32 raw (pre-sigmoid) class logits; cls [n] int32
 *   thr/smin/smax [5][nc] float64 (per-class threshold, min, max InD score); may be NULL -> only scores
 *   scores [5][n] f32, indness [5][n] f32 (NULL to skip), decision [5][n] u8 (0 if score < thr else 1)
 *   sigmoid_mismatch  optional int32[1] counter of boxes whose class is not the arg-max logit
 *                     (the reference asserts on it, :1442)
 */
int oodb200_logit_score_f32(const float* logits, const int32_t* cls, int n, int nc, int method_mask,
                            float t_energy, float t_odin, const double* thr, const double* smin, const double* smax,
                            int clip_indness, float* scores, float* indness, uint8_t* decision,
                            int32_t* sigmoid_mismatch, void* stream);

/* ---- K6: fusion rules, position-wise (/root/reference/ood_utils.py:2906-2940, 3282-3301). */
int oodb200_fuse_u8(const uint8_t* a, const uint8_t* b, const uint8_t* c, int n, int strategy, uint8_t* out, void* stream);
int oodb200_fuse_score_f32(const float* s1, const float* s2, int n, uint8_t* out, void* stream);

/* ---- K2 standalone on already-pooled vectors (fit-time scoring), segmented.
 * Replaces `compute_scores_one_class_one_stride` (/root/reference/ood_utils.py:2000-2036):
 * normalize + pairwise_distances(...).min(axis=0) per (class, stride) segment.
 *   x [n_rows, ld] float32 (first `dim` columns used); seg_off device int64 [n_seg+1] row offsets;
 *   segment g is scored against rows cent_row_off[g] .. +cent_k[g] of cent [*, dim]
 *   dist / argmin [3][n_rows] (slot = metric id; only requested slots written); a segment without centroids
 *   (cent_k == 0) gives 1000 / -1 (ood_utils.py:2159-2164)
 *   thr [3][n_seg] float64 (NaN = no threshold) and decision [3][n_rows] are optional (both NULL to skip):
 *   decision = dist < thr (ood_utils.py:2173-2180)
 */
int oodb200_vec_score_f32(const float* x, int64_t ld, int dim, const int64_t* seg_off, int n_seg, int64_t n_rows,
                          int metric_mask, int normalize,
                          const float* cent, const float* cent_unit, const int64_t* cent_row_off, const int32_t* cent_k,
                          float* dist, int32_t* argmin, const double* thr, uint8_t* decision, void* stream);

/* INDness of a distance score as the reference intends it (/root/reference/ood_utils.py:1599-1604): piecewise linear
 * through (min_dist, +1), (thr, 0), (max_dist, -1), clipped to [-1, 1] when clip != 0.  slot[i] indexes the
 * thr / dmin / dmax tables (float64; NaN thr or slot < 0 -> -1).  The shipped reference always returns -1
 * (SURVEY.md Q2); the host class does the same in compat mode without calling this. */
int oodb200_dist_indness_f32(const float* dist, const int32_t* slot, int64_t n, const double* thr, const double* dmin,
                             const double* dmax, int clip, float* out, void* stream);

/* Row-wise L2 normalisation, `sklearn.preprocessing.normalize(x, axis=1)` on float32 rows
 * (`DistanceMethod.activations_transformation`, /root/reference/ood_utils.py:2404-2409). out may alias x. */
int oodb200_normalize_rows_f32(const float* x, int64_t ld, int dim, int64_t n_rows, float* out, int64_t out_ld, void* stream);

/* ---- K5: one pass of an exact radix select over float32 score bit patterns, segmented.
 * Replaces `np.percentile(scores, q, method='lower')` (/root/reference/ood_utils.py:613,626): the host
 * computes the order-statistic index numpy would use, then narrows 11+11+10 key bits with three passes.
 *   hist [n_seg, 1<<bits] uint32 is ACCUMULATED INTO (caller zeroes); a key is counted when its bits above
 *   (shift+bits) equal prefix[g] (no condition when shift+bits == 32); bin = (key >> shift) & ((1<<bits)-1).
 *   key = order-preserving transform of the float32 bits (negative: ~b, else b | 0x80000000).
 *   minmax optional [n_seg][2] uint32 keys, min/max ACCUMULATED with integer atomics (caller sets 0xFFFFFFFF / 0).
 */
int oodb200_radix_hist_u32(const float* scores, const int64_t* seg_off, int n_seg, int64_t n_rows,
                           const uint32_t* prefix, int shift, int bits, uint32_t* hist, uint32_t* minmax, void* stream);

/* ---- K4: k-means (Lloyd) over segmented data: every (class, stride) problem advances in the same launches.
 * Replaces sklearn `lloyd_iter_chunked_dense` / `_average_centers` / `_center_shift` behind
 * `KMeans(n_clusters=k, random_state=10).fit_predict(X)` (/root/reference/cluster_utils.py:62-73).
 *
 * kmeans_step: one CTA per block of rows (block b covers rows [block_row0[b], block_row1[b]) of segment
 *   block_seg[b]; blocks never straddle segments).  labels[r] = argmin_k ||c_k||^2 - 2 x_r.c_k (first minimum);
 *   n_changed[g] += rows whose label changed (labels is read as the previous assignment; -1 initially).
 *   update == 2: labels are NOT recomputed; the sums/counts of the given labels are produced (member means).
 *   update != 0: psums [n_blocks, k, dim] / pcounts [n_blocks, k] receive the block's per-cluster sums and
 *   counts, accumulated in row order (bit-reproducible).  active [n_seg] (or NULL): segments with 0 are skipped.
 *   x [n_rows, dim] float32, already mean-centred per segment by the caller; cent [n_seg, k, dim]; seg_k [n_seg].
 * kmeans_reduce: out[grp] = sum of in[b] for b in [first[grp], first[grp+1]) in increasing b (fixed order).
 * kmeans_update: centres = sums * (1/count) (sklearn arithmetic), empty clusters keep the old centre and are
 *   counted in n_empty[g] (written, not accumulated); shift_sq[g] = sum_k ||new_k - old_k||^2.  An inactive segment
 *   copies cent_old to cent_new (shift 0), so the caller may ping-pong two centre buffers.
 * sqdist_cand (k-means++ seeding): out_d[j, r] = min(closest[r], d(x_r, cand[g,j,:])) with d evaluated like
 *   sklearn's float64 expansion cast to float32; pot[g, j] += sum_r out_d[j, r] (float64).
 */
int64_t oodb200_kmeans_smem_bytes(int k, int dim);
int oodb200_kmeans_step_f32(const float* x, int dim, int n_seg, int k, const int32_t* seg_k, const float* cent,
                            const int32_t* block_seg, const int64_t* block_row0, const int64_t* block_row1, int n_blocks,
                            const int32_t* active, int32_t* labels, float* psums, float* pcounts, int32_t* n_changed,
                            int update, void* stream);
int oodb200_kmeans_reduce_f32(const float* in, const int32_t* first, int n_groups, int64_t elems, float* out, void* stream);
/* kmeans_reduce_step: the two reductions of an iteration (sums [n_blocks, elems_s] and counts [n_blocks, elems_c] ->
 * out_s / out_c per group, fixed block order) in one launch; also chg_f[g] = (float)n_changed[g], n_changed[g] = 0. */
int oodb200_kmeans_reduce_step_f32(const float* psums, const float* pcounts, const int32_t* first, int n_groups,
                                   int64_t elems_s, int64_t elems_c, float* out_s, float* out_c, int32_t* n_changed,
                                   float* chg_f, void* stream);
int oodb200_kmeans_update_f32(const float* sums, const float* counts, const float* cent_old, const int32_t* seg_k,
                              const int32_t* active, int n_seg, int k, int dim, float* cent_new, float* shift_sq,
                              int32_t* n_empty, void* stream);

/* ---- the Lloyd iteration's all-reduce fused into the centre update (N > 1 ranks on one NVLink / NVSwitch node):
 * every rank has written its reduced partials [n_seg*k*dim sums | n_seg*k counts | n_seg changed-label counts] into a
 * SYMMETRIC buffer (torch.distributed._symmetric_memory); peer_bufs is the device array of the n_peers buffer addresses as
 * mapped into this process.  The kernel reads every peer's values straight over NVLink, adds them in rank order (every
 * rank computes identical bits) and does what oodb200_kmeans_update_f32 does; the summed counts and changed-label counts
 * are written to cnts_out [n_seg, k] / chg_out [n_seg] for oodb200_kmeans_converge_f32.
 * The barrier between the ranks is part of the kernel: peer_flags (device array of the n_peers addresses of a symmetric
 * uint32[n_peers] array, or NULL when the caller orders the ranks itself) -- rank my_rank stores `epoch` into slot my_rank of
 * every peer's array (release, system scope) and waits until every slot of its own array has reached `epoch`; epoch must
 * increase from call to call (wrap-around safe), the partial buffers alternate between iterations.
 * Replaces the NCCL all-reduce + update of `kmeans.kmeans_fit` (same sklearn statements as kmeans_update). */
int oodb200_kmeans_update_peers_f32(const float* const* peer_bufs, int n_peers, int64_t counts_off, int64_t chg_off,
                                    uint32_t* const* peer_flags, int my_rank, uint32_t epoch, const float* cent_old, const int32_t* seg_k, const int32_t* active, int n_seg,
                                    int k, int dim, float* cent_new, float* shift_sq, int32_t* n_empty,
                                    float* cnts_out, float* chg_out, void* stream);

/* ---- Lloyd convergence bookkeeping on the device: replaces the per-iteration host decisions of sklearn's
 * `_kmeans_single_lloyd` (`_kmeans.py:712-740`, behind /root/reference/cluster_utils.py:62-73) so that the host does
 * not read flags back every iteration.  For every ACTIVE segment: state[0][g] += 1 (iterations), state[3][g] +=
 * n_empty[g], counts[g] = cnts[g]; no label changed -> state[1][g] = 1 (strict), active[g] = 0; else squared centre
 * shift <= tol_abs[g] -> state[2][g] = 1 (final E-step needed), active[g] = 0.  any_active[0] = 1 iff a segment is
 * still active.  n_changed: int32 (n_changed_i) or float32 (n_changed_f, when it rode through the all-reduce buffer).
 *   shift [n_seg] f32, n_empty [n_seg] i32, tol_abs [n_seg] f64, cnts / counts [n_seg, k] f32, state [4, n_seg] i32
 */
int oodb200_kmeans_converge_f32(const int32_t* n_changed_i, const float* n_changed_f, const float* shift,
                                const int32_t* n_empty, const double* tol_abs, const float* cnts, int n_seg, int k,
                                int32_t* active, int32_t* state, float* counts, int32_t* any_active, void* stream);
int oodb200_sqdist_cand_f32(const float* x, int dim, const int64_t* seg_off, int n_seg, int64_t max_seg_rows,
                            const float* cand, int n_cand, const float* closest, float* out_d, double* pot,
                            void* stream);

/* ---- K2b: K2 standalone on the tensor pipe for up to 64 centroids per segment (BASELINE config C5: K = 64): the x.c
 * cross-term of 'l2' / 'cosine' runs on tcgen05 (kind::tf32, split-float operands), combined like sklearn
 * (`euclidean_distances`: XX - 2 X.Y^T + YY in float64, cast to float32, sqrt; `cosine_distances`: 1 - cos, clipped).
 * Replaces `pairwise_distances(cluster, activations, metric).min(axis=0)` (/root/reference/ood_utils.py:2422-2430).
 *   x [n_rows, dim] float32 contiguous, ALREADY normalised (normalize_rows) when the method normalises; metric =
 *   OODB200_METRIC_L2 or OODB200_METRIC_COS (for cosine pass the unit-norm centroid rows as `cent`);
 *   blocks: block b = rows [block_row0[b], block_row1[b]) of segment block_seg[b], at most 512 rows, never straddling
 *   segments; dist / argmin / decision / thr laid out like vec_score ([3][n_rows], [3][n_seg]); only slot `metric` is written.
 *   workspace: oodb200_vec_score_tc_workspace_bytes (0 = shape not supported), 256-byte aligned. */
int64_t oodb200_vec_score_tc_workspace_bytes(int n_seg, int64_t n_rows, int dim);
int oodb200_vec_score_tc_f32(const float* x, int64_t n_rows, int dim, int n_seg, int metric, const float* cent,
                             const int64_t* cent_row_off, const int32_t* cent_k, int max_k, const int32_t* block_seg,
                             const int64_t* block_row0, const int64_t* block_row1, int n_blocks, float* dist,
                             int32_t* argmin, const double* thr, uint8_t* decision, void* workspace, void* stream);

/* Mean-centring of every segment and the variance behind sklearn's tolerance (KMeans.fit `X -= X.mean(axis=0)`,
 * `_tolerance`; sklearn/cluster/_kmeans.py:283-293, 1487-1497), float64 accumulation, fixed combination order.
 * segment_colsum: sums[g, d] = sum_r x[r, d] over the rows of segment g (the caller divides by the global size).
 * segment_center: out[r, d] = x[r, d] - mean[g, d]; sq[g] = sum of out^2 over the segment.
 * scratch: oodb200_segment_scratch_doubles(n_seg, dim) doubles. */
int64_t oodb200_segment_scratch_doubles(int n_seg, int dim);
int oodb200_segment_colsum_f64(const float* x, int dim, const int64_t* seg_off, int n_seg, double* scratch, double* sums,
                               void* stream);
int oodb200_segment_center_f32(const float* x, int dim, const int64_t* seg_off, int n_seg, const float* mean, float* out,
                               double* scratch, double* sq, void* stream);

/* ---- K4 on the tensor pipe: same contract as kmeans_step (update 0 or 1) for k <= 16, dim % 32 == 0,
 * 128 <= dim <= 640: the x.c cross-term runs on tcgen05 (kind::tf32, split-float hi/lo pieces = float32-level
 * accuracy) on 128-row tiles streamed by TMA; the partial sums re-read the tile from L2 (csrc/kmeans_tc.cu).
 * kmeans_tc_workspace_bytes: scratch the call needs (0 = shape not supported, use kmeans_step). */
int64_t oodb200_kmeans_tc_workspace_bytes(int n_seg, int k, int dim);
int oodb200_kmeans_step_tc_f32(const float* x, int64_t n_rows, int dim, int n_seg, int k, const int32_t* seg_k,
                               const float* cent, const int32_t* block_seg, const int64_t* block_row0,
                               const int64_t* block_row1, int n_blocks, const int32_t* active, int32_t* labels,
                               float* psums, float* pcounts, int32_t* n_changed, int update, void* workspace,
                               void* stream);

/* ---- K4b: k-means++ seeding on the device (sklearn `_kmeans_plusplus`, sklearn/cluster/_kmeans.py:180-278, behind
 * /root/reference/cluster_utils.py:62-73), every segment in lock-step, no host round trip per centre.
 * seed_sqdist: like sqdist_cand with n_cand <= 4, but the potentials are bit-reproducible: every block writes a
 *   float64 partial to pot_part [n_seg, oodb200_seed_grid(max_seg_rows), 4] and pots [n_seg, n_cand] receives their
 *   fixed-order sum (the caller adds the ranks).
 * seed_scan: cand_id[g, t] = min(searchsorted(cumsum_f32(closest of segment g in global row order),
 *   uniform[g, t] * pot[g]), n_g - 1) -- the cumsum is the sequential float32 sum numpy computes.  Segment g is the
 *   concatenation of n_pieces pieces closest_all[piece_off[g, r] .. + piece_cnt[g, r]) (one piece per rank).
 *   chunk_sum: scratch [n_seg, max_chunks] floats, max_chunks >= 32 * ceil(max segment rows / 4096).
 *   seg_trials [n_seg] (<= n_trials; remaining slots repeat candidate 0), seg_on [n_seg] (0 = skip) may be NULL.
 * seed_gather: vec[g, j, :] = x[seg_off[g] + cand_id[g, j] - shard_first[g]] when this rank owns that row, else 0.
 * seed_pick: best = first minimum of float32(pots[g, :trials]); closest[r] = newd[best, r] for the rows of g;
 *   pot[g] = float32(pots[g, best]); cent_out[g * cent_stride + d] = cand_vec[g, best, d].
 */
int oodb200_seed_grid(int64_t max_seg_rows);
int oodb200_seed_sqdist_f32(const float* x, int dim, const int64_t* seg_off, int n_seg, int64_t max_seg_rows,
                            const float* cand, int n_cand, const float* closest, float* out_d, double* pot_part,
                            double* pots, void* stream);
int oodb200_seed_scan_f32(const float* closest_all, const int64_t* piece_off, const int64_t* piece_cnt, int n_seg,
                          int n_pieces, const double* uniform, const float* pot, const int32_t* seg_trials,
                          const int32_t* seg_on, int n_trials, float* chunk_sum, int64_t max_chunks, int64_t* cand_id,
                          void* stream);
int oodb200_seed_gather_f32(const float* x, int dim, const int64_t* cand_id, int n_seg, int n_cand,
                            const int64_t* seg_off, const int64_t* shard_first, float* vec, void* stream);
int oodb200_seed_pick_f32(const double* pots, const int32_t* seg_trials, const int32_t* seg_on, int n_seg, int n_cand,
                          const int64_t* seg_off, int64_t max_seg_rows, const float* newd, const float* cand_vec, int dim,
                          float* closest, float* pot, float* cent_out, int64_t cent_stride, int32_t* best_out,
                          void* stream);

/* ---- K7: per-cluster distance sums for the silhouette score of the k-search
 * (/root/reference/cluster_utils.py:203-302; :277 `silhouette_score(feature_maps_one_run, cluster_labels, metric=metric)`
 * -> sklearn `silhouette_samples`: pairwise_distances_chunked + `_silhouette_reduce`).
 *   x [n, ld] float32 rows of ONE (class, stride) segment (first `d` columns used; for cosine the rows must already be
 *   unit-norm like sklearn `normalize` leaves them); labels int32 [n] in [0, kc)
 *   sums float64 [n, kc]:  sums[i][c] = sum_{j : labels[j] == c} dist(x_i, x_j), dist(i, i) = 0
 *   metric = OODB200_METRIC_L1 (cityblock) / _L2 (euclidean) / _COS (1 - x.y clipped to [0, 2])
 */
int oodb200_pair_cluster_sums_f32(const float* x, int n, int d, int64_t ld, const int32_t* labels, int kc, int metric,
                                  double* sums, void* stream);

/* The k-search (/root/reference/cluster_utils.py:203-302) scores up to 13 labelings of the SAME rows: the pair distances
 * can be stored once and folded per labeling.
 *   oodb200_pair_dist_matrix_f32:    dist [n, ld_dist] float32, dist[i][j] = dist(x_i, x_j) as above (exactly symmetric,
 *                                    zero diagonal); same x / metric contract as oodb200_pair_cluster_sums_f32
 *   oodb200_matrix_cluster_sums_f32: order int32 [n] = the rows sorted by label (stable: ascending row inside a label),
 *                                    member_off int64 [kc + 1] = start of every label's run in `order`;
 *                                    sums float64 [n, kc] as above, members added in ascending row order (deterministic)
 */
int oodb200_pair_dist_matrix_f32(const float* x, int n, int d, int64_t ld, int metric, float* dist, int64_t ld_dist,
                                 void* stream);
int oodb200_matrix_cluster_sums_f32(const float* dist, int n, int64_t ld_dist, const int32_t* order,
                                    const int64_t* member_off, int kc, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OODB200_H_ */
