/*
 * geom3d.h — C ABI of libgeom3d.so: the sm_100a (B200) kernels behind the 3D-box geometry hot path of
 * DerekGloudemans/3D-playground (anchor<->GT IoU assignment + focal / corner / direction losses, box decode,
 * score filter, NMS, homography state<->image projection, tracker footprint IoU).
 *
 * The reference has no FFI layer: its boundary is the Python module surface.  Every entry point
 * below names the reference function (file:line, relative to the reference checkout) whose arithmetic it replaces;
 * the Python modules in 3d-playground_b200/ bind them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on CUDA device `device` unless the name ends in _host;
 *   - all tensors are dense row-major ("contiguous"); sizes are int64_t; float = IEEE binary32, double = binary64;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is enqueued on it and the
 *     call returns without synchronising the device;
 *   - outputs and workspaces are caller-allocated (query the *_workspace_bytes functions);
 *   - return value: 0 on success, negative G3D_ERR_* otherwise; g3d_last_error() returns a thread-local message;
 *   - no global mutable state: safe to call concurrently from several host threads on different devices/streams
 *     (nn.DataParallel calls the loss that way, train_detector_3D_angle.py:316-318);
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef GEOM3D_H
#define GEOM3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G3D_OK 0
#define G3D_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, unsupported width ...) */
#define G3D_ERR_CUDA (-2)        /* a CUDA runtime call or kernel launch failed */
#define G3D_ERR_UNSUPPORTED (-3) /* valid request outside what the kernels implement */

/* annotation layout selector for the loss entry points */
#define G3D_VARIANT_2D 0 /* retinanet/losses.py: rows (x1,y1,x2,y2,class), 4-d regression */
#define G3D_VARIANT_3D 1 /* pytorch_retinanet_detector_directional/retinanet/losses.py: 16 corner coords + 4 box + class (+vps), 12-d regression */

/* assignment codes written to the `assign` arrays */
#define G3D_ASSIGN_IGNORE (-2)   /* 0.4 <= IoU_max < 0.5 : target row of -1 */
#define G3D_ASSIGN_NEGATIVE (-1) /* IoU_max < 0.4        : target row of 0 */
/* >= 0: positive; value = ORIGINAL annotation row (before the class != -1 filter) of the assigned GT box */

const char* g3d_version(void);
const char* g3d_last_error(void);
/* number of SMs of `device` (used by hosts to size persistent grids / report rooflines); <0 on error */
int g3d_sm_count(int device);

/* ------------------------------------------------------------------------------------------------------------------
 * a1  calc_iou(a, b)            retinanet/losses.py:5-22 == pytorch_retinanet_detector_directional/retinanet/losses.py:5-22
 * a[A,4], b[G,4] (x1,y1,x2,y2) -> out[A,G]; FP32, operation order and roundings of the reference (no FMA contraction,
 * IEEE division, union clamped at 1e-8).
 */
int g3d_calc_iou(const float* a, int64_t A, const float* b, int64_t G, float* out, int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * a6 (prologue)  GT row filter + enclosing 2D box      3D losses.py:54,93-107 ; 2D retinanet/losses.py:46,81
 * ann[B,Gmax,W]; rows whose class column (W-1 for 2D i.e. col 4; col 20 for 3D) equals -1 are dropped, the rest are
 * compacted in order.  variant 3D: box = (min,min,max,max) over the 8 corners in cols 0..15; 2D: cols 0..3.
 * gt_box[B,Gmax,4], gt_row[B,Gmax] (original row of each compacted entry), gt_count[B].
 */
int g3d_gt_prepare(const float* ann, int64_t B, int64_t Gmax, int64_t W, int variant,
                   float* gt_box, int32_t* gt_row, int32_t* gt_count, int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * a2  assignment          3D losses.py:109-131 ; 2D retinanet/losses.py:81-104
 * For every image b and anchor a: IoU_max, IoU_argmax = max_g calc_iou(anchor a, gt_box[b,g]) with torch.max's
 * first-maximal-index rule (index into the compacted GT list; 0 for an image without GT).
 * iou_max[B,A] (nullable), iou_argmax[B,A] int64 (nullable), assign[B,A] int32 codes (nullable), num_pos[B] int32
 * (nullable; zeroed by the call itself).
 */
int g3d_assign(const float* anchors, int64_t A, const float* gt_box, const int32_t* gt_row, const int32_t* gt_count,
               int64_t B, int64_t Gmax, float* iou_max, int64_t* iou_argmax, int32_t* assign, int32_t* num_pos,
               int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * a2-a6  FocalLoss.forward (+ backward) : assignment + focal classification loss + regression loss (+ direction loss)
 *        3D losses.py:27-362 ; 2D retinanet/losses.py:27-177
 * cls[B,A,C] (post-sigmoid), reg[B,A,R] (R = 12 for 3D, 4 for 2D), anchors[A,4], ann[B,Gmax,W]
 * (W >= 21 for 3D - only cols 0..20 are read -, W == 5 for 2D).
 * hyper_host (nullable, host floats[G3D_HYPER_COUNT]): {alpha, gamma, positive IoU threshold, negative IoU threshold,
 *   smooth-L1 beta, top_weighting, clamp min, clamp max}; NULL = the reference's constants {0.25, 2.0, 0.5, 0.4, 1/9, 0.5,
 *   1e-4, 1 - 1e-4} (losses.py:28-30,56,121,124,343,346-348).  gamma == 2 takes the fast kernels.
 * Outputs:
 *   losses[4]        = (classification, regression, direction "vp") batch means exactly as the reference forms them:
 *                      mean over all B images for cls/reg, mean over the images that have >=1 GT row for vp
 *                      (NaN if there is none: the reference raises there, the Python wrapper turns it into the raise);
 *                      2D variant: losses[2] = 0.  losses[3] = number of images with >= 1 GT row.
 *   per_image[B,4]   = (cls_j, reg_j, vp_j, num_pos_j) per image (vp_j = 0 and flagged by gt_count==0 for empty images)
 *   assign[B,A]      = int32 assignment codes (nullable: the kernels hand over one-byte codes in the workspace)
 *   gt_count_out[B]  = number of GT rows (class != -1) per image (nullable)
 * workspace: g3d_focal_workspace_bytes(B, A, Gmax) bytes, 256-byte aligned, contents undefined on entry.
 * Alignment: cls / dcls 32 bytes (256-bit row accesses), reg / dreg / anchors / per_image 16 bytes.
 *
 * g3d_focal_loss_fwd      : the losses only.
 * g3d_focal_loss_fwd_bwd  : the losses AND, in the same pass, the complete gradients dcls[B,A,C], dreg[B,A,R] of
 *                           grad_expected_host[0] * losses[0] + [1] * losses[1] + [2] * losses[2] (host floats: the
 *                           upstream gradients the caller expects: {1,1,1} for `(cls + reg + vp).backward()`,
 *                           train_detector_3D_angle.py:374-382).  g3d_focal_loss_bwd(have_grads = 1) verifies the
 *                           expectation on the device and recomputes what does not hold.
 *                           dcls == dreg == NULL degrades to g3d_focal_loss_fwd.
 *                           shard_stats (nullable, double[5]): sum_j cls_j, sum_j reg_j, sum over images with GT of
 *                           vp_j, B, number of images with GT - what a rank contributes to the global batch means
 *                           when the images are sharded over several GPUs (see g3d_combine_shard_stats).
 *                           trace_events (nullable): n_trace_events <= 6 cudaEvent_t handles recorded on `stream` before
 *                           the first launch and after the prologue, the assignment, the resolve pass, the streaming
 *                           pass and the positives / reduction launch (per-kernel timing without a profiler;
 *                           bench.py's roofline figures come from these).
 *                           dreg_state: G3D_DREG_UNDEFINED - dreg holds anything; all of it is written (the 48 B of
 *                           zeros per non-positive row are 43 % of the bytes the step moves).  G3D_DREG_CLEAN - the
 *                           caller keeps dreg AND the workspace from step to step and promises that dreg is what the
 *                           previous g3d_focal_loss_fwd_bwd call on this workspace left in it (or all zeros, with a zeroed
 *                           workspace, before the first step), and that no other call used the workspace in between:
 *                           only that step's positive rows (~1 % of the rows) are zeroed again, nothing is filled.
 *                           Same result.  (Honoured on the GT-centric path; the anchor-centric kernel writes every row.)
 * pyramid_host (nullable, host doubles): tells the call that `anchors` is the regular pyramid table of Anchors.forward
 *   (retinanet/anchors.py:21-40; g3d_generate_anchors): {L, S, L x (rows, cols, stride), L x S x (anchor width, height)},
 *   L <= 8 levels, S <= 16 shapes per cell, sum rows*cols*S == A.  With it (and Gmax <= 256) the assignment runs
 *   GT-centric: per GT row only the window of cells whose anchors can reach IoU 0.385 is evaluated (same arithmetic on
 *   the table's values, same codes), and the dense zero part of dreg is written by bulk async copies (TMA) spread over
 *   all launches of the step.  Without it: the anchor-centric kernel, valid for any anchor table (it writes the zeros
 *   itself, behind its instruction stream).
 */
#define G3D_HYPER_COUNT 8
#define G3D_DREG_UNDEFINED 0
#define G3D_DREG_CLEAN 1
int64_t g3d_focal_workspace_bytes(int64_t B, int64_t A, int64_t Gmax);
int g3d_focal_loss_fwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                       int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                       const float* hyper_host, float* losses, float* per_image, int32_t* assign, int32_t* gt_count_out,
                       void* workspace, int64_t workspace_bytes, const double* pyramid_host, int device, void* stream);
int g3d_focal_loss_fwd_bwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                           int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                           const float* hyper_host, const float* grad_expected_host, float* losses, float* per_image,
                           int32_t* assign, int32_t* gt_count_out, double* shard_stats, float* dcls, float* dreg,
                           void* workspace, int64_t workspace_bytes, const double* pyramid_host,
                           void* const* trace_events, int n_trace_events, int dreg_state, int device, void* stream);

/* backward of the above (autograd of the reference graph, same file:lines) for arbitrary upstream gradients.
 * grad_out[3] = upstream gradients of the three returned [1]-tensors (device memory); grad_scale[3] (nullable, device)
 * multiplies them element-wise (multi-GPU: local -> global mean, from g3d_combine_shard_stats).
 * workspace = what the forward wrote (keys / byte codes, per-image lists of positive anchors, GT tables: keep it
 * untouched between the two calls); hyper_host / pyramid_host: the values the forward was given.  On return dcls[B,A,C] and dreg[B,A,R] are complete (dreg is zero on non-positive anchors).
 * have_grads = 0: dcls / dreg are uninitialised; everything is computed here.
 * have_grads = 1: they come from g3d_focal_loss_fwd_bwd(grad_expected_host).  The kernels compare grad_out * grad_scale
 *   with grad_expected_host ON THE DEVICE: equal (the usual training step) -> both launches exit in one wave; a different
 *   classification gradient -> dcls is recomputed; a different regression / direction gradient -> the rows of the
 *   positive anchors are recomputed.  No host synchronisation is needed to pick the path.
 */
int g3d_focal_loss_bwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                       int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                       const float* hyper_host, const float* grad_out, const float* grad_scale, int have_grads,
                       const float* grad_expected_host, const void* workspace, int64_t workspace_bytes,
                       const double* pyramid_host, float* dcls, float* dreg, int device, void* stream);

/* process-wide tuning knobs of the loss path (benchmark sweeps; the defaults are the measured optimum on B200):
 * G3D_TUNE_PDL: 0 launches the positives / reduction kernel without programmatic dependent launch (default 1);
 * G3D_TUNE_FORCE_ANCHOR_CENTRIC != 0: ignore pyramid_host. */
#define G3D_TUNE_PDL 1
#define G3D_TUNE_FORCE_ANCHOR_CENTRIC 3
int g3d_set_tuning(int key, int64_t value);

/* multi-GPU reduction of the loss (replaces nn.DataParallel + .mean(), train_detector_3D_angle.py:316-318,374-378):
 * gathered[world][5] = the shard_stats of every rank (all-gathered by the host, e.g. NCCL), summed in rank order ->
 * losses[3] = global (cls, reg, vp) means as the reference forms them on one device, scale[3] = this rank's
 * d(global mean)/d(local mean) for g3d_focal_loss_bwd's grad_scale.
 */
int g3d_combine_shard_stats(const double* gathered, int64_t world, int64_t rank, float* losses, float* scale,
                            int device, void* stream);

/* The same reduction without a collective library call, over peer-mapped ("symmetric") memory of the GPUs of one
 * NVLink / NVSwitch box: every rank owns an exchange buffer of g3d_exchange_buffer_doubles(world) doubles, zero before the
 * first call, mapped into every peer; peer_ptrs_dev = DEVICE array of `world` pointers, entry r = rank r's buffer as
 * addressable from this device (torch.distributed._symmetric_memory: rendezvous(...).buffer_ptrs_dev).  One launch of one
 * warp: stores this rank's shard_stats[5] into every peer's buffer, publishes them (system-scope fence + epoch flag), waits
 * for the flags of all ranks in its own buffer (bounded: NaN losses after ~3 s instead of a hang) and forms losses[3] /
 * scale[3] exactly as g3d_combine_shard_stats does (rank-order sums).  Every rank must make the same sequence of calls.
 */
int64_t g3d_exchange_buffer_doubles(int64_t world);
int g3d_exchange_shard_stats(const double* shard_stats, const void* peer_ptrs_dev, int64_t world, int64_t rank,
                             float* losses, float* scale, int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * a7  3D BBoxTransform.forward     pytorch_retinanet_detector_directional/retinanet/utils.py:102-149
 * anchors[A,4] (the reference's boxes[1,A,4]), reg[B,A,12] -> out[B,A,20].  Bit-identical to the eager reference
 * (separate FP32 mul and add, left-to-right sums).
 */
int g3d_decode3d(const float* anchors, const float* reg, int64_t B, int64_t A, float* out, int device, void* stream);

/* a8  2D BBoxTransform.forward     retinanet/utils.py:102-126
 * anchors[Ba,A,4] with Ba == 1 or B; deltas[B,A,4]; mean[4], std[4] host arrays -> out[B,A,4].
 * If clip != 0 the ClipBoxes clamp (a9) for an image of clip_w x clip_h is fused into the same pass.
 */
int g3d_decode2d(const float* anchors, int64_t Ba, const float* deltas, int64_t B, int64_t A,
                 const float* mean_host, const float* std_host, int clip, float clip_w, float clip_h,
                 float* out, int device, void* stream);

/* a9  ClipBoxes.forward (in place)  retinanet/utils.py:134-144 ; 3D copy utils.py:157-167
 * boxes[N,K] with K >= 4: col0,col1 = max(.,0); col2 = min(., width); col3 = min(., height).
 */
int g3d_clip_boxes(float* boxes, int64_t N, int64_t K, float width, float height, int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * a10 score filter.
 * Score vectors are described by (outer, inner, N): vector s = o*inner + c is scores[o*outer_pitch + n*inner + c],
 * n in [0,N).  classification[B,A,C] taken per (image, class) is (outer=B, inner=C, N=A, outer_pitch=A*C)
 * (retinanet/model.py:287-289, 3D model.py:365-374); a flat vector is (outer=1, inner=1) (3D model.py:320-328).
 *
 * g3d_rowmax: scores, classes = torch.max(classification, dim=1) (3D model.py:320): cls[rows,C] -> smax[rows],
 *   amax[rows] int64 (first maximal index).
 * g3d_threshold_ladder: the adaptive ladder of 3D model.py:322-328 (start 1e-7, MULTI_FRAME) and :368-374 (1e-25):
 *   thresholds_host[L] = the float32-cast rungs, ascending (the caller forms them in double by repeated
 *   multiplication, as the Python loop does); one pass builds per-vector rung histograms, then
 *   rung_out[s] = first rung whose count of (score > rung) is <= keep_max (-1 if none), count_out[s] that count,
 *   thr_out[s] = the rung value (device array, feeds g3d_filter_compact without a host round trip).
 * g3d_filter_compact: for each vector append the element index n of every score > thr[s] to idx_out[s*cap ...]
 *   (arrival order - the NMS front end orders by (score, index)); count_out[s] = number of passing scores, which may
 *   exceed cap (then only cap indices were stored: the caller must retry with a larger cap).
 */
int g3d_rowmax(const float* cls, int64_t rows, int64_t C, float* smax, int64_t* amax, int device, void* stream);
int64_t g3d_ladder_workspace_bytes(int64_t S, int64_t L);
int g3d_threshold_ladder(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                         const float* thresholds_host, int64_t L, int64_t keep_max, int32_t* rung_out,
                         int32_t* count_out, float* thr_out, void* workspace, int64_t workspace_bytes, int device,
                         void* stream);
int g3d_filter_compact(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                       const float* thr, int64_t cap, int32_t* idx_out, int32_t* count_out, int device, void* stream);
/* g3d_gather_candidates: packs the candidates of g3d_filter_compact for the NMS front end.  Per vector s the stored
 *   indices (min(count[s], cap) of them, cap <= 16384) are sorted ascending - the order of the reference's boolean-mask
 *   gather (3D model.py:380-382, retinanet/model.py:294-296) - and written contiguously from seg_offsets[s]
 *   (seg_offsets[S+1] = exclusive scan, produced here): cand_scores[.], cand_src[.] (the element index n) and, when
 *   boxes != NULL, cand_boxes[.,4] = boxes[(o*N + n)*box_stride + box_col ...].
 */
int g3d_gather_candidates(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                          const float* boxes, int64_t box_stride, int64_t box_col, const int32_t* idx,
                          const int32_t* count, int64_t cap, int32_t* seg_offsets, float* cand_scores,
                          float* cand_boxes, int32_t* cand_src, int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * the detection tail without the decoded tensor.  Thresholding happens BEFORE decoding, so only the
 * candidates' NMS boxes and the kept rows are ever decoded - bit-identical to the rows g3d_decode3d / g3d_decode2d
 * (with the same clip) would produce.  Replaces, together with g3d_filter_compact / g3d_nms_segmented, everything
 * ResNet.forward does after the heads: 3D model.py:346-397 (regressBoxes, ladder, per-class nms, cat), 2D
 * retinanet/model.py:270-311 (regressBoxes, clipBoxes, scores > 0.05, nms, cat).
 *
 * g3d_gather_candidates_decoded: as g3d_gather_candidates, but cand_boxes[.,4] is decoded from reg[B,A,12] (variant 3D:
 *   columns 16..19 of the BBoxTransform row) or reg[B,A,4] (variant 2D: the BBoxTransform row, clipped if clip != 0),
 *   anchors[Ba,A,4] with Ba == 1 or B.  mean_host / std_host: 2D only (may be NULL for 3D).
 * g3d_exclusive_scan_i32: offsets[S+1] = exclusive scan of count[S] (offsets[S] = total; one 4-byte read tells the host
 *   how many detections to allocate).
 * g3d_assemble_detections: rows out_offsets[s] ... of out_scores[K], out_classes[K] (int64), out_image[K] (int64) and
 *   out_boxes[K,20|4] from the keep lists of g3d_nms_segmented(relative = 0): segment s = o*inner + c contributes its
 *   keep_count[s] kept candidates in NMS order (descending score) - image-major, then class, as the reference
 *   concatenates them.  `capacity` = rows the four output arrays hold: rows beyond it are not written (a caller may
 *   launch into buffers sized from an estimate before it has read the number of detections, and repeat if it was short).
 */
int g3d_gather_candidates_decoded(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                                  const float* anchors, int64_t Ba, const float* reg, int variant,
                                  const float* mean_host, const float* std_host, int clip, float clip_w, float clip_h,
                                  const int32_t* idx, const int32_t* count, int64_t cap, int32_t* seg_offsets,
                                  float* cand_scores, float* cand_boxes, int32_t* cand_src, int device, void* stream);
int g3d_exclusive_scan_i32(const int32_t* count, int64_t S, int32_t* offsets, int device, void* stream);
/* g3d_detect_tail: g3d_filter_compact -> g3d_gather_candidates_decoded -> g3d_nms_segmented(relative = 0) ->
 *   g3d_exclusive_scan_i32 issued back to back from one call (the chain is launch-latency bound; see detect_tail.cu),
 *   S = outer*inner segments, thr[S] on the device.  Outputs as the individual entry points define them: count[S],
 *   seg_offsets[S+1], cand_scores[S*cap], cand_src[S*cap], keep[S*cap], keep_count[S], out_offsets[S+1]; summary[4] =
 *   {number of detections = out_offsets[S], max count[s], 0, 0} for the ONE device->host read the caller needs before
 *   g3d_assemble_detections.  `summary` may also be mapped pinned host memory: entry 3 is written last, after a system-wide
 *   fence, so a host that set it to a negative value beforehand can poll it and then read entries 0..2 (no copy, no event).  workspace: g3d_detect_tail_workspace_bytes(S, cap) bytes, 256-byte aligned. */
int64_t g3d_detect_tail_workspace_bytes(int64_t S, int64_t cap);
int g3d_detect_tail(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch, const float* thr,
                    int64_t cap, const float* anchors, int64_t Ba, const float* reg, int variant,
                    const float* mean_host, const float* std_host, int clip, float clip_w, float clip_h,
                    double iou_threshold, int32_t* count, int32_t* seg_offsets, float* cand_scores, int32_t* cand_src,
                    int64_t* keep, int32_t* keep_count, int32_t* out_offsets, int32_t* summary, void* workspace,
                    int64_t workspace_bytes, int device, void* stream);
/* g3d_detect_tail_short: the same call for the usual case of SHORT segments (at most 1024 candidates per (image, class),
 *   iou_threshold >= 0): after the score filter ONE launch does gather + decode + sort + NMS per segment in shared memory
 *   (pair search through a grid over the box centres, exact torchvision test on the pairs found; detect_tail.cu), then
 *   one launch writes out_offsets and summary.  Same arguments, same results, with two differences in layout:
 *   the candidate tables of segment s start at seg_offsets[s] = s*cap and are in NMS order (score descending, anchor
 *   ascending) - keep / g3d_assemble_detections work on them unchanged - and summary has 4 entries:
 *   {detections, max count[s], segments NOT processed, 0}.  A segment that is too long, has more than a few thousand suppressing
 *   pairs or holds boxes with sides < 1e-10 or coordinates > 1e15 gets keep_count[s] = -1 and is counted in summary[2]:
 *   the caller then repeats the tail with g3d_detect_tail.  Workspace as g3d_detect_tail. */
int g3d_detect_tail_short(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                          const float* thr, int64_t cap, const float* anchors, int64_t Ba, const float* reg, int variant,
                          const float* mean_host, const float* std_host, int clip, float clip_w, float clip_h,
                          double iou_threshold, int32_t* count, int32_t* seg_offsets, float* cand_scores,
                          int32_t* cand_src, int64_t* keep, int32_t* keep_count, int32_t* out_offsets, int32_t* summary,
                          void* workspace, int64_t workspace_bytes, int device, void* stream);
int g3d_assemble_detections(const int64_t* keep, const int32_t* keep_count, const int32_t* seg_offsets,
                            const int32_t* out_offsets, const float* cand_scores, const int32_t* cand_src,
                            int64_t outer, int64_t inner, int64_t N, const float* anchors, int64_t Ba,
                            const float* reg, int variant, const float* mean_host, const float* std_host,
                            int clip, float clip_w, float clip_h, float* out_scores, int64_t* out_classes,
                            float* out_boxes, int64_t* out_image, int64_t capacity, int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * a11 NMS       torchvision.ops.nms as called at retinanet/model.py:297, 3D model.py:383 (and :336 through
 *               batched_nms :19-57), perform_3D_detection_on_video_sequences.py:78, MC3D_crop_tracker.py:507,614,634
 * Segmented greedy NMS: S independent segments, segment s = entries seg_offsets[s] .. seg_offsets[s+1]-1 of
 * boxes[N, box_stride floats] (the 4 box columns start at column box_col) and scores[N].
 * Per segment: stable descending score order (ties: lower index first); box j is suppressed by a kept box i iff
 * IoU(i,j) > iou_threshold, IoU in FP32 as torchvision computes it, the comparison against the double threshold.
 * keep_out[N] int64: for segment s the kept indices (relative to the segment start when relative != 0, else global)
 * in descending score order, stored from position seg_offsets[s]; keep_count[S] int32.
 * seg_offsets is a DEVICE int32 array of S+1 entries (NULL is allowed for S == 1: one segment [0, N)); max_seg_len is a
 * host upper bound of the longest segment.
 * Three execution shapes, one result: many short segments (S >= 64) run one 256-thread CTA per segment; a single
 * segment with capacity max_seg_len == N of 256..16384 boxes (the trackers' calls) is rank-sorted (N <= 4096), builds a
 * 64x64-block suppression bit-matrix with the whole GPU and scans it with one warp fed from shared memory; everything
 * else runs the one-CTA-per-segment greedy pass.  The segment bounds are always read on the device, so a captured
 * launch sequence (CUDA graph) with S == 1 and seg_offsets = {lo, hi} serves any hi - lo <= N.
 */
int64_t g3d_nms_workspace_bytes(int64_t N, int64_t S, int64_t max_seg_len);
int g3d_nms_segmented(const float* boxes, int64_t box_stride, int64_t box_col, const float* scores, int64_t N,
                      const int32_t* seg_offsets, int64_t S, int64_t max_seg_len, double iou_threshold, int relative,
                      int64_t* keep_out, int32_t* keep_count, void* workspace, int64_t workspace_bytes,
                      int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * a12 Homography.i24_state_to_space        homography.py:305-320
 * states[d,S] (S >= 6; x,y,l,w,h,direction in cols 0..5) -> out[d,8,3] float32.
 */
int g3d_state_to_space(const float* states, int64_t d, int64_t S, float* out, int device, void* stream);

/* a13 Homography.space_to_im               homography.py:438-476 (wrapper select :849-856)
 * pts[d,m,3] float32 or float64 (pts_is_f64), P[ncam,2,3,4] float64 = per camera the 3x4 projection of the first
 * (hg1 / EB) and second (hg2 / WB) homography; cam[d] uint8 camera index per object (nullable -> camera `cam_const`);
 * wrapper != 0 selects the second matrix for objects whose pts[i,0,1] > 60 (Homography_Wrapper), else always the first.
 * out[d,m,2] float64.
 */
int g3d_space_to_im(const void* pts, int pts_is_f64, int64_t d, int64_t m, const double* P, int64_t ncam,
                    const uint8_t* cam, int cam_const, int wrapper, double* out, int device, void* stream);

/* a14 Homography.state_to_im               homography.py:479-488 (wrapper :861-862)
 * fused a12 o a13 without the [d,8,3] intermediate.  out[d,8,2] float64 (out_f32 == 0) or float32 (opt-in).
 * all_cams != 0: project every state into all ncam cameras -> out[d,ncam,8,2] (cam ignored).
 */
int g3d_state_to_im(const float* states, int64_t d, int64_t S, const double* P, int64_t ncam,
                    const uint8_t* cam, int cam_const, int wrapper, int all_cams, void* out, int out_f32,
                    int device, void* stream);

/* a15 Homography.im_to_space               homography.py:388-435 (wrapper :840-847)
 * pts[d,8,2] float32/float64, heights[d] float32/float64 (same flag), H[ncam,2,3,3] float64 image->road-plane
 * homographies; out[d,8,3] float64: (x,y) = H.(u,v,1) dehomogenised, z = 0 for corners 0..3, +height for 4..7.
 * wrapper != 0: objects whose first-matrix result has out[i,0,1] > 60 are recomputed with the second matrix.
 */
int g3d_im_to_space(const void* pts, const void* heights, int in_is_f64, int64_t d, const double* H, int64_t ncam,
                    const uint8_t* cam, int cam_const, int wrapper, double* out, int device, void* stream);

/* a16 Homography.i24_space_to_state        homography.py:274-303
 * pts[d,8,3] float32/float64 -> out[d,6] float32.
 */
int g3d_space_to_state(const void* pts, int pts_is_f64, int64_t d, float* out, int device, void* stream);

/* a17 Homography.im_to_state               homography.py:491-500 (wrapper :858-859): fused a16 o a15. */
int g3d_im_to_state(const void* pts, const void* heights, int in_is_f64, int64_t d, const double* H, int64_t ncam,
                    const uint8_t* cam, int cam_const, int wrapper, float* out, int device, void* stream);

/* a18 Homography.height_from_template      homography.py:519-551
 * template_boxes[d,8,2], template_heights[d], boxes[d,8,2], each float32 or float64 (its *_is_f64 flag); every
 * operand is evaluated in the dtype torch's type promotion would use; out[d] is float64 if any input is, else float32.
 */
int g3d_height_from_template(const void* template_boxes, int tb_is_f64, const void* template_heights, int th_is_f64,
                             const void* boxes, int bx_is_f64, int64_t d, void* out, int device, void* stream);

/* a17+a14+a18+a17: the trackers' two-pass height refinement idiom  MC3D_crop_tracker.py:364-370, :1222-1227;
 * mot_evaluator.py:169-176; fit_filter_3D.py:262-266:
 *   s0 = im_to_state(pts, h0); repro = state_to_im(s0); h1 = height_from_template(repro, h0, pts);
 *   out = im_to_state(pts, h1).  P and H as above.  heights_out[d] (nullable) receives h1 (double).
 */
int g3d_im_to_state_refined(const void* pts, const void* heights, int in_is_f64, int64_t d, const double* H,
                            const double* P, int64_t ncam, const uint8_t* cam, int cam_const, int wrapper,
                            float* out, double* heights_out, int device, void* stream);

/* a19 footprint box from state             MC3D_crop_tracker.py:625-632 (also :271-275,:498-502,:668-682)
 * states[d,S] -> out[d,4] float32 (xmin,ymin,xmax,ymax of the 4 bottom corners of state_to_space).
 */
int g3d_state_footprint(const float* states, int64_t d, int64_t S, float* out, int device, void* stream);

/* a19' image box from 8 corners            MC3D_crop_tracker.py:602-607 ; minimal_3D_track.py:505-510
 * pts[d,8,2] float32/float64 -> out[d,4] same dtype (min x, min y, max x, max y).
 */
int g3d_corners_to_box(const void* pts, int is_f64, int64_t d, void* out, int device, void* stream);

/* a20 md_iou, un-broadcast form            MC3D_crop_tracker.py:1030-1049 (callers :280,:689,:1013)
 * first[n,4], second[m,4] float32 (promoted to float64 exactly as the callers' .double()) -> out[n,m] float64
 * = IoU(first[i], second[j]) with NO epsilon (0/0 -> NaN) when eps == 0; eps is added to the union for the
 * evaluator variant (mot_evaluator.py:115).  one_minus != 0 writes 1 - IoU (match_hungarian's cost, :689).
 */
int g3d_pairwise_iou_f64(const float* first, int64_t n, const float* second, int64_t m, double eps, int one_minus,
                         double* out, int device, void* stream);
/* a20 literal form: a[n,4], b[n,4] float64 element-wise -> out[n] (the reference signature with pre-broadcast inputs) */
int g3d_md_iou(const double* a, const double* b, int64_t n, double* out, int device, void* stream);
/*
 * SURVEY §8(f)-3  estimate_ts_bias pair mining              MC3D_crop_tracker.py:277-289
 * All pairs (i, j), i < j, cams[i] != cams[j], md_iou(boxes[j], boxes[i]) > threshold (float64 IoU of the float32
 * footprints boxes[d,4], no epsilon), in the reference's loop order (i outer, j inner).  Two passes around
 * g3d_exclusive_scan_i32: row_offsets == NULL writes row_count[d]; row_offsets[d+1] != NULL writes pairs[.,2] (int32 i, j)
 * at the rows' offsets (entries at positions >= capacity are dropped).  The d x d matrix is never materialised.
 */
int g3d_cross_camera_pairs(const float* boxes, const int32_t* cams, int64_t d, double threshold, int32_t* row_count,
                           const int32_t* row_offsets, int32_t* pairs, int64_t capacity, int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * SURVEY §8(f)-4  Torch_KF.predict / Torch_KF.update        util_track/kf.py:292-336, :339-403
 * The batched Kalman filter of the trackers: X[n,S] float32 states, P[n,S,S] float32 covariances (S <= 8; the trackers
 * use S = 6, M = 5).  One thread per object, everything in registers, same FP32 / FP64 split as the reference.
 *
 * g3d_kf_predict: F_rep = F with F_rep[0,5] = D[i] * dt (S > 5);  X = F_rep X;  P = F_rep P F_rep^T + Q * dt / dt_default.
 *   dt: one host scalar (dt_per_object == NULL; FP32 scaling of Q as torch does for a python float) or a device
 *   float64 array [n] (the `dts = filter.get_dt(...)` form; the scaling and the sum in FP64, then rounded to FP32).
 *   T[n] (nullable, float64) += dt.   F_host[S*S], Q_host[S*S]: host arrays.
 * g3d_kf_update: for the objects rows[j] (int64 indices into X / P, distinct) with measurements z[m,M] float64:
 *   y = z + mu_R - H x;  S = H P H^T + R;  K = P H^T S^-1;  x += K y;  P = (I - K H) P.
 *   H_host[M*S], R_host[M*M], mu_R_host[M] (nullable = 0): host arrays.
 */
int g3d_kf_predict(float* X, float* P, const float* D, const double* dt_per_object, double dt_scalar, double dt_default,
                   double* T, int64_t n, int64_t S, const float* F_host, const float* Q_host, int device, void* stream);
int g3d_kf_update(float* X, float* P, const int64_t* rows, const double* z, int64_t m_count, int64_t S, int64_t M,
                  const float* H_host, const float* R_host, const float* mu_R_host, int device, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * SURVEY §8(f)-2  Anchors.forward                           retinanet/anchors.py:21-40 (generate_anchors :42-74, shift :109-129)
 * Writes the float32 [A,4] anchor table on the device instead of building it in numpy and copying it on every forward.
 * Order: level -> cell row-major (y outer, x inner) -> the S shapes of the level.  anchors[i] =
 * float32(shapes[level][s] + ((col|row) + 0.5) * stride[level]) with the sum in FP64 (numpy's float64 then astype f32).
 * shapes_host[L][S][4] (x1,y1,x2,y2 around the origin, from generate_anchors), strides_host[L], rows_host[L], cols_host[L]
 * (feature-map sizes ceil(H / 2^level), ceil(W / 2^level)): host arrays.  L <= 8, L*S <= 96.  A = sum rows*cols*S.
 */
int g3d_generate_anchors(const double* shapes_host, const double* strides_host, const int64_t* rows_host,
                         const int64_t* cols_host, int64_t L, int64_t S, float* anchors, int64_t A, int device,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOM3D_H */
