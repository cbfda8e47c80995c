/*
 * vyolo.h -- C ABI of the B200-native VideoYOLO detection post-processing path.
 *
 * The reference (HaydenFaulkner/VideoYOLO) has no FFI layer of its own: the path sits behind
 * Gluon HybridBlocks and one MXNet contrib-op call.  Each entry point below names the reference
 * interface it replaces (file:line under /root/reference).  INTEGRATION.md shows the ctypes stub
 * a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory unless the name says host_;
 *     the small shape/anchor arrays (H, W, stride, anchors, head[]) are HOST arrays;
 *   - no allocation, no ownership transfer, no implicit synchronisation: work is enqueued on
 *     `stream` (a cudaStream_t passed as void*) and the caller owns ordering;
 *   - the caller has made the target GPU current (cudaSetDevice) before calling;
 *   - returns VY_OK (0) or a negative VY_E* code; vy_last_error() gives a thread-local message;
 *   - sm_100a only.  There is no CPU fallback anywhere in this library.
 */
#ifndef VYOLO_H_
#define VYOLO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VY_ABI_VERSION 1

#define VY_OK            0
#define VY_EINVAL       -1   /* bad shape / argument                       */
#define VY_EALIGN       -2   /* pointer not aligned as documented          */
#define VY_EWORKSPACE   -3   /* workspace missing or too small             */
#define VY_ECUDA        -4   /* CUDA runtime/driver error, see last_error  */
#define VY_EUNSUPPORTED -5   /* valid request this build cannot serve      */

#define VY_MAX_SCALES   4
#define VY_MAX_ANCHORS  8    /* anchors per scale */

#define VY_FMT_CORNER   0    /* MXNet box_nms in_format/out_format 'corner' */
#define VY_FMT_CENTER   1    /* 'center'                                    */

typedef void *vy_stream_t;   /* cudaStream_t */

int         vy_version(void);
const char *vy_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Launch accounting (measurement only; no reference counterpart).  Every kernel the library
 * launches is counted under one of the ids below.  While vy_prof_enable(1) is in force each launch
 * is additionally bracketed by cudaEventRecord on the stream it is launched on; vy_prof_read()
 * synchronises those events, returns summed milliseconds and launch counts per id, and resets.
 */
#define VY_K_DECODE        0
#define VY_K_SELECT_HEADS  1
#define VY_K_SELECT_ROWS   2
#define VY_K_FINALIZE      3
#define VY_K_FILL          4
#define VY_K_IOU           5
#define VY_K_FUSION_CONV   6
#define VY_K_TEMPORAL_POOL 7
#define VY_K_NMS_LARGE     8
#define VY_K_LAYOUT        9
#define VY_K_SAMPLE       10
#define VY_K_STREAM       11
#define VY_K_TABLE        12      /* (development builds with -DVY_STREAM_ALT only) */
#define VY_K_COUNT        13
const char *vy_kernel_name(int kernel_id);
int vy_launch_counts(long long *host_counts, int n);     /* cumulative since load; returns VY_K_COUNT */
int vy_prof_enable(int on);
int vy_prof_read(double *host_ms, long long *host_launches, int n);

/* ---------------------------------------------------------------------------------------------
 * Anchor decode of the YOLOv3 output layers into the reference's detection tensor.
 * Replaces: YOLOOutputV3.hybrid_forward, models/definitions/yolo/yolo3.py:151-199 (dup
 *           yolo3_temporal.py:137-179) for each scale, plus the scale concat yolo3.py:523.
 *   host_head[s]  device ptr, (B, A*(5+C), H[s], W[s]) fp32 NCHW = output of the 1x1 `prediction`
 *                 conv (yolo3.py:157); scales in network order (stride 32, 16, 8: yolo3.py:416-417)
 *   host_anchors  n_scales*A*2 floats (w,h) in the same scale order
 *   dets          (B, R, 6) rows [cls, score, x1, y1, x2, y2];  R = Ceff * sum_s H*W*A,
 *                 Ceff = agnostic ? 1 : C; row = Ceff*sum_{s'<s} n_s' + c*n_s + (y*W+x)*A + a
 *   agnostic      yolo3.py:184-188 branch: one row per box, id 0, score = objectness
 */
int vy_decode_f32(const float *const *host_head, const int *host_H, const int *host_W,
                  const float *host_stride, const float *host_anchors, int n_scales,
                  int B, int A, int C, int agnostic, float *dets, vy_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * box_nms on a materialised detection tensor.
 * Replaces: F.contrib.box_nms(...) at yolo3.py:526-528 (+ :807-809, :1198-1200, :1578-1580,
 *           :1862-1864, yolo3_temporal.py:545-547) = MXNet `_contrib_box_nms`, and optionally the
 *           slice_axis(0, post_nms) that follows it (yolo3.py:529-530).
 *   data       (B, R, W_elem) fp32; B = product of all leading dims
 *   out_rows   rows written per image: R for the operator's own output shape, or post_nms to
 *              fuse the slice.  out is (B, out_rows, W_elem), kept_rows (B, out_rows) int32 or NULL.
 *              Rows after the survivors are filled with -1 (kept_rows too).
 *   topk<0 means all R rows take part; background_id<0 disables the id filter (MXNet >= 1.5).
 */
size_t vy_box_nms_workspace_bytes(int B, long R, int W_elem, int topk);
int vy_box_nms_f32(const float *data, int B, long R, int W_elem,
                   float overlap_thresh, float valid_thresh, int topk,
                   int coord_start, int score_index, int id_index, int background_id,
                   int force_suppress, int in_format, int out_format,
                   long out_rows, float *out, int32_t *kept_rows,
                   void *workspace, size_t workspace_bytes, vy_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused decode + box_nms + post_nms slice: head maps -> (ids, scores, bboxes) without ever
 * materialising the (B, R, 6) tensor.
 * Replaces: the inference tail of YOLOV3*.hybrid_forward, yolo3.py:496,523-534 (and the same
 *           tail in YOLOV3T :1159,1195-1206, YOLOV3TS, YOLOV3TB, YOLOV3_noback :1859-1870,
 *           YOLOV3Temporal yolo3_temporal.py:542-550).
 *   out        (B, post_nms, 6) rows [cls, score, x1, y1, x2, y2], -1 padded; the caller's
 *              ids/scores/bboxes are column slices of it (yolo3.py:531-533)
 *   kept_rows  (B, post_nms) int32 reference row index of every output row (-1 padded), or NULL
 *   requires 1 <= min(topk, R) <= 1024 and post_nms >= 1 (else use vy_decode_f32 + vy_box_nms_f32)
 */
size_t vy_decode_nms_workspace_bytes(const int *host_H, const int *host_W, int n_scales,
                                     int B, int A, int C, int agnostic, int topk);
int vy_decode_nms_f32(const float *const *host_head, const int *host_H, const int *host_W,
                      const float *host_stride, const float *host_anchors, int n_scales,
                      int B, int A, int C, int agnostic,
                      float overlap_thresh, float valid_thresh, int topk, int force_suppress,
                      int post_nms, float *out, int32_t *kept_rows,
                      void *workspace, size_t workspace_bytes, vy_stream_t stream);

/* The same call with everything that depends only on the shapes and arguments resolved ONCE (head description, job /
 * tile / table plan, workspace layout): a hybridized Gluon block runs the same graph on the same shapes for every
 * batch (YOLOV3.hybridize(), detect_yolo3.py:204), so the per-call host work should be a handful of launches.
 *   create   host arrays as for vy_decode_nms_f32 (no head pointers); *plan is owned by the caller until destroy
 *   launch   host_head[s] device pointers of this call's head maps; everything else as vy_decode_nms_f32.
 *            A plan may be launched from several host threads / on several streams at once as long as every
 *            concurrent launch has its own workspace and outputs.
 * vy_decode_nms_f32 is create + launch on a stack plan (same kernels, same results). */
typedef struct vy_decode_nms_plan vy_decode_nms_plan_t;
int    vy_decode_nms_plan_create(const int *host_H, const int *host_W, const float *host_stride,
                                 const float *host_anchors, int n_scales, int B, int A, int C, int agnostic,
                                 float overlap_thresh, float valid_thresh, int topk, int force_suppress,
                                 int post_nms, vy_decode_nms_plan_t **plan);
size_t vy_decode_nms_plan_workspace_bytes(const vy_decode_nms_plan_t *plan);
int    vy_decode_nms_plan_launch(const vy_decode_nms_plan_t *plan, const float *const *host_head, float *out,
                                 int32_t *kept_rows, void *workspace, size_t workspace_bytes, vy_stream_t stream);
void   vy_decode_nms_plan_destroy(vy_decode_nms_plan_t *plan);

/* ---------------------------------------------------------------------------------------------
 * Pairwise IoU.  Replaces: utils/bbox.py:11-38 bbox_iou(bbox_a, bbox_b, offset).
 *   a (N, lda>=4), b (M, ldb>=4) row-major [x1,y1,x2,y2,...]; out (N, M).
 */
int vy_bbox_iou_f32(const float *a, int N, int lda, const float *b, int M, int ldb,
                    float offset, float *out, vy_stream_t stream);
int vy_bbox_iou_f64(const double *a, int N, int lda, const double *b, int M, int ldb,
                    double offset, double *out, vy_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Anchor matching of the prefetch target generator ("next" row f4).
 * Replaces: models/definitions/yolo/yolo_target.py:86-94 -- shift_gt_boxes / shift_anchor_boxes, nd.contrib.box_iou
 *           (MXNet, corner format) and `ious.argmax(axis=1)`: for every ground-truth box the anchor whose zero-centred
 *           box overlaps its zero-centred box best.
 *   gt_boxes (B, M, 4) corner boxes (padding rows of -1 get match 0, as the reference computes before it skips them),
 *   16-byte aligned; anchors (A, 2) = all_anchors (:63) as (w, h), A <= 32; matches (B, M) int32;
 *   ious (B, A, M) or NULL = `ious` after the transpose at :92.  Accounted under VY_K_IOU. */
int vy_anchor_match_f32(const float *gt_boxes, int B, int M, const float *anchors, int A, int32_t *matches,
                        float *ious, vy_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Batched pairwise IoU of the dynamic-target step ("next" row f4).
 * Replaces: gluoncv.nn.bbox.BBoxBatchIOU as called at models/definitions/yolo/yolo_target.py:171,202
 *           (defaults: corner format, offset 0, eps 1e-15), fused with `ious.max(axis=-1)` (:203) and the
 *           ignore mask `(ious_max > ignore_iou_thresh) * -1` (:204).
 *   a (B, N, 4), b (B, M, 4) corner boxes, 16-byte aligned; any of the outputs may be NULL (not all):
 *   ious (B, N, M), ious_max (B, N), objness (B, N) = -1 where ious_max > ignore_thresh else 0.
 *   Accounted under VY_K_IOU. */
int vy_bbox_batch_iou_f32(const float *a, const float *b, int B, int N, int M, float offset, float eps,
                          float ignore_thresh, float *ious, float *ious_max, float *objness, vy_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The consumer step that follows net(x) in the reference, on the device.
 * Replaces: `bboxes.clip(0, W)` (detect_yolo3.py:226, train_yolov3.py:477), `valid_pred = id >= 0` and
 *           `box / W` (detect_yolo3.py:254-258) -- the part of detect()/validate() between the forward and
 *           the python lists / metric update.
 *   dets     (B, P, 6) rows [id, score, x1, y1, x2, y2] as written by vy_decode_nms_f32 / vy_box_nms_f32
 *   clipped  (B, P, 4) every row's box clipped to [0, clip_hi] (padding rows clip to 0, like NDArray.clip)
 *   normed   (B, P, 4) clipped / norm for rows with id >= 0, -1 elsewhere
 *   counts   (B) int32 number of rows with id >= 0 (they are the first rows of each image)
 *   Accounted under VY_K_IOU. */
int vy_detect_consume_f32(const float *dets, int B, int P, float clip_hi, float norm, float *clipped,
                          float *normed, int32_t *counts, vy_stream_t stream);

/* hierarchical_nms on the device ("next" row f3).
 * Replaces: hierarchical_nms(predictions, dataset, ov_thresh, conf_thresh, level_thresh) with its iou() helper,
 *           detect_yolo3.py:712-789 -- per image, the python double loop over the detections.
 *   boxes   (B, N, 6) rows [cls, conf, x1, y1, x2, y2] in the order detect() collected them; rows with cls < 0 = padding
 *   lifted  (n_cls) int32: the class the `while levels[cls] > level_thresh` walk (:766-767) ends at, per class
 *   branch  (n_cls, n_cls) uint8: dataset.on_branch(i, j) (:743-747)
 *   out     (B, N, 6) the new prediction rows in the order the reference appends them, -1 padded; counts (B) rows kept
 *   float32 arithmetic in the reference's operation order (detect() hands iou() np.float32 values).  N <= 1024. */
int vy_hier_nms_f32(const float *boxes, int B, int N, const int32_t *lifted, const unsigned char *branch, int n_cls,
                    float ov_thresh, float conf_thresh, float *out, int32_t *counts, vy_stream_t stream);

/* The per-image part of the VOC metric update on the device ("next" row f3).
 * Replaces: VOCMApMetric.update, metrics/pascalvoc.py:116-184 (strip padding, per class: sort by score, bbox_iou against
 *           the class's ground truths, greedy true-positive assignment), i.e. what validate() calls per batch
 *           (train_yolov3.py:473-488).
 *   dets      (B, P, 6) rows [id, score, x1, y1, x2, y2], -1 padded (vy_decode_nms_f32 / vy_box_nms_f32 output)
 *   gt_boxes  (B, M, 4); gt_labels (B, M) (< 0 = padding); gt_difficult (B, M) or NULL
 *   out_label / out_score / out_match (B, P): the valid predictions ordered by (class ascending, score descending, later
 *             row first among equal scores), match in {1 true positive, 0 false positive, -1 difficult}; -2 padded
 *   counts (B) valid predictions; n_pos (B, n_class) non-difficult ground truths per class
 *   The host then only extends its per-class lists.  P <= 1024, M <= 512. */
int vy_voc_match_f32(const float *dets, const float *gt_boxes, const float *gt_labels, const float *gt_difficult,
                     int B, int P, int M, int n_class, float iou_thresh, int32_t *out_label, float *out_score,
                     int32_t *out_match, int32_t *counts, int32_t *n_pos, vy_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Temporal fusion convolution: LeakyReLU(BN(ConvND(x))), use_bias=False, stride 1, groups 1.
 * Replaces: Conv / _conv2d / _conv3d / _conv21d cells, models/definitions/layers.py:63-89,135-158,
 *           as used by YOLODetectionBlockV3 (yolo3.py:229-253) -- one call per conv+BN+LReLU cell
 *           ('21' = two calls: (1,3,3) then (3,1,1), layers.py:82-89).  tcgen05/TMEM implicit GEMM
 *           fed by TMA.
 *
 * Activations use the library's "P layout": [T][B][H+2][W+2][C] bf16 (fp32 allowed for the last
 * output), time outermost, channels innermost, with a one-pixel ZERO spatial border.  The conv
 * writes a valid P-layout tensor (border included), so cells chain without repacking.  2-D convs
 * are T = 1; the 'cat' join (yolo3.py:1135-1136) is T = 1 with C = K*C channels.
 *   x        P layout, Cin channels
 *   w        (Cout, kt, kh, kw, Cin) bf16  (the reference's (Cout, Cin, kt, kh, kw) permuted)
 *   scale, shift  per-Cout fp32 folded inference BatchNorm: y = conv*scale + shift
 *                 (scale = gamma/sqrt(var+eps), shift = beta - mean*scale; layers.py:68,77)
 *   y        P layout, Cout channels, bf16 (y_is_f32 = 0) or fp32 (1); 'same' padding p = k/2
 *   requires Cin % 64 == 0, Cout % 64 == 0, kt,kh,kw in {1,3}; needs no workspace.
 * The tile shape is chosen per call: 128 x {64,128,256} tiles on one CTA, or 256 x {128,256} tiles on a CTA pair
 * (tcgen05 cta_group::2, cluster of 2) where that does not cost whole extra rounds of tiles.
 */
size_t vy_p_layout_elems(int B, int T, int H, int W, int C);
size_t vy_fusion_conv_workspace_bytes(int B, int T, int H, int W, int Cin, int Cout,
                                      int kt, int kh, int kw);
int vy_fusion_conv_bf16(const void *x, const void *w, const float *scale, const float *shift,
                        float leaky_slope, int B, int T, int H, int W, int Cin, int Cout,
                        int kt, int kh, int kw, void *y, int y_is_f32,
                        void *workspace, size_t workspace_bytes, vy_stream_t stream);

/* The 1x1 prediction conv over the 'cat'-joined window (reshape (0,-3,-2), yolo3.py:1135-1136) without materialising
 * the join: x is a P-layout tensor of T frames x C channels, the GEMM's K runs over rep*T*C channels -- k-block cb reads
 * channels (cb % (C/64))*64 of frame (cb / (C/64)) % T; rep > 1 walks the window again (weights split into bf16 hi | lo
 * parts: w is (Cout, rep*T*C)).  T = 1, rep = 2 serves the 'max' / 'mean' joins.  Output as vy_fusion_conv_bf16_nchw. */
int vy_fusion_conv_bf16_nchw_joined(const void *x, const void *w, const float *scale, const float *shift,
                                    float leaky_slope, int B, int T, int H, int W, int C, int rep, int Cout,
                                    float *y, int out_channels, vy_stream_t stream);

/* The same cell with the late 'max' join over the window (TemporalPooling(k, 'max'), layers.py:201-205, as used at
 * yolo3.py:1134-1138) done in its epilogue: y is ONE P-layout frame (T = 1, Cout channels, bf16) = max over the T
 * output frames; the un-pooled tip is never written.  Bit-identical to vy_fusion_conv_bf16 + vy_temporal_pool_bf16. */
int vy_fusion_conv_bf16_maxpool(const void *x, const void *w, const float *scale, const float *shift,
                                float leaky_slope, int B, int T, int H, int W, int Cin, int Cout,
                                int kt, int kh, int kw, void *y, vy_stream_t stream);

/* The same cell for T = 1 with the result written in the reference's own layout: y is the fp32 (B, out_channels, H, W)
 * tensor (NCHW), the first out_channels <= Cout channels, interior pixels only -- the 1x1 `prediction` conv of
 * YOLOOutputV3 (yolo3.py:62,157; Cout = all_pred padded to a multiple of 64, leaky_slope = 1, bias as shift) hands its
 * head maps to vy_decode_nms_f32 without a layout pass in between. */
int vy_fusion_conv_bf16_nchw(const void *x, const void *w, const float *scale, const float *shift,
                             float leaky_slope, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                             float *y, int out_channels, vy_stream_t stream);

/* Layout conversion between the reference's fp32 tensors and the P layout.  Element (b, c, t, h, w)
 * of the fp32 tensor lives at x[b*stride_b + c*stride_c + t*stride_t + h*W + w], which covers both
 * NCDHW (after the swapaxes of yolo3.py:256-262) and (B, K, C, H, W) (before it), and T = 1 NCHW. */
int vy_pack_f32_to_p_bf16(const float *x, long long stride_b, long long stride_c, long long stride_t,
                          int B, int C, int T, int H, int W, void *y_p, vy_stream_t stream);
int vy_unpack_p_to_f32(const void *y_p, int p_is_f32, int B, int C, int T, int H, int W, float *x,
                       long long stride_b, long long stride_c, long long stride_t, vy_stream_t stream);
/* the same for the first C of the Cp channels of a P-layout tensor (a conv whose Cout was padded to a multiple of 64,
 * e.g. the 1x1 `prediction` conv of YOLOOutputV3, yolo3.py:62, with A*(5+classes) = 75 / 105 / 255 outputs) */
int vy_unpack_p_channels_to_f32(const void *y_p, int p_is_f32, int B, int Cp, int C, int T, int H, int W, float *x,
                                long long stride_b, long long stride_c, long long stride_t, vy_stream_t stream);

/* The 1x1 `prediction` conv of YOLOOutputV3 (yolo3.py:62,157: Conv2D with bias, fp32 in the reference) runs in the
 * fusion-conv kernel with fp32-grade operands: a value v is carried as hi = bf16(v), lo = bf16(v - hi) in separate
 * channels and w*v is accumulated in fp32 as w_hi*v_hi + w_lo*v_hi + w_hi*v_lo (dropped term ~2^-16 relative).
 *   vy_pack_f32_split_to_p_bf16   fp32 tensor (strides as vy_pack_f32_to_p_bf16) -> P layout with 3*Cpad channels
 *                                 [hi | hi | lo], channels C..Cpad of each third zero; pairs with weights [w_hi | w_lo | w_hi]
 *   vy_cat_repeat_bf16            P layout (T, B, H+2, W+2, C) bf16 -> (1, B, H+2, W+2, rep*T*C): channel r*T*C + t*C + c
 *                                 = frame t, channel c.  T = K, rep = 1 is the 'cat' late join (reshape (0,-3,-2),
 *                                 yolo3.py:1135-1136); rep = 2 feeds a bf16 activation to split weights [w_hi | w_lo].
 *   Both are accounted under VY_K_LAYOUT. */
int vy_pack_f32_split_to_p_bf16(const float *x, long long stride_b, long long stride_c, long long stride_t,
                                int B, int C, int Cpad, int T, int H, int W, void *y_p, vy_stream_t stream);
int vy_cat_repeat_bf16(const void *x, int B, int T, int H, int W, int C, int rep, void *y, vy_stream_t stream);

/* TemporalPooling 'direct' style (layers.py:201-205) on P-layout data: x (T, inner) -> y (inner),
 * inner = B*(H+2)*(W+2)*C, mode 0 = max, 1 = mean.  bf16 in/out. */
int vy_temporal_pool_bf16(const void *x, int T, long inner, int mode, void *y, vy_stream_t stream);

/* The join between two scales of the YOLO neck on P-layout data ("next" row f2).
 * Replaces: `_upsample(x, stride=2)` (layers.py:11-20: pixel repeat), `F.slice_like(upsample, route_now, axes=(3,4))` and
 *           `F.concat(..., route_now, dim=2)` (yolo3.py:1170-1177; 4-D variant :1177).
 *   up    (T, B, Hu+2, Wu+2, Cu) bf16, route (T, B, H+2, W+2, Cr) bf16, 2*Hu >= H, 2*Wu >= W (slice_like crops)
 *   out   (T, B, H+2, W+2, Cu+Cr) bf16 = [upsampled | route] per pixel, zero border
 *   requires Cu % 8 == 0 and Cr % 8 == 0.  Accounted under VY_K_LAYOUT. */
int vy_upsample_concat_bf16(const void *up, const void *route, int B, int T, int H, int W, int Hu, int Wu,
                            int Cu, int Cr, void *out, vy_stream_t stream);

/* Depthwise temporal merge of a window of T frames into one.
 * Replaces: _conv1d(channels, kernel=T, padding=0, strides=1), models/definitions/layers.py:50-60
 *           = Conv3D(kernel (T,1,1), groups=channels, use_bias=False) + BatchNorm + LeakyReLU(0.1), as
 *           applied per window by HDarknet (models/definitions/darknet/h_darknet.py:97-119).
 *   x      P layout (T, B, H+2, W+2, C) bf16;  w (C, T) fp32 = the reference weight (C, 1, T, 1, 1)
 *   scale, shift  folded inference BatchNorm per channel;  y P layout (1, B, H+2, W+2, C) bf16
 *   requires C % 8 == 0.  Accounted under VY_K_TEMPORAL_POOL. */
int vy_temporal_dwconv_bf16(const void *x, const float *w, const float *scale, const float *shift,
                            float leaky_slope, int B, int T, int H, int W, int C, void *y, vy_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VYOLO_H_ */
