// vy_iou.cu -- pairwise IoU, utils/bbox.py:11-38 (same operation order as the numpy source).
#include "vy_common.cuh"
#include <math_constants.h>

template <typename T>
__global__ void vy_bbox_iou_kernel(const T *__restrict__ a, int N, int lda, const T *__restrict__ b, int M,
                                   int ldb, T offset, T *__restrict__ out) {
    const long long total = (long long)N * M;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / M), j = (int)(idx % M);
        const T *pa = a + (size_t)i * lda, *pb = b + (size_t)j * ldb;
        const T tlx = pa[0] > pb[0] ? pa[0] : pb[0], tly = pa[1] > pb[1] ? pa[1] : pb[1];   // :32
        const T brx = pa[2] < pb[2] ? pa[2] : pb[2], bry = pa[3] < pb[3] ? pa[3] : pb[3];   // :33
        const T valid = (tlx < brx && tly < bry) ? (T)1 : (T)0;
        const T area_i = ((brx - tlx + offset) * (bry - tly + offset)) * valid;            // :35
        const T area_a = (pa[2] - pa[0] + offset) * (pa[3] - pa[1] + offset);              // :36
        const T area_b = (pb[2] - pb[0] + offset) * (pb[3] - pb[1] + offset);              // :37
        out[idx] = area_i / (area_a + area_b - area_i);                                    // :38
    }
}

template <typename T>
static int launch_iou(const T *a, int N, int lda, const T *b, int M, int ldb, T offset, T *out, vy_stream_t st) {
    if (N < 0 || M < 0 || lda < 4 || ldb < 4) VY_FAIL(VY_EINVAL, "bbox_iou: boxes need >= 4 columns");   // :29-30
    if (N == 0 || M == 0) return VY_OK;
    if (!a || !b || !out) VY_FAIL(VY_EINVAL, "bbox_iou: null pointer");
    const long long total = (long long)N * M;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)vy_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    VY_KERNEL(VY_K_IOU, (cudaStream_t)st,
              (vy_bbox_iou_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)st>>>(a, N, lda, b, M, ldb, offset, out)));
    VY_LAUNCH_CHECK("vy_bbox_iou_kernel");
    return VY_OK;
}

extern "C" int vy_bbox_iou_f32(const float *a, int N, int lda, const float *b, int M, int ldb, float offset,
                               float *out, vy_stream_t st) {
    return launch_iou<float>(a, N, lda, b, M, ldb, offset, out, st);
}
extern "C" int vy_bbox_iou_f64(const double *a, int N, int lda, const double *b, int M, int ldb, double offset,
                               double *out, vy_stream_t st) {
    return launch_iou<double>(a, N, lda, b, M, ldb, offset, out, st);
}

// ------------------------------------------------------------------------------------------------
// Batched pairwise IoU of the dynamic-target step (models/definitions/yolo/yolo_target.py:171,202-204:
// batch_ious = BBoxBatchIOU()(box_preds, gt_boxes); ious_max = batch_ious.max(-1); objness = -(ious_max > thr)).
// BBoxBatchIOU is gluoncv.nn.bbox (not vendored in the reference); its arithmetic, restated in oracle/:
//   iw = clip(min(ar, br) - max(al, bl) + offset, 0, 65504), ih likewise, i = iw * ih,
//   area = (r - l + offset) * (b - t + offset), iou = i / (area_a + area_b - i + eps)      -- corner format.
// One thread per predicted box: the M ground-truth boxes of its image sit in shared memory; the (B, N, M) tensor
// is optional -- the maximum over M and the ignore mask are produced in the same pass.
constexpr int BIOU_NT = 256;
constexpr int BIOU_MTILE = 256;
__global__ void __launch_bounds__(BIOU_NT)
vy_bbox_batch_iou_kernel(const float *__restrict__ a, const float *__restrict__ b, int N, int M, float offset,
                         float eps, float ignore_thresh, float *__restrict__ ious, float *__restrict__ ious_max,
                         float *__restrict__ objness) {
    __shared__ float4 gt[BIOU_MTILE];
    __shared__ float gt_area[BIOU_MTILE];
    const int img = blockIdx.y;
    const int n = blockIdx.x * BIOU_NT + threadIdx.x;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) bx = *reinterpret_cast<const float4 *>(a + ((size_t)img * N + n) * 4);
    const float area_a = (bx.z - bx.x + offset) * (bx.w - bx.y + offset);
    float best = -CUDART_INF_F;
    for (int m0 = 0; m0 < M; m0 += BIOU_MTILE) {
        const int mt = min(BIOU_MTILE, M - m0);
        __syncthreads();
        for (int j = threadIdx.x; j < mt; j += BIOU_NT) {
            const float4 g = *reinterpret_cast<const float4 *>(b + ((size_t)img * M + m0 + j) * 4);
            gt[j] = g;
            gt_area[j] = (g.z - g.x + offset) * (g.w - g.y + offset);
        }
        __syncthreads();
        if (n < N) {
            float *o = ious ? ious + ((size_t)img * N + n) * M + m0 : nullptr;
            for (int j = 0; j < mt; ++j) {
                const float4 g = gt[j];
                const float left = fmaxf(bx.x, g.x), right = fminf(bx.z, g.z);
                const float top = fmaxf(bx.y, g.y), bot = fminf(bx.w, g.w);
                const float iw = fminf(fmaxf(right - left + offset, 0.0f), 6.55040e+04f);
                const float ih = fminf(fmaxf(bot - top + offset, 0.0f), 6.55040e+04f);
                const float inter = iw * ih;
                const float v = inter / (area_a + gt_area[j] - inter + eps);
                if (o) o[j] = v;
                best = fmaxf(best, v);
            }
        }
    }
    if (n < N) {
        if (ious_max) ious_max[(size_t)img * N + n] = best;
        if (objness) objness[(size_t)img * N + n] = best > ignore_thresh ? -1.0f : 0.0f;   // yolo_target.py:204
    }
}

extern "C" int vy_bbox_batch_iou_f32(const float *a, const float *b, int B, int N, int M, float offset, float eps,
                                     float ignore_thresh, float *ious, float *ious_max, float *objness,
                                     vy_stream_t st) {
    if (B < 0 || N < 0 || M < 1) VY_FAIL(VY_EINVAL, "bbox_batch_iou: need B, N >= 0 and M >= 1");
    if (B == 0 || N == 0) return VY_OK;
    if (!a || !b || (!ious && !ious_max && !objness)) VY_FAIL(VY_EINVAL, "bbox_batch_iou: null pointer");
    if ((((uintptr_t)a) | ((uintptr_t)b)) & 15) VY_FAIL(VY_EALIGN, "bbox_batch_iou: boxes must be 16-byte aligned");
    if (B > 65535) VY_FAIL(VY_EINVAL, "bbox_batch_iou: B > 65535");
    const dim3 grid((unsigned)((N + BIOU_NT - 1) / BIOU_NT), (unsigned)B);
    VY_KERNEL(VY_K_IOU, (cudaStream_t)st,
              (vy_bbox_batch_iou_kernel<<<grid, BIOU_NT, 0, (cudaStream_t)st>>>(a, b, N, M, offset, eps, ignore_thresh,
                                                                                ious, ious_max, objness)));
    VY_LAUNCH_CHECK("vy_bbox_batch_iou_kernel");
    return VY_OK;
}

// ------------------------------------------------------------------------------------------------
// The consumer step of detect() / validate() on the device (detect_yolo3.py:226,254-265,
// train_yolov3.py:477): bboxes.clip(0, W) for every row, then per image the rows with id >= 0 and their
// boxes divided by the input size.  dets: (B, P, 6) rows [id, score, x1, y1, x2, y2] as written by
// vy_decode_nms_f32 (survivors first, -1 padding).  One warp per image.
__global__ void vy_detect_consume_kernel(const float *__restrict__ dets, int B, int P, float clip_hi, float norm,
                                         float *__restrict__ clipped, float *__restrict__ normed,
                                         int32_t *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int b = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= B) return;
    int n = 0;
    for (int r0 = 0; r0 < P; r0 += 32) {
        const int r = r0 + lane;
        bool valid = false;
        if (r < P) {
            const float *d = dets + ((size_t)b * P + r) * 6;
            valid = d[0] >= 0.0f;                                            // detect_yolo3.py:256
            float c[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) c[k] = fminf(fmaxf(d[2 + k], 0.0f), clip_hi);      // :226  clip(0, W)
            float *oc = clipped + ((size_t)b * P + r) * 4;
            float *on = normed + ((size_t)b * P + r) * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) { oc[k] = c[k]; on[k] = valid ? __fdiv_rn(c[k], norm) : -1.0f; }   // :257
        }
        n += __popc(__ballot_sync(0xffffffffu, valid));
    }
    if (lane == 0) counts[b] = n;
}

extern "C" int vy_detect_consume_f32(const float *dets, int B, int P, float clip_hi, float norm, float *clipped,
                                     float *normed, int32_t *counts, vy_stream_t st) {
    if (!dets || !clipped || !normed || !counts || B < 1 || P < 1) VY_FAIL(VY_EINVAL, "vy_detect_consume_f32: bad arguments");
    if (!(norm > 0.0f)) VY_FAIL(VY_EINVAL, "vy_detect_consume_f32: norm must be > 0");
    const int warps_per_cta = 4;
    const int blocks = (B + warps_per_cta - 1) / warps_per_cta;
    VY_KERNEL(VY_K_IOU, (cudaStream_t)st, (vy_detect_consume_kernel<<<blocks, warps_per_cta * 32, 0, (cudaStream_t)st>>>(
        dets, B, P, clip_hi, norm, clipped, normed, counts)));
    VY_LAUNCH_CHECK("vy_detect_consume_kernel");
    return VY_OK;
}

// ------------------------------------------------------------------------------------------------
// Anchor matching of the prefetch target generator (yolo_target.py:86-93): every ground-truth box, moved to the
// origin (shift_gt_boxes :89), against the zero-centred anchors (anchor_boxes / bbox2corner :90-91) with
// nd.contrib.box_iou (:92), then argmax over the anchors (:94).  MXNet's box_iou (corner format): per axis
// w = min(right) - max(left), 0 if negative; i = w_x * w_y; 0 if i <= 0, else i / (area_l + area_r - i) with
// area = 0 for a negative extent (bounding_box-inl.h: Intersect / BoxArea / compute_overlap); argmax = first maximum.
// One thread per ground-truth box; fp32, un-contracted, in that operation order.
// ------------------------------------------------------------------------------------------------
constexpr int AM_MAX_A = 32;
__global__ void __launch_bounds__(128)
vy_anchor_match_kernel(const float *__restrict__ gt, long long n, int M, const float *__restrict__ anchors, int A,
                       int32_t *__restrict__ matches, float *__restrict__ ious) {
    __shared__ float aw[AM_MAX_A], ah[AM_MAX_A];
    if (threadIdx.x < A) { aw[threadIdx.x] = anchors[2 * threadIdx.x]; ah[threadIdx.x] = anchors[2 * threadIdx.x + 1]; }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 g = *(const float4 *)(gt + 4 * i);
    // bbox2center (BBoxCornerToCenter): width = xmax - xmin, height = ymax - ymin; the shifted box is (-w/2, -h/2, w/2, h/2)
    const float gw = __fsub_rn(g.z, g.x), gh = __fsub_rn(g.w, g.y);
    const float gl = __fmul_rn(-0.5f, gw), gt_ = __fmul_rn(-0.5f, gh), gr = __fmul_rn(0.5f, gw), gb = __fmul_rn(0.5f, gh);
    const float gwx = __fsub_rn(gr, gl), ghy = __fsub_rn(gb, gt_);
    const float garea = (gwx < 0.0f || ghy < 0.0f) ? 0.0f : __fmul_rn(gwx, ghy);
    const long long b = i / M, m = i - b * M;
    int best = 0;
    float best_v = -CUDART_INF_F;
    for (int a = 0; a < A; ++a) {
        // bbox2corner (BBoxCenterToCorner) of (0, 0, aw, ah): (-aw/2, -ah/2, aw/2, ah/2)
        const float hw = __fdiv_rn(aw[a], 2.0f), hh = __fdiv_rn(ah[a], 2.0f);
        const float al = __fsub_rn(0.0f, hw), at = __fsub_rn(0.0f, hh), ar = __fadd_rn(0.0f, hw), ab = __fadd_rn(0.0f, hh);
        float wx = __fsub_rn(fminf(ar, gr), fmaxf(al, gl));
        wx = wx < 0.0f ? 0.0f : wx;
        float wy = __fsub_rn(fminf(ab, gb), fmaxf(at, gt_));
        wy = wy < 0.0f ? 0.0f : wy;
        const float inter = __fmul_rn(wx, wy);
        float v = 0.0f;
        if (inter > 0.0f) {
            const float awx = __fsub_rn(ar, al), ahy = __fsub_rn(ab, at);
            const float aarea = (awx < 0.0f || ahy < 0.0f) ? 0.0f : __fmul_rn(awx, ahy);
            v = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, garea), inter));
        }
        if (ious) ious[(b * A + a) * M + m] = v;              // (B, A, M): ious.transpose((1, 0, 2)) of the (A, B, M) result
        if (v > best_v) { best_v = v; best = a; }             // first maximum (NaN never wins)
    }
    matches[i] = best;
}

extern "C" int vy_anchor_match_f32(const float *gt_boxes, int B, int M, const float *anchors, int A, int32_t *matches,
                                   float *ious, vy_stream_t st) {
    if (B < 0 || M < 0 || A < 1 || A > AM_MAX_A) VY_FAIL(VY_EINVAL, "anchor_match: need B, M >= 0 and 1 <= A <= %d", AM_MAX_A);
    if (B == 0 || M == 0) return VY_OK;
    if (!gt_boxes || !anchors || !matches) VY_FAIL(VY_EINVAL, "anchor_match: null pointer");
    if (((uintptr_t)gt_boxes) & 15) VY_FAIL(VY_EALIGN, "anchor_match: gt_boxes must be 16-byte aligned");
    const long long n = (long long)B * M;
    VY_KERNEL(VY_K_IOU, (cudaStream_t)st,
              (vy_anchor_match_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)st>>>(gt_boxes, n, M, anchors, A,
                                                                                                  matches, ious)));
    VY_LAUNCH_CHECK("vy_anchor_match_kernel");
    return VY_OK;
}
