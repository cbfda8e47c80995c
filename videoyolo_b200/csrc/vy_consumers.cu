// vy_consumers.cu -- what the reference does with net(x)'s (ids, scores, bboxes) on the host, on the device
// ("next" row f3):
//   vy_hier_nms_f32    hierarchical_nms of detect_yolo3.py:736-789 (with its iou helper :712-733)
//   vy_voc_match_f32   the per-image part of VOCMApMetric.update, metrics/pascalvoc.py:116-184
// Both are small, latency-bound bookkeeping kernels (a few hundred boxes per image): one warp / one CTA per image,
// the reference's float32 arithmetic in its own operation order (the library is built with -fmad=false).
#include "vy_common.cuh"
#include <math_constants.h>

// ------------------------------------------------------------------------------------------------
// hierarchical_nms.  boxes (B, N, 6) rows [cls, conf, x1, y1, x2, y2]; rows with cls < 0 are padding (the reference's
// prediction lists only hold valid rows, detect_yolo3.py:254-265).  Per image, exactly the python loop:
//   for box in sorted(boxes, key=cls, reverse=True):            stable: ties keep their input order
//       skip if conf < conf_thresh; cls = lifted[cls]            (the `while levels[cls] > level_thresh` walk, :766-767)
//       best = the FIRST new box with the largest iou > ov_thresh
//       none: append;  else if not branch[cls][best.cls]: append;  else if cls == best.cls: best.conf = max(best.conf, conf)
// One warp per image: the outer loop is sequential by nature, the scan over the new boxes is spread over the lanes.
// ------------------------------------------------------------------------------------------------
constexpr int HN_MAX = 1024;         // boxes per image

__device__ __forceinline__ float hn_iou(const float *bb, const float *bg) {
    // detect_yolo3.py:712-733, float32 (detect() hands it np.float32 scalars; `+ 1` is a weak python scalar)
    const float iw = __fadd_rn(__fsub_rn(fminf(bb[2], bg[2]), fmaxf(bb[0], bg[0])), 1.0f);
    const float ih = __fadd_rn(__fsub_rn(fminf(bb[3], bg[3]), fmaxf(bb[1], bg[1])), 1.0f);
    if (!(iw > 0.0f && ih > 0.0f)) return 0.0f;
    const float inter = __fmul_rn(iw, ih);
    const float aa = __fmul_rn(__fadd_rn(__fsub_rn(bb[2], bb[0]), 1.0f), __fadd_rn(__fsub_rn(bb[3], bb[1]), 1.0f));
    const float ab = __fmul_rn(__fadd_rn(__fsub_rn(bg[2], bg[0]), 1.0f), __fadd_rn(__fsub_rn(bg[3], bg[1]), 1.0f));
    const float ua = __fsub_rn(__fadd_rn(aa, ab), inter);
    return __fdiv_rn(inter, ua);
}

__global__ void __launch_bounds__(32)
vy_hier_nms_kernel(const float *__restrict__ boxes, int N, const int32_t *__restrict__ lifted,
                   const unsigned char *__restrict__ branch, int n_cls, float ov_thresh, float conf_thresh,
                   float *__restrict__ out, int32_t *__restrict__ counts) {
    extern __shared__ float hn_smem[];                  // in [N][6], new [N][6], order [N] (int)
    float *in = hn_smem, *nw = hn_smem + (size_t)N * 6;
    int *order = (int *)(nw + (size_t)N * 6);
    const int lane = threadIdx.x, b = blockIdx.x;
    const float *src = boxes + (size_t)b * N * 6;
    for (int i = lane; i < N * 6; i += 32) in[i] = src[i];
    __syncwarp();
    // stable sort by class descending among the valid rows: rank = #{class greater} + #{equal class, earlier row}
    int n_valid = 0;
    for (int i0 = 0; i0 < N; i0 += 32) {
        const int i = i0 + lane;
        const bool valid = i < N && in[i * 6] >= 0.0f;
        if (valid) {
            const float c = in[i * 6];
            int rank = 0;
            for (int j = 0; j < N; ++j) {
                const float cj = in[j * 6];
                if (cj >= 0.0f && (cj > c || (cj == c && j < i))) ++rank;
            }
            order[rank] = i;
        }
        n_valid += __popc(__ballot_sync(0xffffffffu, valid));
    }
    __syncwarp();
    int n_new = 0;                                      // warp-uniform
    for (int r = 0; r < n_valid; ++r) {
        const float *bx = in + order[r] * 6;
        const float conf = bx[1];
        if (conf < conf_thresh) continue;               // :763
        int cls = (int)bx[0];
        cls = (cls >= 0 && cls < n_cls) ? lifted[cls] : cls;          // :766-767
        // the first new box with the largest overlap above the threshold
        float best = 0.0f;
        int best_idx = -1;
        for (int j = lane; j < n_new; j += 32) {
            const float ov = hn_iou(bx + 2, nw + j * 6 + 2);
            if (ov > ov_thresh && ov > best) { best = ov; best_idx = j; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
            if (oi >= 0 && (best_idx < 0 || ob > best || (ob == best && oi < best_idx))) { best = ob; best_idx = oi; }
        }
        bool append = best_idx < 0;
        if (!append) {
            const int cb = (int)nw[best_idx * 6];
            const bool on_branch = cls >= 0 && cls < n_cls && cb >= 0 && cb < n_cls && branch[(size_t)cls * n_cls + cb] != 0;
            if (!on_branch) append = true;                                               // :783-784
            else if (cls == cb && lane == 0) nw[best_idx * 6 + 1] = fmaxf(nw[best_idx * 6 + 1], conf);   // :786-787
        }
        if (append) {
            if (lane < 6) nw[n_new * 6 + lane] = lane == 0 ? (float)cls : bx[lane];
            ++n_new;
        }
        __syncwarp();
    }
    float *dst = out + (size_t)b * N * 6;
    for (int i = lane; i < N * 6; i += 32) dst[i] = i < n_new * 6 ? nw[i] : -1.0f;
    if (lane == 0) counts[b] = n_new;
}

extern "C" int vy_hier_nms_f32(const float *boxes, int B, int N, const int32_t *lifted, const unsigned char *branch,
                               int n_cls, float ov_thresh, float conf_thresh, float *out, int32_t *counts,
                               vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!boxes || !lifted || !branch || !out || !counts || B < 1 || N < 1 || n_cls < 1)
        VY_FAIL(VY_EINVAL, "vy_hier_nms_f32: bad arguments");
    if (N > HN_MAX) VY_FAIL(VY_EUNSUPPORTED, "vy_hier_nms_f32: N=%d boxes per image exceed %d", N, HN_MAX);
    const size_t smem = (size_t)N * (12 * sizeof(float) + sizeof(int));
    VY_CUDA_CHECK(vy_ensure_dyn_smem((const void *)vy_hier_nms_kernel, smem));
    VY_KERNEL(VY_K_IOU, st, (vy_hier_nms_kernel<<<B, 32, smem, st>>>(boxes, N, lifted, branch, n_cls, ov_thresh, conf_thresh,
                                                                      out, counts)));
    VY_LAUNCH_CHECK("vy_hier_nms_kernel");
    return VY_OK;
}

// ------------------------------------------------------------------------------------------------
// VOC metric update for one batch.  dets (B, P, 6) rows [id, score, x1, y1, x2, y2] (-1 padding), gt_boxes (B, M, 4),
// gt_labels (B, M) (< 0: padding), gt_difficult (B, M) or null.  Per image (metrics/pascalvoc.py:116-184):
//   valid predictions ordered by (class ascending, score descending); among equal scores the later row first (what
//   `argsort()[::-1]` of a stable sort gives -- numpy's default sort leaves that order to the platform);
//   every prediction's ground truth = first argmax of bbox_iou over the ground truths of its class, -1 below iou_thresh;
//   match = -1 for a difficult ground truth, 1 for the first prediction (in that order) of a ground truth, else 0;
//   n_pos[class] = its non-difficult ground truths.
// Outputs, in that order, -2-padded: out_label / out_score / out_match (B, P); counts (B) valid predictions;
// n_pos (B, n_class).  One CTA per image.
// ------------------------------------------------------------------------------------------------
constexpr int VM_NT = 256, VM_PMAX = 1024, VM_MMAX = 512;

__global__ void __launch_bounds__(VM_NT)
vy_voc_match_kernel(const float *__restrict__ dets, const float *__restrict__ gt_boxes, const float *__restrict__ gt_labels,
                    const float *__restrict__ gt_difficult, int P, int M, int n_class, float iou_thresh,
                    int32_t *__restrict__ out_label, float *__restrict__ out_score, int32_t *__restrict__ out_match,
                    int32_t *__restrict__ counts, int32_t *__restrict__ n_pos) {
    __shared__ float4 g_box[VM_MMAX];
    __shared__ int g_lab[VM_MMAX], g_first[VM_MMAX];
    __shared__ unsigned char g_diff[VM_MMAX];
    __shared__ float p_score[VM_PMAX];
    __shared__ int p_lab[VM_PMAX];
    __shared__ int s_count;
    const int tid = threadIdx.x, b = blockIdx.x;
    const float *d = dets + (size_t)b * P * 6;
    if (tid == 0) s_count = 0;
    for (int j = tid; j < M; j += VM_NT) {
        const float *g = gt_boxes + ((size_t)b * M + j) * 4;
        g_box[j] = make_float4(g[0], g[1], g[2], g[3]);
        const float l = gt_labels[(size_t)b * M + j];
        g_lab[j] = l >= 0.0f ? (int)l : -1;                                   // :127-129
        g_diff[j] = gt_difficult ? (gt_difficult[(size_t)b * M + j] != 0.0f) : 0;
        g_first[j] = 0x7fffffff;
    }
    for (int i = tid; i < P; i += VM_NT) {
        const float l = d[i * 6];
        p_lab[i] = l >= 0.0f ? (int)l : -1;                                   // :118-120
        p_score[i] = d[i * 6 + 1];
    }
    for (int c = tid; c < n_class; c += VM_NT) n_pos[(size_t)b * n_class + c] = 0;
    __syncthreads();
    for (int j = tid; j < M; j += VM_NT)                                      // :149
        if (g_lab[j] >= 0 && g_lab[j] < n_class && !g_diff[j]) atomicAdd(&n_pos[(size_t)b * n_class + g_lab[j]], 1);
    // every prediction: its place in the (class, score desc, later row first) order and its ground truth
    for (int i0 = 0; i0 < P; i0 += VM_NT) {
        const int i = i0 + tid;
        int rank = -1, gi = -1;
        if (i < P && p_lab[i] >= 0) {
            const int l = p_lab[i];
            const float s = p_score[i];
            rank = 0;
            for (int j = 0; j < P; ++j) {
                const int lj = p_lab[j];
                if (lj < 0) continue;
                const float sj = p_score[j];
                if (lj < l || (lj == l && (sj > s || (sj == s && j > i)))) ++rank;
            }
            // bbox_iou (utils/bbox.py:11-38 == gluoncv's) against the ground truths of the class, first argmax (:166-169)
            const float4 a = make_float4(d[i * 6 + 2], d[i * 6 + 3], d[i * 6 + 4], d[i * 6 + 5]);
            const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
            float best = -CUDART_INF_F;
            for (int j = 0; j < M; ++j) {
                if (g_lab[j] != l) continue;
                const float4 g = g_box[j];
                const float tlx = fmaxf(a.x, g.x), tly = fmaxf(a.y, g.y), brx = fminf(a.z, g.z), bry = fminf(a.w, g.w);
                float area_i = __fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly));
                area_i = (tlx < brx && tly < bry) ? area_i : __fmul_rn(area_i, 0.0f);
                const float area_b = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
                const float iou = __fdiv_rn(area_i, __fsub_rn(__fadd_rn(area_a, area_b), area_i));
                if (gi < 0 || iou > best) { best = iou; gi = j; }
            }
            if (gi >= 0 && best < iou_thresh) gi = -1;
            if (gi >= 0) atomicMin(&g_first[gi], rank);
            atomicAdd(&s_count, 1);
        }
        __syncthreads();                                                      // g_first is complete (uniform trip count)
        if (rank >= 0) {
            int m = 0;
            if (gi >= 0) m = g_diff[gi] ? -1 : (g_first[gi] == rank ? 1 : 0);   // :173-184
            out_label[(size_t)b * P + rank] = p_lab[i];
            out_score[(size_t)b * P + rank] = p_score[i];
            out_match[(size_t)b * P + rank] = m;
        }
        __syncthreads();
    }
    const int n = s_count;
    for (int i = n + tid; i < P; i += VM_NT) {
        out_label[(size_t)b * P + i] = -2; out_score[(size_t)b * P + i] = -2.0f; out_match[(size_t)b * P + i] = -2;
    }
    if (tid == 0) counts[b] = n;
}

extern "C" int vy_voc_match_f32(const float *dets, const float *gt_boxes, const float *gt_labels, const float *gt_difficult,
                                int B, int P, int M, int n_class, float iou_thresh, int32_t *out_label, float *out_score,
                                int32_t *out_match, int32_t *counts, int32_t *n_pos, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!dets || !gt_boxes || !gt_labels || !out_label || !out_score || !out_match || !counts || !n_pos || B < 1 || P < 1 ||
        M < 1 || n_class < 1)
        VY_FAIL(VY_EINVAL, "vy_voc_match_f32: bad arguments");
    if (P > VM_PMAX || M > VM_MMAX) VY_FAIL(VY_EUNSUPPORTED, "vy_voc_match_f32: P <= %d and M <= %d", VM_PMAX, VM_MMAX);
    VY_KERNEL(VY_K_IOU, st, (vy_voc_match_kernel<<<B, VM_NT, 0, st>>>(dets, gt_boxes, gt_labels, gt_difficult, P, M, n_class,
                                                                      iou_thresh, out_label, out_score, out_match, counts, n_pos)));
    VY_LAUNCH_CHECK("vy_voc_match_kernel");
    return VY_OK;
}
