/*
 * oracle/vy_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the arithmetic on VideoYOLO's per-frame detection
 * post-processing path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (videoyolo_b200/) never does.
 *
 * PARITY: /root/reference has no tests, fixtures or golden vectors of its own, and
 * its arithmetic lives in un-vendored, un-pinned third-party packages
 * (mxnet-cu100, gluoncv; requirements.txt:1-2).  The decode (and the YOLOV3 tail
 * around it) is PINNED BY THE REFERENCE'S OWN CODE: yolo3.py imported unmodified
 * under a numpy-fp32 stand-in for MXNet's array ops and run in the build container
 * (tests/golden/make_golden.py -> tests/golden/decode_ref_*.npz); bbox_iou by
 * outputs of the reference's utils/bbox.py (tests/golden/bbox_iou_ref.npz).
 * box_nms stays PARITY UNPINNED by the reference (the operator's source is in
 * MXNet, not in /root/reference): restated from MXNet's published algorithm and
 * pinned by (a) the known-answer vectors of MXNet's public box_nms documentation
 * / unit test (tests/golden/box_nms_mxnet_doc.json), (b) an independent python
 * twin (oracle.box_nms_py), (c) torchvision's per-class nms.
 *
 * What each function follows:
 *   vy_oracle_decode_f32   models/definitions/yolo/yolo3.py:151-199
 *                          (dup models/definitions/yolo/yolo3_temporal.py:137-179),
 *                          scale concat yolo3.py:523, constants yolo3.py:46-74,416-417
 *   vy_oracle_box_nms_f32  the F.contrib.box_nms call at yolo3.py:525-530 (+5 sites);
 *                          operator semantics = MXNet _contrib_box_nms
 *                          (src/operator/contrib/bounding_box-inl.h, BoxNMSForward;
 *                          restated from the published algorithm)
 *   vy_oracle_bbox_iou_f64 utils/bbox.py:11-38
 *
 * Build: see oracle/Makefile  (-O2 -ffp-contract=off so that no FMA contraction
 * changes the fp32 association order the operator defines).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- tiny pthread parallel-for over the batch axis (frames are independent;
 * this is only so the CPU baseline can use all host cores) ---------------- */
static int g_threads = 1;
void vy_oracle_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int vy_oracle_get_threads(void) { return g_threads; }

typedef void (*vy_body_fn)(int b, void *ctx);
typedef struct { vy_body_fn fn; void *ctx; int n; volatile int *next; } vy_pf_t;
static void *vy_pf_worker(void *p)
{
    vy_pf_t *w = (vy_pf_t *)p;
    for (;;) {
        int b = __sync_fetch_and_add(w->next, 1);
        if (b >= w->n) break;
        w->fn(b, w->ctx);
    }
    return NULL;
}
static void vy_parallel_for(int n, vy_body_fn fn, void *ctx)
{
    int nt = g_threads < n ? g_threads : n;
    volatile int next = 0;
    vy_pf_t w = { fn, ctx, n, &next };
    if (nt <= 1) { vy_pf_worker(&w); return; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nt - 1; ++i)
        if (pthread_create(&th[started], NULL, vy_pf_worker, &w) == 0) ++started;
    vy_pf_worker(&w);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

#define VY_CORNER 0
#define VY_CENTER 1

/* mshadow_op::sigmoid: 1/(1+exp(-x)) in fp32 (yolo3.py:172,174,175 F.sigmoid) */
static inline float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }

/*
 * Decode ONE scale into the concatenated detection tensor.
 *   head : (B, A*P, H, W) fp32 NCHW, channel = a*P + p          (yolo3.py:158-160)
 *   dets : (B, R_total, 6) fp32; this scale's rows start at row_offset
 *          row = row_offset + c*(H*W*A) + (y*W+x)*A + a          (yolo3.py:191-197, :523)
 *          agnostic: row = row_offset + (y*W+x)*A + a, id 0, score = objectness (:184-188)
 */
typedef struct {
    const float *head; int A, C, H, W; float stride; const float *anchors;
    int agnostic; float *dets; long R_total, row_offset;
} vy_dec_ctx;

static void vy_decode_frame(int b, void *p)
{
    const vy_dec_ctx *k = (const vy_dec_ctx *)p;
    const int A = k->A, C = k->C, H = k->H, W = k->W;
    const int P = 5 + C;
    const long HW = (long)H * W;
    const long n_s = HW * A;
    const float stride = k->stride;
    const float *anchors = k->anchors;
    const float *hb = k->head + (long)b * A * P * HW;
    float *db = k->dets + (long)b * k->R_total * 6;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const long pos = (long)y * W + x;
            for (int a = 0; a < A; ++a) {
                const float *pa = hb + (long)a * P * HW + pos;
                const float tx = pa[0 * HW], ty = pa[1 * HW];
                const float tw = pa[2 * HW], th = pa[3 * HW];
                const float to = pa[4 * HW];
                /* yolo3.py:172  (sigmoid(raw_centers) + offsets) * stride */
                const float cx = (sigmoid_f32(tx) + (float)x) * stride;
                const float cy = (sigmoid_f32(ty) + (float)y) * stride;
                /* yolo3.py:173  exp(raw_scales) * anchors */
                const float bw = expf(tw) * anchors[2 * a + 0];
                const float bh = expf(th) * anchors[2 * a + 1];
                /* yolo3.py:174 */
                const float conf = sigmoid_f32(to);
                /* yolo3.py:176-177  wh = scales / 2.0 ; (centers - wh, centers + wh) */
                const float hw = bw / 2.0f, hh = bh / 2.0f;
                const float x1 = cx - hw, y1 = cy - hh, x2 = cx + hw, y2 = cy + hh;
                if (k->agnostic) {
                    float *r = db + (k->row_offset + pos * A + a) * 6;
                    r[0] = 0.0f; r[1] = conf; r[2] = x1; r[3] = y1; r[4] = x2; r[5] = y2;
                    continue;
                }
                for (int c = 0; c < C; ++c) {
                    /* yolo3.py:175  sigmoid(class_pred) * confidence */
                    const float sc = sigmoid_f32(pa[(long)(5 + c) * HW]) * conf;
                    float *r = db + (k->row_offset + (long)c * n_s + pos * A + a) * 6;
                    r[0] = (float)c; r[1] = sc; r[2] = x1; r[3] = y1; r[4] = x2; r[5] = y2;
                }
            }
        }
    }
}

void vy_oracle_decode_f32(const float *head, int B, int A, int C, int H, int W,
                          float stride, const float *anchors /* A*2: w,h */,
                          int agnostic, float *dets, long R_total, long row_offset)
{
    vy_dec_ctx k = { head, A, C, H, W, stride, anchors, agnostic, dets, R_total, row_offset };
    vy_parallel_for(B, vy_decode_frame, &k);
}

/* ---- box_nms ---------------------------------------------------------- */

typedef struct { float score; int32_t row; } cand_t;

/* upstream SortByKey(scores, index, is_ascend=false) is a std::stable_sort with
 * '>' on the score: ties keep ascending original row.  merge sort == stable. */
static void merge_sort_desc(cand_t *a, cand_t *tmp, long n)
{
    if (n < 2) return;
    long h = n / 2;
    merge_sort_desc(a, tmp, h);
    merge_sort_desc(a + h, tmp, n - h);
    long i = 0, j = h, k = 0;
    while (i < h && j < n) {
        /* take right only when strictly greater: keeps left (lower row) on ties */
        if (a[j].score > a[i].score) tmp[k++] = a[j++]; else tmp[k++] = a[i++];
    }
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, (size_t)n * sizeof(cand_t));
}

/* upstream BoxArea (bounding_box-inl.h): corner -> (x2-x1)*(y2-y1); center -> w*h;
 * zero when width<0 or height<0 */
static inline float box_area(const float *b, int fmt)
{
    float w, h;
    if (fmt == VY_CORNER) { w = b[2] - b[0]; h = b[3] - b[1]; }
    else                  { w = b[2];        h = b[3]; }
    if (w < 0 || h < 0) return 0.0f;
    return w * h;
}

/* upstream Intersect: 1-D overlap of [a0,a2] / [b0,b2] (stride-2 pick: x then y) */
static inline float intersect_1d(const float *a, const float *b, int fmt)
{
    float a1 = a[0], a2 = a[2], b1 = b[0], b2 = b[2], w;
    if (fmt == VY_CORNER) {
        float left = a1 > b1 ? a1 : b1;
        float right = a2 < b2 ? a2 : b2;
        w = right - left;
    } else {
        float aw = a2 / 2, bw = b2 / 2;
        float al = a1 - aw, ar = a1 + aw, bl = b1 - bw, br = b1 + bw;
        float left = bl > al ? bl : al;
        float right = br < ar ? br : ar;
        w = right - left;
    }
    return w > 0 ? w : 0.0f;
}

/*
 * data/out: (B, R, W_elem); record: (B, R) int32 = source row of each output
 * row, -1 for padding (upstream's hidden second output).
 * Per image:
 *  1 valid = score > valid_thresh (strict); and id != background_id when both >= 0
 *  2 stable sort by score descending
 *  3 keep first k = (topk<0 ? R : min(R,topk))
 *  4/5 greedy suppression, iou = inter/(area_ref+area_pos-inter) > overlap_thresh (strict)
 *  6 survivors compacted to the front in score order, rest -1
 *  out_format != in_format converts the 4 coords of surviving rows whose first
 *  coordinate is >= 0 (upstream corner_to_center / center_to_corner early-return quirk).
 */
typedef struct {
    const float *data; long R; int W_elem; float overlap_thresh, valid_thresh; long k;
    int coord_start, score_index, id_index, background_id, force_suppress, in_format, out_format;
    float *out; int32_t *record; volatile int rc;
} vy_nms_ctx;

static void vy_nms_frame(int b, void *p)
{
    vy_nms_ctx *q = (vy_nms_ctx *)p;
    const long R = q->R; const int W_elem = q->W_elem;
    const int coord_start = q->coord_start, id_index = q->id_index;
    const float *d = q->data + (long)b * R * W_elem;
    float *o = q->out + (long)b * R * W_elem;
    int32_t *rec = q->record ? q->record + (long)b * R : NULL;
    for (long i = 0; i < R * W_elem; ++i) o[i] = -1.0f;
    if (rec) for (long i = 0; i < R; ++i) rec[i] = -1;

    cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * (size_t)(R > 0 ? R : 1) * 2);
    if (!cand) { q->rc = -2; return; }
    long n = 0;
    for (long i = 0; i < R; ++i) {
        const float s = d[i * W_elem + q->score_index];
        if (!(s > q->valid_thresh)) continue;                      /* strict, NaN dropped */
        if (id_index >= 0 && q->background_id >= 0 &&
            (int)d[i * W_elem + id_index] == q->background_id) continue;
        cand[n].score = s; cand[n].row = (int32_t)i; ++n;
    }
    merge_sort_desc(cand, cand + R, n);
    const long m = n < q->k ? n : q->k;
    float *area = (float *)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
    char *dead = (char *)calloc((size_t)(m > 0 ? m : 1), 1);
    if (!area || !dead) { q->rc = -2; free(cand); free(area); free(dead); return; }
    for (long i = 0; i < m; ++i)
        area[i] = box_area(d + (long)cand[i].row * W_elem + coord_start, q->in_format);
    for (long r = 0; r < m; ++r) {
        if (dead[r]) continue;                                     /* suppressed never suppress */
        const float *rb = d + (long)cand[r].row * W_elem;
        for (long t = r + 1; t < m; ++t) {
            if (dead[t]) continue;
            const float *pb = d + (long)cand[t].row * W_elem;
            if (!q->force_suppress && id_index >= 0) {
                if ((int)rb[id_index] != (int)pb[id_index]) continue;   /* different class */
            }
            float inter = intersect_1d(rb + coord_start, pb + coord_start, q->in_format);
            inter *= intersect_1d(rb + coord_start + 1, pb + coord_start + 1, q->in_format);
            const float iou = inter / (area[r] + area[t] - inter);
            if (iou > q->overlap_thresh) dead[t] = 1;              /* strict; NaN keeps */
        }
    }
    long w = 0;
    for (long i = 0; i < m; ++i) {
        if (dead[i]) continue;
        float *dst = o + w * W_elem;
        memcpy(dst, d + (long)cand[i].row * W_elem, sizeof(float) * (size_t)W_elem);
        if (q->in_format != q->out_format) {
            float *c = dst + coord_start;
            if (!(c[0] < 0)) {
                if (q->out_format == VY_CENTER) {          /* corner_to_center */
                    float l = c[0], t = c[1], r2 = c[2], bt = c[3];
                    c[0] = (l + r2) / 2; c[1] = (t + bt) / 2; c[2] = r2 - l; c[3] = bt - t;
                } else {                                    /* center_to_corner */
                    float x = c[0], y = c[1], hw = c[2] / 2, hh = c[3] / 2;
                    c[0] = x - hw; c[1] = y - hh; c[2] = x + hw; c[3] = y + hh;
                }
            }
        }
        if (rec) rec[w] = cand[i].row;
        ++w;
    }
    free(cand); free(area); free(dead);
}

int vy_oracle_box_nms_f32(const float *data, int B, long R, int W_elem,
                          float overlap_thresh, float valid_thresh, int topk,
                          int coord_start, int score_index, int id_index,
                          int background_id, int force_suppress,
                          int in_format, int out_format,
                          float *out, int32_t *record)
{
    if (W_elem < coord_start + 4 || score_index >= W_elem || score_index < 0 ||
        id_index >= W_elem) return -1;
    vy_nms_ctx q = { data, R, W_elem, overlap_thresh, valid_thresh,
                     (topk < 0) ? R : (topk < R ? topk : R),
                     coord_start, score_index, id_index, background_id, force_suppress,
                     in_format, out_format, out, record, 0 };
    vy_parallel_for(B, vy_nms_frame, &q);
    return q.rc;
}

/* utils/bbox.py:11-38, evaluated in float64 exactly as numpy does for float64 inputs */
void vy_oracle_bbox_iou_f64(const double *a, int N, const double *b, int M,
                            double offset, double *out)
{
    for (int i = 0; i < N; ++i) {
        const double *pa = a + 4 * i;
        const double area_a = (pa[2] - pa[0] + offset) * (pa[3] - pa[1] + offset);
        for (int j = 0; j < M; ++j) {
            const double *pb = b + 4 * j;
            const double tlx = pa[0] > pb[0] ? pa[0] : pb[0];
            const double tly = pa[1] > pb[1] ? pa[1] : pb[1];
            const double brx = pa[2] < pb[2] ? pa[2] : pb[2];
            const double bry = pa[3] < pb[3] ? pa[3] : pb[3];
            const double valid = (tlx < brx && tly < bry) ? 1.0 : 0.0;
            const double area_i = ((brx - tlx + offset) * (bry - tly + offset)) * valid;
            const double area_b = (pb[2] - pb[0] + offset) * (pb[3] - pb[1] + offset);
            out[(long)i * M + j] = area_i / (area_a + area_b - area_i);
        }
    }
}

int vy_oracle_version(void) { return 1; }
