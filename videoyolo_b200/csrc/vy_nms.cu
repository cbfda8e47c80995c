// vy_nms.cu -- candidate selection + box_nms kernels (fused-from-heads and generic-from-rows).
//
// Pipeline per call (two launches, no host sync):
//   1. *_select_kernel  : streams the input ONCE (head maps: class/objectness planes only;
//                         rows: the score column), keeps per-CTA top-K candidates in shared
//                         memory under a rising threshold shared per image through global
//                         memory, and appends the few survivors to a per-image list.
//   2. nms_finalize     : one CTA per image: exact top-K + sort of the list, decode of the K
//                         boxes, IoU suppression bitmask in warp tiles (ballot), greedy scan,
//                         compaction, write of the (post_nms, W) rows + kept source rows.
// Semantics follow MXNet _contrib_box_nms as called at yolo3.py:525-530.
#include "vy_select.cuh"
#include <math_constants.h>

// ------------------------------------------------------------------------------------------------
// tiling plan shared by workspace sizing and launch
// ------------------------------------------------------------------------------------------------
struct SelPlan {
    int K;                      // min(topk, R)
    int tiles_per_frame;
    int csplit, cper;           // heads: class range split
    int tile_begin[VY_MAX_SCALES + 1];
    int chunks[VY_MAX_SCALES];  // heads: position chunks per (scale, anchor)
    long long rows_per_tile;    // rows: rows per tile
    int list_cap;               // keys per image in the global list
    float valid_thresh;
};

struct SelGlobal {              // workspace views
    u64 *thr;                   // [B]
    int *count;                 // [B]
    u64 *list;                  // [B][list_cap]
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t sel_workspace_layout(int B, int list_cap, SelGlobal *g, void *base) {
    size_t off = 0;
    const size_t o_thr = off;   off = align_up(off + sizeof(u64) * (size_t)B, 256);
    const size_t o_cnt = off;   off = align_up(off + sizeof(int) * (size_t)B, 256);
    const size_t o_list = off;  off = align_up(off + sizeof(u64) * (size_t)B * (size_t)list_cap, 256);
    if (g && base) {
        g->thr = (u64 *)((char *)base + o_thr);
        g->count = (int *)((char *)base + o_cnt);
        g->list = (u64 *)((char *)base + o_list);
    }
    return off;
}
static inline size_t sel_header_bytes(int B) {   // the part that must be zeroed per call
    return align_up(align_up(sizeof(u64) * (size_t)B, 256) + sizeof(int) * (size_t)B, 256);
}

// ------------------------------------------------------------------------------------------------
// device: threshold sharing + flush
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Append the CTA's surviving keys to the image's global list and publish its threshold.
__device__ void sel_flush(SelBuf &S, int K, u64 *g_thr_b, int *g_count_b, u64 *g_list_b, int list_cap) {
    const int tid = threadIdx.x, lane = tid & 31;
    __syncthreads();
    int n = S.count;
    if (n > K + (K >> 2)) n = sel_compact(S, n, K, false);
    if (tid == 0) {
        const u64 mine = S.thr;
        const u64 old = mine ? atomicMax(g_thr_b, mine) : ld_relaxed_u64(g_thr_b);
        S.thr = old > mine ? old : mine;
    }
    __syncthreads();
    const u64 thr = S.thr;
    for (int base = 0; base < n; base += blockDim.x) {
        const int idx = base + tid;
        const u64 key = idx < n ? S.keys[idx] : 0ull;
        const bool p = idx < n && key >= thr;
        const u32 m = __ballot_sync(0xffffffffu, p);
        if (m) {
            const int leader = __ffs(m) - 1;
            int pos = 0;
            if (lane == leader) pos = atomicAdd(g_count_b, __popc(m));
            pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(m & ((1u << lane) - 1u));
            if (p && pos < list_cap) g_list_b[pos] = key;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// device: generic adaptive block loop
//   Src provides: int n_iters; int max_push_per_iter;
//                 void refresh(u64 thr);                       recompute prefilter state
//                 void run(SelBuf&, int it0, int it1, u64 thr);  stream iterations [it0,it1)
// Between CTA barriers a block of U iterations is streamed; U adapts to the observed push rate
// so that the shared buffer cannot overflow in the steady state; if it still does (cold start,
// adversarial order) the block's pushes are discarded, the buffer is compacted and the block is
// replayed with a smaller U.  U == 1 always fits: max_push_per_iter + K <= SEL_CAP.
// ------------------------------------------------------------------------------------------------
template <class Src>
__device__ void sel_stream(SelBuf &S, Src &src, int K, u64 *g_thr_b) {
    constexpr int UMAX = 16;
    const int tid = threadIdx.x;
    __syncthreads();
    int it = 0;
    int cnt0 = S.count;                              // CTA-uniform running count
    int U = (S.thr == 0ull) ? 1 : 4;
    u64 thr_seen = ~0ull, published = S.thr;
    __syncthreads();                                 // nobody pushes before everybody has read
    while (it < src.n_iters) {                       // CTA-uniform loop
        const u64 thr = S.thr;
        u64 gthr = 0;
        if (tid == 0) gthr = ld_relaxed_u64(g_thr_b);          // consumed after the block
        if (thr != thr_seen) { src.refresh(thr); thr_seen = thr; }
        const int Ub = min(U, src.n_iters - it);
        src.run(S, it, it + Ub, thr);
        __syncthreads();                              // pushes of this block complete
        if (tid == 0) {
            const int c = S.count;
            S.flag = c > SEL_CAP;
            if (S.flag) S.count = cnt0;
            S.snap = S.flag ? cnt0 : c;
            if (gthr > S.thr) S.thr = gthr;
        }
        __syncthreads();
        if (S.flag) {                                 // overflow: replay with a smaller block
            cnt0 = sel_compact(S, cnt0, K, true);
            U = max(1, Ub >> 1);
            continue;
        }
        int cnt = S.snap;
        const int pushed = cnt - cnt0;
        it += Ub;
        if (cnt > K + (K >> 2) && cnt >= SEL_CAP / 2) {
            cnt = sel_compact(S, cnt, K, false);
            const u64 nthr = S.thr;                  // stable: thread 0 next writes it after a barrier
            if (tid == 0 && nthr > published) atomicMax(g_thr_b, nthr);
            published = nthr;
        }
        const int rate = (pushed + Ub - 1) / Ub;
        U = min(UMAX, max(1, (SEL_CAP - cnt) / (2 * rate + 1)));
        cnt0 = cnt;
    }
}

// ------------------------------------------------------------------------------------------------
// source 1: YOLO head maps (fused decode).  A tile = (image b, scale s, anchor a, a chunk of
// SEL_NT*VEC positions, a class range).  Each thread owns VEC consecutive positions and walks the
// class planes with one 128-bit load per plane.  score = sigmoid(t_c)*conf >= smin is tested in
// the logit domain against a per-box bound (one compare per element); only elements that pass
// pay for the exact sigmoid.
// ------------------------------------------------------------------------------------------------
// conservative logit bound: score(t) >= smin  ==>  t >= vy_tcmin(smin, conf)
__device__ __forceinline__ float vy_tcmin(float smin, float conf) {
    if (!(smin > 0.0f)) return -CUDART_INF_F;
    const float q = __fmul_rn(__fdiv_rn(smin, conf), 1.0f - 1e-4f);
    if (!(q < 1.0f)) return CUDART_INF_F;            // also conf == 0 / NaN: no class can pass
    return logf(__fdiv_rn(q, 1.0f - q)) - 1e-4f;
}

template <int VEC>
struct HeadSrc {
    int n_iters, max_push_per_iter;
    const float *plane;          // -> channel a*P + 0 at this thread's first position
    size_t HW;
    int c0;
    bool active;
    float conf[VEC], tcmin[VEC];
    u32 row0;                    // row_off + pos0*A + a  (add c*n_s + v*A)
    u32 n_s, A;
    float valid_thresh;

    __device__ void refresh(u64 thr) {
        const float ts = thr ? vy_key_score(thr) : valid_thresh;
        const float smin = fmaxf(ts, valid_thresh);
#pragma unroll
        for (int v = 0; v < VEC; ++v) tcmin[v] = active ? vy_tcmin(smin, conf[v]) : CUDART_INF_F;
    }
    __device__ __forceinline__ void hit(SelBuf &S, float t, int v, int c, u64 thr) {
        const float s = vy_score(t, conf[v]);
        if (s > valid_thresh) {
            const u64 key = vy_make_key(s, row0 + (u32)c * n_s + (u32)v * A);
            if (key >= thr) sel_push(S, key);
        }
    }
    __device__ void run(SelBuf &S, int it0, int it1, u64 thr) {
        if (!active) return;
        for (int it = it0; it < it1; it += 4) {
            if (VEC == 4) {
                float4 t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (it + u < it1) t[u] = vy_ldg128(plane + (size_t)(5 + c0 + it + u) * HW);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (it + u < it1) {
                        const int c = c0 + it + u;
                        if (t[u].x >= tcmin[0]) hit(S, t[u].x, 0, c, thr);
                        if (t[u].y >= tcmin[1 % VEC]) hit(S, t[u].y, 1 % VEC, c, thr);
                        if (t[u].z >= tcmin[2 % VEC]) hit(S, t[u].z, 2 % VEC, c, thr);
                        if (t[u].w >= tcmin[3 % VEC]) hit(S, t[u].w, 3 % VEC, c, thr);
                    }
                }
            } else {
                float t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (it + u < it1) t[u] = vy_ldg32(plane + (size_t)(5 + c0 + it + u) * HW);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (it + u < it1 && t[u] >= tcmin[0]) hit(S, t[u], 0, c0 + it + u, thr);
            }
        }
    }
};

template <int VEC>
__device__ void head_tile(SelBuf &S, const VyHeads &hd, const VyScale &sc, int b, int a, int chunk,
                          int c0, int c1, int K, float valid_thresh, u64 *g_thr_b) {
    const int tid = threadIdx.x;
    const int pos0 = (chunk * SEL_NT + tid) * VEC;
    HeadSrc<VEC> src;
    src.active = pos0 < sc.HW;
    src.HW = (size_t)sc.HW;
    src.plane = sc.head + ((size_t)(b * hd.A + a) * hd.P) * src.HW + (src.active ? pos0 : 0);
    src.c0 = c0;
    src.n_s = (u32)sc.n_s;
    src.A = (u32)hd.A;
    src.row0 = (u32)(sc.row_off + (long long)pos0 * hd.A + a);
    src.valid_thresh = valid_thresh;
    src.max_push_per_iter = SEL_NT * VEC;
    if (VEC == 4) {
        float4 to = make_float4(0, 0, 0, 0);
        if (src.active) to = vy_ldg128(src.plane + 4 * src.HW);
        src.conf[0] = vy_sigmoid(to.x); src.conf[1 % VEC] = vy_sigmoid(to.y);
        src.conf[2 % VEC] = vy_sigmoid(to.z); src.conf[3 % VEC] = vy_sigmoid(to.w);
    } else {
        src.conf[0] = src.active ? vy_sigmoid(vy_ldg32(src.plane + 4 * src.HW)) : 0.0f;
    }
    if (hd.agnostic) {
        // yolo3.py:184-188: one candidate per box, score = objectness, row = off + pos*A + a
        src.n_iters = 0;
        __syncthreads();
        const int cnt = S.count;
        __syncthreads();
        if (cnt > SEL_CAP - SEL_NT * VEC) sel_compact(S, cnt, K, true);
        const u64 thr = S.thr;
        if (src.active) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float s = src.conf[v];
                if (s > valid_thresh) {
                    const u64 key = vy_make_key(s, src.row0 + (u32)v * src.A);
                    if (key >= thr) sel_push(S, key);
                }
            }
        }
        __syncthreads();
        return;
    }
    src.n_iters = c1 - c0;
    sel_stream(S, src, K, g_thr_b);
}

__global__ void __launch_bounds__(SEL_NT, 4)
vy_decode_select_kernel(VyHeads hd, SelPlan pl, SelGlobal g, int total_tiles) {
    __shared__ SelBuf S;
    int cur_b = -1;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t % hd.B;
        const int j = t / hd.B;
        if (b != cur_b) {
            if (cur_b >= 0)
                sel_flush(S, pl.K, g.thr + cur_b, g.count + cur_b, g.list + (size_t)cur_b * pl.list_cap, pl.list_cap);
            __syncthreads();
            if (threadIdx.x == 0) { S.count = 0; S.thr = ld_relaxed_u64(g.thr + b); }
            cur_b = b;
            __syncthreads();
        }
        int s = 0;
        while (s + 1 < hd.n_scales && j >= pl.tile_begin[s + 1]) ++s;
        int jj = j - pl.tile_begin[s];
        const int cpart = jj % pl.csplit; jj /= pl.csplit;
        const int chunk = jj % pl.chunks[s];
        const int a = jj / pl.chunks[s];
        const int c0 = cpart * pl.cper;
        const int c1 = min(hd.C, c0 + pl.cper);
        if (hd.sc[s].vec == 4)
            head_tile<4>(S, hd, hd.sc[s], b, a, chunk, c0, c1, pl.K, pl.valid_thresh, g.thr + b);
        else
            head_tile<1>(S, hd, hd.sc[s], b, a, chunk, c0, c1, pl.K, pl.valid_thresh, g.thr + b);
    }
    if (cur_b >= 0)
        sel_flush(S, pl.K, g.thr + cur_b, g.count + cur_b, g.list + (size_t)cur_b * pl.list_cap, pl.list_cap);
}

// ------------------------------------------------------------------------------------------------
// source 2: materialised detection rows (generic box_nms).  Iteration = SEL_NT consecutive rows.
// ------------------------------------------------------------------------------------------------
struct RowSrc {
    int n_iters, max_push_per_iter;
    const float *img;            // data + b*R*W
    long long row_begin, row_end;
    int W, score_index, id_index, background_id;
    float valid_thresh, smin;

    __device__ void refresh(u64 thr) {
        smin = thr ? vy_key_score(thr) : -CUDART_INF_F;
    }
    __device__ void run(SelBuf &S, int it0, int it1, u64 thr) {
        for (int it = it0; it < it1; it += 4) {
            float sc[4];
            long long r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                r[u] = row_begin + (long long)(it + u) * SEL_NT + threadIdx.x;
                sc[u] = (it + u < it1 && r[u] < row_end) ? vy_ldg32(img + r[u] * W + score_index) : CUDART_NAN_F;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float s = sc[u];
                if (s > valid_thresh && s >= smin) {          // NaN fails both
                    if (id_index >= 0 && background_id >= 0 &&
                        (int)img[r[u] * W + id_index] == background_id) continue;
                    const u64 key = vy_make_key(s, (u32)r[u]);
                    if (key >= thr) sel_push(S, key);
                }
            }
        }
    }
};

__global__ void __launch_bounds__(SEL_NT, 4)
vy_rows_select_kernel(RowParams rp, int B, SelPlan pl, SelGlobal g, int total_tiles) {
    __shared__ SelBuf S;
    int cur_b = -1;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t % B;
        const long long j = t / B;
        if (b != cur_b) {
            if (cur_b >= 0)
                sel_flush(S, pl.K, g.thr + cur_b, g.count + cur_b, g.list + (size_t)cur_b * pl.list_cap, pl.list_cap);
            __syncthreads();
            if (threadIdx.x == 0) { S.count = 0; S.thr = ld_relaxed_u64(g.thr + b); }
            cur_b = b;
            __syncthreads();
        }
        RowSrc src;
        src.img = rp.data + (size_t)b * (size_t)rp.R * rp.W;
        src.row_begin = j * pl.rows_per_tile;
        src.row_end = min(rp.R, src.row_begin + pl.rows_per_tile);
        src.W = rp.W; src.score_index = rp.score_index; src.id_index = rp.id_index;
        src.background_id = rp.background_id; src.valid_thresh = rp.valid_thresh;
        src.n_iters = (int)((src.row_end - src.row_begin + SEL_NT - 1) / SEL_NT);
        src.max_push_per_iter = SEL_NT;
        sel_stream(S, src, pl.K, g.thr + b);
    }
    if (cur_b >= 0)
        sel_flush(S, pl.K, g.thr + cur_b, g.count + cur_b, g.list + (size_t)cur_b * pl.list_cap, pl.list_cap);
}

// ------------------------------------------------------------------------------------------------
// finalize: one CTA per image
// ------------------------------------------------------------------------------------------------
constexpr int FIN_NT = 512;

struct FinParams {
    int K, post_rows;            // rows written per image
    long long out_stride_rows;   // rows per image in `out` (== post_rows)
    float overlap_thresh;
    int force_suppress, in_format, out_format;
    int W;                       // output row width (6 for heads)
    int fill_rest;               // 1: this kernel writes the -1 padding rows itself
    float *out;
    int *kept_rows;
};

// upstream BoxArea / Intersect (bounding_box-inl.h), same association order as the oracle
__device__ __forceinline__ float nms_area(float4 b, int fmt) {
    float w, h;
    if (fmt == VY_FMT_CORNER) { w = __fsub_rn(b.z, b.x); h = __fsub_rn(b.w, b.y); }
    else { w = b.z; h = b.w; }
    if (w < 0 || h < 0) return 0.0f;
    return __fmul_rn(w, h);
}
__device__ __forceinline__ float nms_isect(float a1, float a2, float b1, float b2, int fmt) {
    float w;
    if (fmt == VY_FMT_CORNER) {
        const float left = a1 > b1 ? a1 : b1;
        const float right = a2 < b2 ? a2 : b2;
        w = __fsub_rn(right, left);
    } else {
        const float aw = __fdiv_rn(a2, 2.0f), bw = __fdiv_rn(b2, 2.0f);
        const float al = __fsub_rn(a1, aw), ar = __fadd_rn(a1, aw);
        const float bl = __fsub_rn(b1, bw), br = __fadd_rn(b1, bw);
        const float left = bl > al ? bl : al;
        const float right = br < ar ? br : ar;
        w = __fsub_rn(right, left);
    }
    return w > 0 ? w : 0.0f;
}

template <int SRC>   // 0: head maps, 1: rows
__global__ void __launch_bounds__(FIN_NT)
vy_nms_finalize_kernel(VyHeads hd, RowParams rp, SelPlan pl, SelGlobal g, FinParams fp) {
    __shared__ SelBuf S;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = FIN_NT / 32;
    const int b = blockIdx.x;
    const int K = pl.K;
    const int nwK = (K + 31) >> 5;
    float4 *box = (float4 *)dyn;                       // K
    float *area = (float *)(box + K);                  // K
    int *cls = (int *)(area + K);                      // K
    u32 *mask = (u32 *)(cls + K);                      // K * nwK
    u32 *rowany = mask + (size_t)K * nwK;              // 32
    u32 *keepw = rowany + 32;                          // 32
    int *kprefix = (int *)(keepw + 32);                // 33

    // ---- 1. exact top-K of the image's candidate list, sorted descending
    if (tid == 0) { S.count = 0; S.thr = g.thr[b]; }
    __syncthreads();
    const int n_list = min(g.count[b], pl.list_cap);
    const u64 *list = g.list + (size_t)b * pl.list_cap;
    int cnt = 0;
    for (int off = 0; off < n_list;) {
        const int take = min(SEL_CAP - cnt, n_list - off);          // CTA-uniform
        const u64 thr = S.thr;
        for (int i = tid; i < take; i += FIN_NT) {
            const u64 key = list[off + i];
            if (key >= thr) sel_push(S, key);
        }
        __syncthreads();
        cnt = S.count;
        __syncthreads();
        off += take;
        if (off < n_list && cnt > K) cnt = sel_compact(S, cnt, K, false);
    }
    const int m = sel_compact(S, cnt, K, true);         // <= K candidates take part
    int npow2 = 32;
    while (npow2 < m) npow2 <<= 1;
    for (int i = m + tid; i < npow2; i += FIN_NT) S.keys[i] = 0ull;
    if (tid < 32) { rowany[tid] = 0; keepw[tid] = 0; }
    __syncthreads();
    sel_sort_desc(S, npow2);

    // ---- 2. boxes / classes of the m candidates
    const int nw = (m + 31) >> 5;
    for (int i = tid; i < m; i += FIN_NT) {
        const u32 row = vy_key_row(S.keys[i]);
        float4 bx; int c;
        if (SRC == 0) {
            int s = 0;
            while (s + 1 < hd.n_scales && (long long)row >= hd.sc[s + 1].row_off) ++s;
            const VyScale &sc = hd.sc[s];
            const u32 rel = row - (u32)sc.row_off;
            c = (int)(rel / (u32)sc.n_s);
            const u32 rem = rel % (u32)sc.n_s;
            const int pos = (int)(rem / (u32)hd.A), a = (int)(rem % (u32)hd.A);
            const int y = pos / sc.W, x = pos % sc.W;
            const float *p = sc.head + ((size_t)(b * hd.A + a) * hd.P) * (size_t)sc.HW + pos;
            bx = vy_box(p[0], p[sc.HW], p[2 * (size_t)sc.HW], p[3 * (size_t)sc.HW], x, y,
                        sc.stride, sc.aw[a], sc.ah[a]);
            if (hd.agnostic) c = 0;
        } else {
            const float *p = rp.data + ((size_t)b * (size_t)rp.R + row) * rp.W;
            bx = make_float4(p[rp.coord_start], p[rp.coord_start + 1], p[rp.coord_start + 2], p[rp.coord_start + 3]);
            c = rp.id_index >= 0 ? (int)p[rp.id_index] : 0;
        }
        box[i] = bx; cls[i] = c; area[i] = nms_area(bx, fp.in_format);
    }
    __syncthreads();

    // ---- 3. suppression bitmask, one warp per (ref row i, 32-candidate word w >= i/32)
    const bool all_pairs = fp.force_suppress || (SRC == 1 && rp.id_index < 0);
    for (int i = warp; i < m; i += nwarps) {
        const float4 bi = box[i];
        const float ai = area[i];
        const int ci = cls[i];
        u32 any = 0;
        for (int w = i >> 5; w < nw; ++w) {
            const int jx = (w << 5) + lane;
            bool sup = false;
            if (jx > i && jx < m && (all_pairs || cls[jx] == ci)) {
                const float4 bj = box[jx];
                float inter = nms_isect(bi.x, bi.z, bj.x, bj.z, fp.in_format);
                inter = __fmul_rn(inter, nms_isect(bi.y, bi.w, bj.y, bj.w, fp.in_format));
                const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, area[jx]), inter));
                sup = iou > fp.overlap_thresh;
            }
            const u32 bits = __ballot_sync(0xffffffffu, sup);
            if (lane == 0) mask[(size_t)i * nwK + w] = bits;
            any |= bits;
        }
        if (lane == 0 && any) atomicOr(&rowany[i >> 5], 1u << (i & 31));
    }
    __syncthreads();

    // ---- 4. greedy scan in score order (warp 0).  lane w holds the suppressed-bits of word w.
    if (warp == 0) {
        u32 removed = 0;
        for (int blk = 0; blk < nw; ++blk) {
            const int r = (blk << 5) + lane;
            const u32 validm = (m - (blk << 5) >= 32) ? 0xffffffffu : ((1u << (m - (blk << 5))) - 1u);
            const u32 ra = rowany[blk] & validm;
            const u32 diag = (r < m && ((ra >> lane) & 1u)) ? mask[(size_t)r * nwK + blk] : 0u;
            u32 rem = __shfl_sync(0xffffffffu, removed, blk);
            // refs of this block whose row is non-empty, resolved in order
            u32 pend = ra;
            while (pend) {
                const int i = __ffs(pend) - 1;
                pend &= pend - 1;
                const u32 di = __shfl_sync(0xffffffffu, diag, i);
                if (!((rem >> i) & 1u)) rem |= di;
            }
            const u32 keep = ~rem & validm;
            if (lane == 0) keepw[blk] = keep;
            // propagate surviving refs with non-empty rows to the later words
            u32 act = keep & ra;
            if (lane > blk && lane < nw) {
                u32 acc = 0;
                while (act) {
                    const int i = __ffs(act) - 1;
                    act &= act - 1;
                    acc |= mask[(size_t)((blk << 5) + i) * nwK + lane];
                }
                removed |= acc;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int w = 0; w < nw; ++w) { kprefix[w] = run; run += __popc(keepw[w]); }
        kprefix[32] = run;
    }
    __syncthreads();
    const int n_keep = kprefix[32];

    // ---- 5. survivors to the front in score order; the rest is -1
    const int W = fp.W;
    float *out_b = fp.out + (size_t)b * (size_t)fp.out_stride_rows * W;
    int *kept_b = fp.kept_rows ? fp.kept_rows + (size_t)b * (size_t)fp.out_stride_rows : nullptr;
    for (int i = tid; i < m; i += FIN_NT) {
        const u32 kw = keepw[i >> 5];
        if (!((kw >> (i & 31)) & 1u)) continue;
        const int p = kprefix[i >> 5] + __popc(kw & ((1u << (i & 31)) - 1u));
        if (p >= fp.post_rows) continue;
        const u64 key = S.keys[i];
        const u32 row = vy_key_row(key);
        float *o = out_b + (size_t)p * W;
        if (SRC == 0) {
            const float4 bx = box[i];
            o[0] = (float)cls[i]; o[1] = vy_key_score(key);
            o[2] = bx.x; o[3] = bx.y; o[4] = bx.z; o[5] = bx.w;
        } else {
            const float *src = rp.data + ((size_t)b * (size_t)rp.R + row) * rp.W;
            for (int c = 0; c < W; ++c) o[c] = src[c];
            if (fp.in_format != fp.out_format) {
                float *q = o + rp.coord_start;
                if (!(q[0] < 0)) {
                    if (fp.out_format == VY_FMT_CENTER) {   // corner_to_center
                        const float l = q[0], t = q[1], r2 = q[2], bt = q[3];
                        q[0] = __fdiv_rn(__fadd_rn(l, r2), 2.0f); q[1] = __fdiv_rn(__fadd_rn(t, bt), 2.0f);
                        q[2] = __fsub_rn(r2, l); q[3] = __fsub_rn(bt, t);
                    } else {                                 // center_to_corner
                        const float x = q[0], y = q[1];
                        const float hw = __fdiv_rn(q[2], 2.0f), hh = __fdiv_rn(q[3], 2.0f);
                        q[0] = __fsub_rn(x, hw); q[1] = __fsub_rn(y, hh);
                        q[2] = __fadd_rn(x, hw); q[3] = __fadd_rn(y, hh);
                    }
                }
            }
        }
        if (kept_b) kept_b[p] = (int)row;
    }
    if (fp.fill_rest) {
        const int first = min(n_keep, fp.post_rows);
        const long long total = (long long)(fp.post_rows - first) * W;
        float *o = out_b + (size_t)first * W;
        for (long long i = tid; i < total; i += FIN_NT) o[i] = -1.0f;
        if (kept_b) for (int i = first + tid; i < fp.post_rows; i += FIN_NT) kept_b[i] = -1;
    }
}

__global__ void vy_fill_kernel(float *out, int *kept, size_t n_out, size_t n_kept) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) out[i] = -1.0f;
    if (kept) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_kept; i += stride) kept[i] = -1;
}

// ------------------------------------------------------------------------------------------------
// host: planning + launches
// ------------------------------------------------------------------------------------------------
static size_t fin_dyn_smem(int K) {
    const int nwK = (K + 31) / 32;
    return (size_t)K * (16 + 4 + 4) + (size_t)K * nwK * 4 + 32 * 4 + 32 * 4 + 33 * 4 + 16;
}

static int plan_heads(const VyHeads &hd, int topk, float valid_thresh, SelPlan *pl) {
    long long K = topk < 0 ? hd.R : (topk < hd.R ? topk : hd.R);
    if (K < 1 || K > SEL_KMAX) return VY_EUNSUPPORTED;
    pl->K = (int)K;
    pl->valid_thresh = valid_thresh;
    int base_tiles = 0;
    for (int s = 0; s < hd.n_scales; ++s) {
        const int per = SEL_NT * hd.sc[s].vec;
        pl->chunks[s] = (hd.sc[s].HW + per - 1) / per;
        base_tiles += pl->chunks[s] * hd.A;
    }
    // split the class range while the grid is too small to fill the machine, keeping >= 4
    // class planes per tile so the per-box objectness work stays amortised
    int csplit = 1;
    if (!hd.agnostic) {
        const long long want = 4096;
        const long long have = (long long)base_tiles * hd.B;
        csplit = (int)((want + have - 1) / have);
        const int max_split = hd.C / 4 > 1 ? hd.C / 4 : 1;
        if (csplit > max_split) csplit = max_split;
        if (csplit < 1) csplit = 1;
    }
    pl->cper = (hd.C + csplit - 1) / csplit;
    pl->csplit = (hd.C + pl->cper - 1) / pl->cper;
    int tb = 0;
    for (int s = 0; s < hd.n_scales; ++s) {
        pl->tile_begin[s] = tb;
        tb += pl->chunks[s] * hd.A * pl->csplit;
    }
    for (int s = hd.n_scales; s <= VY_MAX_SCALES; ++s) pl->tile_begin[s] = tb;
    pl->tiles_per_frame = tb;
    pl->rows_per_tile = 0;
    const long long cap = (long long)tb * (pl->K + (pl->K >> 2));
    pl->list_cap = (int)(cap < 64 ? 64 : cap);
    return VY_OK;
}

static int plan_rows(int B, long long R, int topk, float valid_thresh, SelPlan *pl) {
    long long K = topk < 0 ? R : (topk < R ? topk : R);
    if (K < 1 || K > SEL_KMAX) return VY_EUNSUPPORTED;
    memset(pl, 0, sizeof(*pl));
    pl->K = (int)K;
    pl->valid_thresh = valid_thresh;
    // ~4096 tiles in flight, each a multiple of SEL_NT rows and at least 16 iterations long
    long long tiles = (4096 + B - 1) / B;
    long long rpt = (R + tiles - 1) / tiles;
    if (rpt < 16 * SEL_NT) rpt = 16 * SEL_NT;
    rpt = (rpt + SEL_NT - 1) / SEL_NT * SEL_NT;
    pl->rows_per_tile = rpt;
    pl->tiles_per_frame = (int)((R + rpt - 1) / rpt);
    const long long cap = (long long)pl->tiles_per_frame * (pl->K + (pl->K >> 2));
    pl->list_cap = (int)(cap < 64 ? 64 : cap);
    pl->csplit = 1; pl->cper = 1;
    return VY_OK;
}

static int select_grid(const void *kernel, int total_tiles) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SEL_NT, 0) != cudaSuccess || per_sm < 1)
        per_sm = 4;
    const long long resident = (long long)per_sm * vy_sm_count();
    return (int)(total_tiles < resident ? total_tiles : resident);
}

template <int SRC>
static int launch_finalize(const VyHeads &hd, const RowParams &rp, const SelPlan &pl, const SelGlobal &g,
                           FinParams fp, int B, cudaStream_t st) {
    const size_t dyn = fin_dyn_smem(pl.K);
    VY_CUDA_CHECK(cudaFuncSetAttribute(vy_nms_finalize_kernel<SRC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    vy_nms_finalize_kernel<SRC><<<B, FIN_NT, dyn, st>>>(hd, rp, pl, g, fp);
    VY_LAUNCH_CHECK("vy_nms_finalize_kernel");
    return VY_OK;
}

extern "C" size_t vy_decode_nms_workspace_bytes(const int *H, const int *W, int n_scales, int B, int A,
                                                int C, int agnostic, int topk) {
    VyHeads hd;
    const float *fake[VY_MAX_SCALES] = {nullptr, nullptr, nullptr, nullptr};
    float st[VY_MAX_SCALES] = {1, 1, 1, 1};
    float an[VY_MAX_SCALES * VY_MAX_ANCHORS * 2] = {0};
    if (vy_fill_heads(&hd, fake, H, W, st, an, n_scales, B, A, C, agnostic) != VY_OK) return 0;
    SelPlan pl;
    if (plan_heads(hd, topk, 0.0f, &pl) != VY_OK) { vy_set_error("topk out of range for the fused path"); return 0; }
    return sel_workspace_layout(B, pl.list_cap, nullptr, nullptr);
}

extern "C" int vy_decode_nms_f32(const float *const *head, const int *H, const int *W, const float *stride,
                                 const float *anchors, int n_scales, int B, int A, int C, int agnostic,
                                 float overlap_thresh, float valid_thresh, int topk, int force_suppress,
                                 int post_nms, float *out, int32_t *kept_rows, void *workspace,
                                 size_t workspace_bytes, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    VyHeads hd;
    int rc = vy_fill_heads(&hd, head, H, W, stride, anchors, n_scales, B, A, C, agnostic);
    if (rc != VY_OK) return rc;
    if (post_nms < 1 || !out) VY_FAIL(VY_EINVAL, "vy_decode_nms_f32: post_nms must be >= 1 and out non-null");
    SelPlan pl;
    rc = plan_heads(hd, topk, valid_thresh, &pl);
    if (rc != VY_OK) VY_FAIL(rc, "vy_decode_nms_f32: min(topk,R)=%d outside [1,%d]; use vy_decode_f32 + vy_box_nms_f32",
                             topk, SEL_KMAX);
    SelGlobal g;
    const size_t need = sel_workspace_layout(B, pl.list_cap, &g, workspace);
    if (!workspace || workspace_bytes < need)
        VY_FAIL(VY_EWORKSPACE, "vy_decode_nms_f32: workspace %zu < %zu bytes", workspace_bytes, need);
    if (((uintptr_t)workspace & 255) != 0) VY_FAIL(VY_EALIGN, "workspace must be 256-byte aligned");
    VY_CUDA_CHECK(cudaMemsetAsync(workspace, 0, sel_header_bytes(B), st));
    const long long total = (long long)pl.tiles_per_frame * B;
    if (total > 0x7fffffffLL) VY_FAIL(VY_EINVAL, "too many tiles");
    const int grid = select_grid((const void *)vy_decode_select_kernel, (int)total);
    vy_decode_select_kernel<<<grid, SEL_NT, 0, st>>>(hd, pl, g, (int)total);
    VY_LAUNCH_CHECK("vy_decode_select_kernel");
    FinParams fp;
    fp.K = pl.K; fp.post_rows = post_nms; fp.out_stride_rows = post_nms;
    fp.overlap_thresh = overlap_thresh; fp.force_suppress = force_suppress;
    fp.in_format = VY_FMT_CORNER; fp.out_format = VY_FMT_CORNER; fp.W = 6; fp.fill_rest = 1;
    fp.out = out; fp.kept_rows = kept_rows;
    RowParams rp;
    memset(&rp, 0, sizeof(rp));
    return launch_finalize<0>(hd, rp, pl, g, fp, B, st);
}

// implemented in vy_nms_large.cu: topk < 0 or min(topk,R) > SEL_KMAX
size_t vy_box_nms_large_workspace_bytes(int B, long long R, int W_elem);
int vy_box_nms_large(const RowParams &rp, int B, long long K, float overlap_thresh, int force_suppress,
                     int in_format, int out_format, long long out_rows, float *out, int32_t *kept_rows,
                     void *workspace, size_t workspace_bytes, cudaStream_t st);

extern "C" size_t vy_box_nms_workspace_bytes(int B, long R, int W_elem, int topk) {
    SelPlan pl;
    if (B < 1 || R < 1) return 0;
    if (plan_rows(B, R, topk, 0.0f, &pl) == VY_OK) return sel_workspace_layout(B, pl.list_cap, nullptr, nullptr);
    return vy_box_nms_large_workspace_bytes(B, R, W_elem);
}

extern "C" int vy_box_nms_f32(const float *data, int B, long R, int W_elem, float overlap_thresh,
                              float valid_thresh, int topk, int coord_start, int score_index, int id_index,
                              int background_id, int force_suppress, int in_format, int out_format,
                              long out_rows, float *out, int32_t *kept_rows, void *workspace,
                              size_t workspace_bytes, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!data || !out || B < 1 || R < 1 || W_elem < 1) VY_FAIL(VY_EINVAL, "vy_box_nms_f32: bad data/out/B/R/W");
    if (R > 0xfffffffeLL) VY_FAIL(VY_EINVAL, "vy_box_nms_f32: R too large");
    if (coord_start < 0 || coord_start + 4 > W_elem || score_index < 0 || score_index >= W_elem ||
        id_index >= W_elem)
        VY_FAIL(VY_EINVAL, "vy_box_nms_f32: coord_start/score_index/id_index outside the row (W=%d)", W_elem);
    if ((in_format != VY_FMT_CORNER && in_format != VY_FMT_CENTER) ||
        (out_format != VY_FMT_CORNER && out_format != VY_FMT_CENTER))
        VY_FAIL(VY_EINVAL, "vy_box_nms_f32: bad format");
    if (out_rows < 1 || out_rows > R) VY_FAIL(VY_EINVAL, "vy_box_nms_f32: out_rows must be in [1, R]");
    RowParams rp;
    rp.data = data; rp.R = R; rp.W = W_elem; rp.coord_start = coord_start; rp.score_index = score_index;
    rp.id_index = id_index; rp.background_id = background_id; rp.valid_thresh = valid_thresh;
    SelPlan pl;
    if (plan_rows(B, R, topk, valid_thresh, &pl) != VY_OK) {
        const long long K = topk < 0 ? R : (topk < R ? topk : R);
        if (K < 1) {      // topk == 0: nothing takes part
            vy_fill_kernel<<<vy_sm_count() * 4, 256, 0, st>>>(out, kept_rows, (size_t)B * out_rows * W_elem,
                                                               (size_t)B * out_rows);
            VY_LAUNCH_CHECK("vy_fill_kernel");
            return VY_OK;
        }
        return vy_box_nms_large(rp, B, K, overlap_thresh, force_suppress, in_format, out_format, out_rows,
                                out, kept_rows, workspace, workspace_bytes, st);
    }
    SelGlobal g;
    const size_t need = sel_workspace_layout(B, pl.list_cap, &g, workspace);
    if (!workspace || workspace_bytes < need)
        VY_FAIL(VY_EWORKSPACE, "vy_box_nms_f32: workspace %zu < %zu bytes", workspace_bytes, need);
    if (((uintptr_t)workspace & 255) != 0) VY_FAIL(VY_EALIGN, "workspace must be 256-byte aligned");
    VY_CUDA_CHECK(cudaMemsetAsync(workspace, 0, sel_header_bytes(B), st));
    const long long total = (long long)pl.tiles_per_frame * B;
    if (total > 0x7fffffffLL) VY_FAIL(VY_EINVAL, "too many tiles");
    const int grid = select_grid((const void *)vy_rows_select_kernel, (int)total);
    vy_rows_select_kernel<<<grid, SEL_NT, 0, st>>>(rp, B, pl, g, (int)total);
    VY_LAUNCH_CHECK("vy_rows_select_kernel");
    FinParams fp;
    fp.K = pl.K; fp.post_rows = (int)(out_rows < pl.K ? out_rows : pl.K); fp.out_stride_rows = out_rows;
    fp.overlap_thresh = overlap_thresh; fp.force_suppress = force_suppress;
    fp.in_format = in_format; fp.out_format = out_format; fp.W = W_elem;
    fp.out = out; fp.kept_rows = kept_rows;
    if (out_rows > pl.K) {
        // at most K rows can survive: pad everything first, survivors overwrite the front
        vy_fill_kernel<<<vy_sm_count() * 8, 256, 0, st>>>(out, kept_rows, (size_t)B * out_rows * W_elem,
                                                           (size_t)B * out_rows);
        VY_LAUNCH_CHECK("vy_fill_kernel");
        fp.fill_rest = 0;
    } else {
        fp.fill_rest = 1;
    }
    VyHeads hd;
    memset(&hd, 0, sizeof(hd));
    return launch_finalize<1>(hd, rp, pl, g, fp, B, st);
}
