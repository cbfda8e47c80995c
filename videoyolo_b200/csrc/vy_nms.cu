// vy_nms.cu -- candidate selection + box_nms kernels (fused-from-heads and generic-from-rows).
//
// Fused decode + box_nms from the head maps, large class-aware inputs (R >= 131072): four launches chained with
// programmatic dependent launch, no memset, no host sync --
//   1. vy_decode_sample_kernel  : 1/S of every image -> a per-image score bound (an ESTIMATE of the rank-4K score)
//   2. vy_decode_stream_kernel  : the one pass over the class planes at that fixed bound; the few survivors are queued,
//                                 scored 32 at a time and appended to a per-image candidate list
//   3. vy_decode_select_kernel  : rescue pass for images whose list is unusable (normally exits at once)
//   4. vy_nms_finalize_kernel   : one CTA per image: exact top-K of the list (bucket sort), regroup by class, decode of
//                                 the K boxes, suppression (lane per reference slot, or 32 x 32 tiles for long
//                                 segments), greedy resolution per segment, compaction, (post_nms, W) rows + kept rows
// Small / class-agnostic inputs and materialised rows (vy_box_nms_f32, topk <= 1024): a memset, one adaptive
// streaming select (*_select_kernel: per-CTA top-K under a rising threshold shared per image) and the same finalize.
// Semantics follow MXNet _contrib_box_nms as called at yolo3.py:525-530.
#include "vy_select.cuh"
#include "vy_nms_math.cuh"
#include <math_constants.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

// ------------------------------------------------------------------------------------------------
// job plan shared by workspace sizing and launch
//   An image is split into G equal shares; job (b, g) streams share g of image b from start to end
//   in ONE CTA, so its candidate buffer and threshold persist for the whole share.
// ------------------------------------------------------------------------------------------------
#ifndef SEL_CTAS_PER_SM
#define SEL_CTAS_PER_SM 4
#endif

struct SelPlan {
    int K;                      // min(topk, R)
    int G, Kq;                  // CTAs per image, ceil(K / G)
    int n_jobs;                 // B * G
    // heads: an "item" is 4 consecutive positions of one (scale, anchor) plane set
    int items_per_plane[VY_MAX_SCALES];     // ceil(HW / 4)
    int item_begin[VY_MAX_SCALES + 1];      // first item of each scale (A * items_per_plane each)
    int items_per_frame;
    long long rows_per_job;     // rows: multiple of SEL_NT
    int list_cap;               // keys per image in the global list
    float valid_thresh;
    // ---- streaming path (heads only; stream == 0: not used)
    int stream;
    int samp_stride;            // every samp_stride-th PAIR of items is sampled
    int samp_items;             // sampled items per image (even)
    int Gs, Ksq;                // sample CTAs per image (= samp_ib * plane blocks), ceil(K / Gs)
    int samp_ib, samp_ipj;      // item blocks per image, sampled items per job (<= SAMP_NT)
    int samp_ppj, samp_pls;     // class planes per job, plane lanes per item (threads sharing an item)
    int n_groups, PU;           // class planes are cut into n_groups groups of PU planes
    int chunks[VY_MAX_SCALES];  // 128-position chunks per plane
    int unit_begin[VY_MAX_SCALES + 1];
    int units_per_image;
    long long n_units;
    // ---- tile streaming (stream2): every (b, scale, anchor) block of class planes is one contiguous array of C*HW
    // floats, cut into tiles of S2_TILE bytes; the global tile sequence (b, s, a, t) is dealt to the CTAs in equal
    // contiguous ranges.  The per-position logit bounds come from a table built by vy_decode_table_kernel.
    int tiles_blk[VY_MAX_SCALES];           // tiles per (b, s, a) block
    int tile_tpp[VY_MAX_SCALES];            // planes of >= S2_TILE bytes: tiles per plane (else 0)
    int tile_ppt[VY_MAX_SCALES];            // smaller planes: whole planes per tile (else 0)
    int tile_begin[VY_MAX_SCALES + 1];      // first tile of scale s within an image (order: s, a, t)
    int tiles_per_image;
    long long n_tiles;
    int tab_hwp[VY_MAX_SCALES];             // floats per (s, a) table segment: HW + 3 wrap-around copies, rounded up to 4
    int tab_off[VY_MAX_SCALES + 1];         // first float of scale s in an image's table
    int tab_max;                            // largest segment (floats)
    int s3_groups[VY_MAX_SCALES];           // segment streaming: warp groups per CTA (each with its own table) at scale s
};

struct SelGlobal {              // workspace views
    u64 *thr;                   // [B]        best bound per image
    int *count;                 // [B]        list fill
    u64 *slots;                 // [B][G]     per-CTA ceil(K/G)-th largest
    u64 *list;                  // [B][list_cap]
    // streaming path only (see "sample + stream" below); null otherwise
    int *scount;                // [B]        fill of the streamed candidate list
    int *sdone;                 // [B]        sample jobs finished
    u64 *sslots;                // [B][Gs]    COMPLEMENT of every sample job's bound (each job stores its own: nothing to zero)
    int Gs;
    u64 *slist;                 // [B][slist_cap]
    int slist_cap;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Gs == 0: no streaming-path arrays
static size_t sel_workspace_layout(int B, int G, int list_cap, SelGlobal *g, void *base, size_t *header,
                                   int Gs = 0, int slist_cap = 0) {
    size_t off = 0;
    const size_t o_thr = off;   off = align_up(off + sizeof(u64) * (size_t)B, 256);
    const size_t o_cnt = off;   off = align_up(off + sizeof(int) * (size_t)B, 256);
    const size_t o_slot = off;  off = align_up(off + sizeof(u64) * (size_t)B * G, 256);
    const size_t o_scnt = off;  off = align_up(off + sizeof(int) * (size_t)B * (Gs ? 1 : 0), 256);
    const size_t o_sdone = off; off = align_up(off + sizeof(int) * (size_t)B * (Gs ? 1 : 0), 256);
    const size_t o_sslot = off; off = align_up(off + sizeof(u64) * (size_t)B * Gs, 256);
    if (header) *header = off;  // the part that must be zeroed per call
    const size_t o_list = off;  off = align_up(off + sizeof(u64) * (size_t)B * (size_t)list_cap, 256);
    const size_t o_slist = off; off = align_up(off + sizeof(u64) * (size_t)B * (size_t)slist_cap, 256);
    if (g && base) {
        g->thr = (u64 *)((char *)base + o_thr);
        g->count = (int *)((char *)base + o_cnt);
        g->slots = (u64 *)((char *)base + o_slot);
        g->list = (u64 *)((char *)base + o_list);
        g->Gs = Gs;
        g->scount = Gs ? (int *)((char *)base + o_scnt) : nullptr;
        g->sdone = Gs ? (int *)((char *)base + o_sdone) : nullptr;
        g->sslots = Gs ? (u64 *)((char *)base + o_sslot) : nullptr;
        g->slist = Gs ? (u64 *)((char *)base + o_slist) : nullptr;
        g->slist_cap = slist_cap;
    }
    return off;
}

// Is the streamed candidate list of image b usable?  It is not when it overflowed, or when it holds
// fewer than K keys although a bound above valid_thresh was in force (the sample's bound is an
// ESTIMATE of a rank a few times K, not a guaranteed lower bound of the K-th largest score: with
// probability ~1e-7 per image, or for adversarial layouts, it can be too high).  Those images are
// redone exactly by vy_decode_select_kernel.
// COMPLEMENT of the bound the streaming pass works with for image b = the minimum over the image's sample jobs (may be
// optimistic, see below); ~0 = no bound
__device__ __forceinline__ u64 stream_bound_compl(const SelGlobal &g, int b) {
    u64 m = 0;
    for (int j = 0; j < g.Gs; ++j) { const u64 v = g.sslots[(size_t)b * g.Gs + j]; m = v > m ? v : m; }
    return m;
}
__device__ __forceinline__ bool stream_list_ok(const SelGlobal &g, int b, int K) {
    const int n = g.scount[b];
    return n <= g.slist_cap && (n >= K || ~stream_bound_compl(g, b) == 0ull);
}

// ------------------------------------------------------------------------------------------------
// device: flush a finished job to the image's global list
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ void sel_flush(SelBuf &S, const SelJob &jb, int *g_count_b, u64 *g_list_b, int list_cap) {
    const int tid = threadIdx.x, lane = tid & 31;
    __syncthreads();
    int n = S.count;
    __syncthreads();
    n = sel_update(S, n, jb, false);
    if (n > jb.K + (jb.K >> 2)) n = sel_compact(S, n, jb.K, false);
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
// device: generic adaptive block loop over one "round" of a source
//   Src provides: int n_iters;
//                 void refresh(SelBuf&, u64 thr);                  recompute prefilter state
//                 void run(SelBuf&, int it0, int it1);             stream [it0,it1): queue prefilter hits
//                 void eval(SelBuf&, int q0, int q1, u64 thr);     exact keys of queued hits -> sel_push
// A block of U iterations is streamed between CTA barriers; U adapts to the observed hit rate so
// that the hit queue cannot overflow in the steady state.  If it still does (cold start, adversarial
// order) the block is replayed with a smaller U; U == 1 always fits (<= 4*SEL_NT hits).  Queued hits
// are evaluated in chunks that fit the key buffer, which is compacted exactly when it is full, so
// every chunk makes progress (K <= SEL_KMAX leaves >= SEL_CAP - SEL_KMAX free slots).
// `cnt` (CTA-uniform running key count) and `fresh` (keys pushed since the last selection event)
// persist across rounds of one job.
// ------------------------------------------------------------------------------------------------
template <class Src>
static __device__ __forceinline__ void sel_stream(SelBuf &S, Src &src, const SelJob &jb, int &cnt, int &fresh) {
    constexpr int UMAX = 128;
    const int tid = threadIdx.x;
    const int trigger = max(16, jb.Kq >> 1);
    int it = 0;
    int U = (S.thr == 0ull) ? 1 : 8;
    u64 thr_seen = ~0ull;
    while (it < src.n_iters) {                       // CTA-uniform loop
        const u64 thr = S.thr;
        u64 gthr = 0;
        if (tid == 0) gthr = ld_relaxed_u64(jb.g_thr_b);       // consumed after the block
        if (thr != thr_seen) { src.refresh(S, thr); thr_seen = thr; }
        const int Ub = min(U, src.n_iters - it);
        src.run(S, it, it + Ub);
        __syncthreads();                              // hits of this block are queued
        const int nq = S.qcount;                      // not modified before the barriers below
        if (nq > SEL_QCAP) {                          // queue overflow: replay a smaller block
            __syncthreads();
            if (tid == 0) S.qcount = 0;
            __syncthreads();
            U = max(1, Ub >> 1);
            continue;
        }
        int q0 = 0;
        for (;;) {                                    // chunks of queued hits that fit the key buffer
            const int want = nq - q0;
            if (want > SEL_CAP - cnt) { cnt = sel_update(S, cnt, jb, true); fresh = 0; }
            const int take = min(want, SEL_CAP - cnt);
            if (take > 0) src.eval(S, q0, q0 + take, S.thr);
            q0 += take;
            __syncthreads();                          // pushes complete
            if (tid == 0) {
                S.snap = S.count;
                if (q0 >= nq) { S.qcount = 0; if (gthr > S.thr) S.thr = gthr; }
            }
            __syncthreads();
            const int now = S.snap;
            fresh += now - cnt;
            cnt = now;
            if (q0 >= nq) break;
        }
        it += Ub;
        if (cnt >= SEL_CAP / 2 || fresh >= trigger) {
            cnt = sel_update(S, cnt, jb, false);
            if (cnt >= SEL_CAP / 2) cnt = sel_compact(S, cnt, jb.K, true);
            fresh = 0;
        }
        const int rate = (nq + Ub - 1) / Ub;          // hits per iteration
        U = min(UMAX, max(1, SEL_QCAP / (2 * rate + 1)));
    }
}

// ------------------------------------------------------------------------------------------------
// source 1: YOLO head maps (fused decode).  Each thread owns one item = 4 consecutive positions of
// one (scale, anchor) and walks the class planes with one 128-bit load per plane (four 32-bit loads
// when the plane size is not a multiple of 4).  score = sigmoid(t_c)*conf >= smin is tested in the
// logit domain against a per-box bound: one FSETP per element.
// ------------------------------------------------------------------------------------------------
// conservative logit bound: score(t) >= smin  ==>  t >= vy_tcmin(smin, conf).  The score that is
// compared later is fl(S(t)*conf) with S within ~2e-6 of the true sigmoid; q is lowered by 1e-3
// relative and the logit by 2e-3 absolute, which covers those roundings and the error of the
// fast intrinsics used here (__fdividef 2 ulp, __logf <= 1e-6 abs on [0.5,2], 3 ulp elsewhere).
__device__ __forceinline__ float vy_tcmin(float smin, float conf) {
    if (!(smin > 0.0f)) return -CUDART_INF_F;
    const float q = __fdividef(smin, conf) * (1.0f - 1e-3f);
    if (!(q < 1.0f)) return CUDART_INF_F;            // also conf == 0 / NaN: no class can pass
    return __logf(__fdividef(q, 1.0f - q)) - 2e-3f;
}

// any of 16 ordered compares t[u][v] >= c[v]: one FSETP per element, chained through the predicate
__device__ __forceinline__ u32 vy_any_ge16(const float (&t)[4][4], const float (&c)[4]) {
    u32 r;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ge.f32 p, %1, %17;\n\t"
        "setp.ge.or.f32 p, %2, %18, p;\n\t"
        "setp.ge.or.f32 p, %3, %19, p;\n\t"
        "setp.ge.or.f32 p, %4, %20, p;\n\t"
        "setp.ge.or.f32 p, %5, %17, p;\n\t"
        "setp.ge.or.f32 p, %6, %18, p;\n\t"
        "setp.ge.or.f32 p, %7, %19, p;\n\t"
        "setp.ge.or.f32 p, %8, %20, p;\n\t"
        "setp.ge.or.f32 p, %9, %17, p;\n\t"
        "setp.ge.or.f32 p, %10, %18, p;\n\t"
        "setp.ge.or.f32 p, %11, %19, p;\n\t"
        "setp.ge.or.f32 p, %12, %20, p;\n\t"
        "setp.ge.or.f32 p, %13, %17, p;\n\t"
        "setp.ge.or.f32 p, %14, %18, p;\n\t"
        "setp.ge.or.f32 p, %15, %19, p;\n\t"
        "setp.ge.or.f32 p, %16, %20, p;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "f"(t[0][0]), "f"(t[0][1]), "f"(t[0][2]), "f"(t[0][3]),
          "f"(t[1][0]), "f"(t[1][1]), "f"(t[1][2]), "f"(t[1][3]),
          "f"(t[2][0]), "f"(t[2][1]), "f"(t[2][2]), "f"(t[2][3]),
          "f"(t[3][0]), "f"(t[3][1]), "f"(t[3][2]), "f"(t[3][3]),
          "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
    return r;
}

struct HeadSrc {
    int n_iters;                 // class planes
    const VyHeads *hd;
    const float *p5;             // -> class plane 0 (channel a*P + 5) at this thread's first position
    size_t HW;
    int nv;                      // valid positions of this item (0: idle thread)
    int o1, o2, o3;              // element offsets of positions 1..3 (clamped into the plane)
    bool vec;                    // 128-bit loads allowed
    float tcmin[4];
    float valid_thresh;

    __device__ __forceinline__ void refresh(SelBuf &S, u64 thr) {
        const float ts = thr ? vy_key_score(thr) : valid_thresh;
        const float smin = fmaxf(ts, valid_thresh);
#pragma unroll
        for (int v = 0; v < 4; ++v)
            tcmin[v] = v < nv ? vy_tcmin(smin, S.it_conf[v][threadIdx.x]) : CUDART_INF_F;
    }
    __device__ __forceinline__ void load(float (&t)[4], const float *p) const {
        if (vec) {
            const float4 q = vy_ldg128_ca(p);
            t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
        } else {      // positions past the end of the plane re-read the last valid one (tcmin = +inf there)
            t[0] = vy_ldg32_ca(p); t[1] = vy_ldg32_ca(p + o1); t[2] = vy_ldg32_ca(p + o2); t[3] = vy_ldg32_ca(p + o3);
        }
    }
    // planes [c, c+n) of this item into t (n in 1..4); the rest can never hit (tcmin may be -inf)
    __device__ __forceinline__ void load_group(float (&t)[4][4], const float *p, int n) const {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (u < n) load(t[u], p + (size_t)u * HW);
            else t[u][0] = t[u][1] = t[u][2] = t[u][3] = CUDART_NAN_F;   // NaN >= x is false for every x
        }
    }
    __device__ __forceinline__ void run(SelBuf &S, int it0, int it1) {
        if (nv == 0) return;
        float t[4][4];
        const float *p = p5 + (size_t)it0 * HW;
        load_group(t, p, min(4, it1 - it0));
        for (int it = it0; it < it1; it += 4) {
            u32 mask = 0;
            if (vy_any_ge16(t, tcmin)) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) mask |= (t[u][v] >= tcmin[v]) ? (1u << (u * 4 + v)) : 0u;
            }
            // the next group's loads go out before the hits of this one are queued
            p += 4 * HW;
            const int left = it1 - (it + 4);
            if (left >= 4) load_group(t, p, 4);
            else if (left > 0) load_group(t, p, left);
            while (mask) {
                const int k = __ffs(mask) - 1;
                mask &= mask - 1;
                const int slot = atomicAdd(&S.qcount, 1);
                if (slot < SEL_QCAP) S.queue[slot] = ((u32)(it + (k >> 2)) << 10) | (threadIdx.x << 2) | (u32)(k & 3);
            }
        }
    }
    __device__ __forceinline__ void eval(SelBuf &S, int q0, int q1, u64 thr) const {
        for (int q = q0 + threadIdx.x; q < q1; q += SEL_NT) {
            const u32 code = S.queue[q];
            const int v = code & 3, owner = (code >> 2) & 255, c = code >> 10;
            const VyScale &sc = hd->sc[S.it_scale[owner]];
            const float tv = vy_ldg32_ca(sc.head + S.it_off[owner] + (size_t)(5 + c) * (size_t)sc.HW + v);
            const float s = vy_score(tv, S.it_conf[v][owner]);
            if (s > valid_thresh) {
                const u64 key = vy_make_key(s, S.it_row0[owner] + (u32)c * (u32)sc.n_s + (u32)v * (u32)hd->A);
                if (key >= thr) sel_push(S, key);
            }
        }
    }
};

__global__ void __launch_bounds__(SEL_NT, SEL_CTAS_PER_SM)
vy_decode_select_kernel(const __grid_constant__ VyHeads hd, const __grid_constant__ SelPlan pl, SelGlobal g) {
    __shared__ SelBuf S;
    const int tid = threadIdx.x;
    vy_grid_dep_trigger();
    vy_grid_dep_wait();
    for (int job = blockIdx.x; job < pl.n_jobs; job += gridDim.x) {
        const int b = job / pl.G;
        SelJob jb;
        jb.G = pl.G; jb.g = job % pl.G; jb.K = pl.K; jb.Kq = pl.Kq;
        jb.g_thr_b = g.thr + b;
        jb.slots_b = g.slots + (size_t)b * pl.G;
        __syncthreads();
        if (tid == 0) { S.count = 0; S.qcount = 0; S.slot_pub = 0; S.thr = ld_relaxed_u64(jb.g_thr_b); }
        __syncthreads();
        const long long T = pl.items_per_frame;
        const int i0 = (int)(T * jb.g / pl.G), i1 = (int)(T * (jb.g + 1) / pl.G);
        int cnt = 0, fresh = 0;
        for (int base = i0; base < i1; base += SEL_NT) {
            const int idx = base + tid;
            HeadSrc src;
            src.hd = &hd;
            src.nv = 0; src.vec = false; src.p5 = nullptr; src.HW = 0; src.o1 = src.o2 = src.o3 = 0;
            src.valid_thresh = pl.valid_thresh;
            src.n_iters = hd.agnostic ? 0 : hd.C;
            float conf[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            u32 row0 = 0, A = (u32)hd.A;
            if (idx < i1) {
                int s = 0;
                while (s + 1 < hd.n_scales && idx >= pl.item_begin[s + 1]) ++s;
                const VyScale &sc = hd.sc[s];
                const int rel = idx - pl.item_begin[s];
                const int a = rel / pl.items_per_plane[s];
                const int pos0 = (rel % pl.items_per_plane[s]) * 4;
                src.nv = min(4, sc.HW - pos0);
                src.o1 = min(1, src.nv - 1); src.o2 = min(2, src.nv - 1); src.o3 = min(3, src.nv - 1);
                src.vec = sc.vec == 4;
                src.HW = (size_t)sc.HW;
                const size_t off = ((size_t)(b * hd.A + a) * hd.P) * src.HW + pos0;
                src.p5 = sc.head + off + 5 * src.HW;
                row0 = (u32)(sc.row_off + (long long)pos0 * hd.A + a);
                float to[4];
                src.load(to, sc.head + off + 4 * src.HW);
#pragma unroll
                for (int v = 0; v < 4; ++v) conf[v] = v < src.nv ? vy_sigmoid(to[v]) : 0.0f;
                S.it_off[tid] = (u32)off;
                S.it_row0[tid] = row0;
                S.it_scale[tid] = (u32)s;
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) S.it_conf[v][tid] = conf[v];
            if (hd.agnostic) {
                // yolo3.py:184-188: one candidate per box, score = objectness, row = off + pos*A + a
                if (cnt > SEL_CAP - 4 * SEL_NT) { cnt = sel_update(S, cnt, jb, true); fresh = 0; }
                const u64 thr = S.thr;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float s = conf[v];
                    if (v < src.nv && s > pl.valid_thresh) {
                        const u64 key = vy_make_key(s, row0 + (u32)v * A);
                        if (key >= thr) sel_push(S, key);
                    }
                }
                __syncthreads();
                cnt = S.count;
                __syncthreads();
                continue;
            }
            sel_stream(S, src, jb, cnt, fresh);
        }
        sel_flush(S, jb, g.count + b, g.list + (size_t)b * pl.list_cap, pl.list_cap);
    }
}

// ------------------------------------------------------------------------------------------------
// sample + stream: the bandwidth path for head maps.
//
//   vy_decode_sample_kernel   looks at 1/samp_stride of every image (pairs of adjacent items = 32-byte
//                             sectors, all class planes) and bounds the image's K-th largest score from
//                             below: ANY subset's K-th largest is <= the full set's.  Gs CTAs share an
//                             image; CTA g publishes its ceil(K/Gs)-th largest, the minimum over the
//                             Gs CTAs is the bound (disjoint subsets: >= K sampled scores lie above it).
//   vy_decode_stream_kernel   one pass over the head maps with that FIXED bound: no shared state, no
//                             CTA barrier; a warp owns a unit = 128 positions x one group of class
//                             planes, tests 16 bytes per lane and plane against a per-box logit bound
//                             and appends the few survivors (about K * samp_stride per image, whatever
//                             the score distribution) to the image's list through a warp-private buffer.
//   vy_decode_select_kernel   (above) runs afterwards for images whose list overflowed -- a sample
//                             that misjudged the distribution -- and normally exits at once.
// ------------------------------------------------------------------------------------------------
// The sample only ESTIMATES a score of rank ~4K (the result never depends on it), and its 68 sigmoids per thread
// were bound by the MUFU pipe at two ops each: one tanh.approx per sigmoid here (relative error ~2^-11).
__device__ __forceinline__ float samp_sigmoid(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}
constexpr int SAMP_NT = 512;
constexpr int SAMP_MAXK = 16;      // class planes per thread
constexpr int SAMP_BATCH = 8;      // of which this many are loaded together
#ifndef VY_SAMP_RUN
#define VY_SAMP_RUN 8
#endif
constexpr int SAMP_RUN = VY_SAMP_RUN;   // adjacent items sampled together: 8 x 16 B = one 128-byte line per plane

// sampled item j of image b -> where it lives (item = 4 consecutive positions of one (scale, anchor))
struct ItemRef { const float *p; int HW, nv, o1, o2, o3; bool vec; };
__device__ __forceinline__ bool item_ref(const VyHeads &hd, const SelPlan &pl, int b, int idx, ItemRef &r) {
    if (idx >= pl.items_per_frame) return false;
    int s = 0;
    while (s + 1 < hd.n_scales && idx >= pl.item_begin[s + 1]) ++s;
    const VyScale &sc = hd.sc[s];
    const int rel = idx - pl.item_begin[s];
    const int a = rel / pl.items_per_plane[s];
    const int pos0 = (rel - a * pl.items_per_plane[s]) * 4;
    r.HW = sc.HW;
    r.nv = min(4, sc.HW - pos0);
    r.o1 = min(1, r.nv - 1); r.o2 = min(2, r.nv - 1); r.o3 = min(3, r.nv - 1);
    r.vec = sc.vec == 4;
    r.p = sc.head + ((size_t)(b * hd.A + a) * hd.P) * (size_t)sc.HW + pos0;
    return true;
}
__device__ __forceinline__ void item_load(const ItemRef &r, const float *p, float (&t)[4]) {
    if (r.vec) { const float4 q = vy_ldg128_ca(p); t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w; }
    else { t[0] = vy_ldg32_ca(p); t[1] = vy_ldg32_ca(p + r.o1); t[2] = vy_ldg32_ca(p + r.o2); t[3] = vy_ldg32_ca(p + r.o3); }
}

// Job (b, g): item block ib = g % samp_ib, plane block pb = g / samp_ib.  Thread t owns sampled item
// (ib * samp_ipj + t % samp_ipj) and every samp_pls-th plane of the block, starting at t / samp_ipj:
// one objectness load and <= SAMP_MAXK class-plane loads, SAMP_BATCH of them in flight together.
// One 32-bit key per (item, plane) -- the score of the plane's best of <= 4 positions -- is parked in
// shared memory; the CTA then radix-selects its Ksq-th largest key.
__global__ void __launch_bounds__(SAMP_NT, 2)
vy_decode_sample_kernel(const __grid_constant__ VyHeads hd, const __grid_constant__ SelPlan pl, SelGlobal g) {
    __shared__ u32 skey[SAMP_MAXK][SAMP_NT];
    __shared__ u32 hist[256];
    __shared__ int sh_digit, sh_above, sh_in;
    const int tid = threadIdx.x;
    const int b = blockIdx.x / pl.Gs, gj = blockIdx.x % pl.Gs;
    const int ib = gj % pl.samp_ib, pb = gj / pl.samp_ib;
    const int jl = tid % pl.samp_ipj, lanep = tid / pl.samp_ipj;
    const int j = ib * pl.samp_ipj + jl;
    const int c_lo = pb * pl.samp_ppj, c_hi = min(hd.C, c_lo + pl.samp_ppj);
    // The CTA wants its Ksq-th largest key, and Ksq is small against the thread count (13 at COCO 608^2, 145 at VID 320^2):
    // a thread's two best keys stand for its <= 16 (the Ksq-th largest of a subset is <= that of the whole set, so the
    // bound can only come out lower = more careful, and it does only when one thread holds three of the CTA's Ksq
    // best: ~0.5 % of the threads at Ksq = 145).  The radix passes then walk 2 keys per thread instead of 16.
    // (CTA-uniform) every key takes part when Ksq is not small against the threads that hold keys (share > 1/2 per thread)
    const bool full = 2 * pl.Ksq > pl.samp_ipj * min(pl.samp_pls, SAMP_NT / max(pl.samp_ipj, 1));
    if (full) {
#pragma unroll
        for (int k = 0; k < SAMP_MAXK; ++k) skey[k][tid] = 0u;
    }
    u32 top0 = 0u, top1 = 0u;
    int n_mine = 0;
    ItemRef r;
    if (lanep < pl.samp_pls && j < pl.samp_items &&
        item_ref(hd, pl, b, (j / SAMP_RUN) * (SAMP_RUN * pl.samp_stride) + (j % SAMP_RUN), r)) {
        float to[4], cf[4];
        item_load(r, r.p + 4 * (size_t)r.HW, to);
        // the planes of the second batch: start them towards L2 now
#pragma unroll
        for (int k = SAMP_BATCH; k < SAMP_MAXK; ++k) {
            const int c = c_lo + lanep + k * pl.samp_pls;
            if (c < c_hi) asm volatile("prefetch.global.L2 [%0];" :: "l"(r.p + (size_t)(5 + c) * (size_t)r.HW));
        }
#pragma unroll
        for (int k0 = 0; k0 < SAMP_MAXK; k0 += SAMP_BATCH) {
            if (c_lo + lanep + k0 * pl.samp_pls >= c_hi) break;
            float tc[SAMP_BATCH][4];
#pragma unroll
            for (int k = 0; k < SAMP_BATCH; ++k) {
                const int c = c_lo + lanep + (k0 + k) * pl.samp_pls;
                if (c < c_hi) item_load(r, r.p + (size_t)(5 + c) * (size_t)r.HW, tc[k]);
            }
            if (k0 == 0) {
#pragma unroll
                for (int v = 0; v < 4; ++v) cf[v] = v < r.nv ? samp_sigmoid(to[v]) : 0.0f;
            }
            // sigmoid is monotonic but the objectness differs per position, so all <= 4 candidates are
            // scored; the best one is a score of the image, and any subset of an image's scores will do
#pragma unroll
            for (int k = 0; k < SAMP_BATCH; ++k) {
                const int c = c_lo + lanep + (k0 + k) * pl.samp_pls;
                if (c < c_hi) {
                    float best = 0.0f;
#pragma unroll
                    for (int v = 0; v < 4; ++v) if (v < r.nv) best = fmaxf(best, samp_sigmoid(tc[k][v]) * cf[v]);
                    if (best > pl.valid_thresh) {
                        const u32 key = vy_f2ord(best);
                        if (full) skey[k0 + k][tid] = key;
                        top1 = max(top1, min(top0, key));
                        top0 = max(top0, key);
                        ++n_mine;
                    }
                }
            }
        }
    }
    vy_grid_dep_trigger();                              // loads done: the rest of this CTA is shared-memory work
    // ---- CTA-wide: Ksq-th largest of the keys (MSB-first radix select over the parked keys); fewer than
    // Ksq keys in all (seen in the first pass: no bin reaches the rank) leave the bound at 0
    if (tid == 0) sh_digit = -1;
    u32 bound = 0;
    {
        u32 prefix = 0;
        int kk = pl.Ksq;
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            // run-length aggregation: neighbouring keys of a thread mostly share the leading digits
            u32 cur = 0xffffffffu, run = 0;
            auto visit = [&](u32 k) {
                if (k != 0u && (shift == 24 || ((k ^ prefix) >> (shift + 8)) == 0u)) {
                    const u32 d = (k >> shift) & 255u;
                    if (d != cur) { if (run) atomicAdd(&hist[cur], run); cur = d; run = 0; }
                    ++run;
                }
            };
            if (n_mine) {
                if (full) {
#pragma unroll
                    for (int i = 0; i < SAMP_MAXK; ++i) visit(skey[i][tid]);
                } else {
                    visit(top0);
                    visit(top1);
                }
            }
            if (run) atomicAdd(&hist[cur], run);
            __syncthreads();
            if (tid < 32) {
                u32 c[8], sum = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) { c[t] = hist[255 - 8 * tid - t]; sum += c[t]; }
                u32 inc = sum;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const u32 v = __shfl_up_sync(0xffffffffu, inc, off);
                    if (tid >= off) inc += v;
                }
                const u32 exc = inc - sum;
                if (exc < (u32)kk && (u32)kk <= inc) {
                    u32 run2 = exc;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (run2 + c[t] >= (u32)kk) { sh_digit = 255 - 8 * tid - t; sh_above = (int)run2; sh_in = (int)c[t]; break; }
                        run2 += c[t];
                    }
                }
            }
            __syncthreads();
            const int dig = sh_digit, inb = sh_in;
            if (dig < 0) { prefix = 0; break; }        // CTA-uniform: not enough keys
            prefix |= (u32)dig << shift;
            kk -= sh_above;
            __syncthreads();
            if (inb - kk <= (pl.Ksq >> 3)) break;      // #{keys >= prefix} is within Ksq/8 of Ksq
        }
        bound = prefix;
    }
    // ---- publish: the image's bound is the MINIMUM over its jobs, kept as the maximum of the complements
    // (the workspace header starts at zero = "no job yet" = complement of the largest bound)
    if (tid == 0) g.sslots[(size_t)b * pl.Gs + gj] = ~((u64)bound << 32);
    // the first job of an image also zeroes what the later kernels of the call accumulate into for that image (they
    // all start after this grid has finished): no memset in front of the call
    if (gj == 0) {
        if (tid == 0) { g.scount[b] = 0; g.count[b] = 0; g.thr[b] = 0ull; }
        if (tid < pl.G) g.slots[(size_t)b * pl.G + tid] = 0ull;
    }
}

constexpr int STR_NT = 256;
#ifndef STR_CTAS_PER_SM
#define STR_CTAS_PER_SM 4
#endif
#ifndef STR_UN
#define STR_UN 4
#endif

struct StrUnit {                 // what one warp streams: 128 positions x planes [c0, c1) of one (b, s, a)
    const float *pc;             // class plane 0 at this lane's first position
    size_t HW;
    int c0, c1, o1, o2, o3;
    u32 row0, n_s, A;
    u64 thr;
    float conf[4], tcmin[4];
    float valid_thresh;
};

template <bool VEC>
__device__ __forceinline__ void str_load(const StrUnit &un, const float *q, float (&t)[4]) {
    if (VEC) { const float4 w = vy_ldg128_ca(q); t[0] = w.x; t[1] = w.y; t[2] = w.z; t[3] = w.w; }
    else { t[0] = vy_ldg32_ca(q); t[1] = vy_ldg32_ca(q + un.o1); t[2] = vy_ldg32_ca(q + un.o2); t[3] = vy_ldg32_ca(q + un.o3); }
}

// rare path, warp-synchronous: every lane evaluates at most one hit per round; survivors go to the
// warp's buffer, which leaves as one atomic + one 256-byte store per 32 keys
// ring_group: the 128-bit path passes the ring slot of plane c (the logits are re-read from shared
// memory, LDS latency instead of an L2 round trip: cp.async.cg leaves nothing in L1); nullptr: re-read
// from global memory.
template <bool VEC>
__device__ __forceinline__ void str_hits(const StrUnit &un, u32 mask, int c, u64 *wbuf, int &cnt, int b,
                                         const SelGlobal &g, int lane, u32 lt_mask, const float4 *ring_group = nullptr) {
    while (__any_sync(0xffffffffu, mask != 0u)) {
        bool ok = false;
        u64 key = 0;
        if (mask) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            const int u = k >> 2, v = k & 3;
            const int ov = VEC ? v : (v == 0 ? 0 : (v == 1 ? un.o1 : (v == 2 ? un.o2 : un.o3)));
            const float cf = v == 0 ? un.conf[0] : (v == 1 ? un.conf[1] : (v == 2 ? un.conf[2] : un.conf[3]));
            const float tv = ring_group ? ((const float *)(ring_group + u * 32))[v]
                                        : vy_ldg32_ca(un.pc + (size_t)(c + u) * un.HW + ov);
            const float sv = vy_score(tv, cf);
            if (sv > un.valid_thresh) {
                key = vy_make_key(sv, un.row0 + (u32)(c + u) * un.n_s + (u32)v * un.A);
                ok = key >= un.thr;
            }
        }
        const u32 bal = __ballot_sync(0xffffffffu, ok);
        if (bal) {
            if (ok) wbuf[cnt + __popc(bal & lt_mask)] = key;
            cnt += __popc(bal);
            __syncwarp();
            if (cnt >= 32) {
                int base = 0;
                if (lane == 0) base = atomicAdd(g.scount + b, 32);
                base = __shfl_sync(0xffffffffu, base, 0);
                const u64 k0 = wbuf[lane], k1 = wbuf[32 + lane];
                if (base + lane < g.slist_cap) g.slist[(size_t)b * g.slist_cap + base + lane] = k0;
                __syncwarp();
                wbuf[lane] = k1;
                cnt -= 32;
                __syncwarp();
            }
        }
    }
}

template <bool VEC>
__device__ __forceinline__ void str_unit(const StrUnit &un, u64 *wbuf, int &cnt, int b, const SelGlobal &g,
                                         int lane, u32 lt_mask) {
    int c = un.c0;
    const float *q = un.pc + (size_t)c * un.HW;
    for (; c + STR_UN <= un.c1; c += STR_UN, q += (size_t)STR_UN * un.HW) {
        float t[STR_UN][4];
#pragma unroll
        for (int u = 0; u < STR_UN; ++u) str_load<VEC>(un, q + (size_t)u * un.HW, t[u]);
        u32 mask = 0;
#pragma unroll
        for (int u = 0; u < STR_UN; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) mask |= (t[u][v] >= un.tcmin[v]) ? (1u << (u * 4 + v)) : 0u;
        str_hits<VEC>(un, mask, c, wbuf, cnt, b, g, lane, lt_mask);
    }
    for (; c < un.c1; ++c, q += un.HW) {               // < STR_UN planes left
        float t[4];
        str_load<VEC>(un, q, t);
        u32 mask = 0;
#pragma unroll
        for (int v = 0; v < 4; ++v) mask |= (t[v] >= un.tcmin[v]) ? (1u << v) : 0u;
        str_hits<VEC>(un, mask, c, wbuf, cnt, b, g, lane, lt_mask);
    }
}

// ---- the 128-bit path keeps STR_NG groups of STR_UN planes per warp in flight with cp.async: a warp-private
// ring in shared memory (every lane reads back exactly the 16 bytes it fetched, so a per-thread
// cp.async.wait_group is all the synchronisation there is), refilled right after a group has been tested.
#ifndef STR_NG
#define STR_NG 3
#endif
constexpr int STR_RING = STR_NG * STR_UN;          // planes per warp ring

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((u32)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// issue the fetch of plane group gi (planes c0 + gi*STR_UN ...) of the unit into its ring slot; always
// commits exactly one group (an empty one past the end) so that the wait counts stay uniform
__device__ __forceinline__ void str_issue(const float *q0, size_t HW, int nplanes, int gi, float4 *ring_lane) {
    const int slot = (gi % STR_NG) * STR_UN;
#pragma unroll
    for (int u = 0; u < STR_UN; ++u) {
        const int pl_i = gi * STR_UN + u;
        if (pl_i < nplanes) cp_async16(ring_lane + (slot + u) * 32, q0 + (size_t)pl_i * HW);
    }
    cp_async_commit();
}

// ---- hits of the 128-bit path are QUEUED, not scored, where they are found: an entry names the element
// (class plane, owning lane, position of its float4), 32 bits in a warp-private queue.  Full batches of 32
// entries are then scored one per lane: the logit and the objectness come back from L2 (cp.async.cg left
// them there), the score is the decode kernel's, the key is tested against the image's bound and survivors go
// to the warp's key buffer as before.  A hit costs a few instructions in the streaming loop instead of a
// divergent detour through two sigmoids with one or two lanes active.
constexpr int STR_HQ = 64;                              // entries per warp (a batch leaves at 32)
__device__ __forceinline__ void str_hq_score(const StrUnit &un, const u32 *hq, int n, u64 *wbuf, int &cnt,
                                             int b, const SelGlobal &g, int lane, u32 lt_mask) {
    bool ok = false;
    u64 key = 0;
    // from lane 0's pointer / row (always a live lane; an idle lane's own ones point at position 0)
    const u32 row00 = __shfl_sync(0xffffffffu, un.row0, 0);
    const float *pc0 = (const float *)__shfl_sync(0xffffffffu, (unsigned long long)un.pc, 0);
    if (lane < n) {
        const u32 e = hq[lane];
        const int plane = (int)(e >> 7), ls = (int)((e >> 2) & 31u), v = (int)(e & 3u);
        const float *q = pc0 + ls * 4 + v;
        const float tv = vy_ldg32(q + (size_t)plane * un.HW);
        const float to = vy_ldg32(q - un.HW);                           // objectness plane = class plane 0 minus one plane
        const float sv = vy_score(tv, vy_sigmoid(to));
        if (sv > un.valid_thresh) {
            key = vy_make_key(sv, row00 + (u32)(ls * 4 + v) * un.A + (u32)plane * un.n_s);
            ok = key >= un.thr;
        }
    }
    const u32 bal = __ballot_sync(0xffffffffu, ok);
    if (bal) {
        if (ok) wbuf[cnt + __popc(bal & lt_mask)] = key;
        cnt += __popc(bal);
        __syncwarp();
        if (cnt >= 32) {
            int base = 0;
            if (lane == 0) base = atomicAdd(g.scount + b, 32);
            base = __shfl_sync(0xffffffffu, base, 0);
            const u64 k0 = wbuf[lane], k1 = wbuf[32 + lane];
            if (base + lane < g.slist_cap) g.slist[(size_t)b * g.slist_cap + base + lane] = k0;
            __syncwarp();
            wbuf[lane] = k1;
            cnt -= 32;
            __syncwarp();
        }
    }
}
// enqueue the set bits of `mask` (bit u*4+v = plane c + u, position v of this lane's float4); warp-uniform entry.
// One pass takes the lowest bit of every lane at once (the usual case: a hit or two in the group); what is
// left then sits in a few lanes (a confident box passes in many classes) and is unloaded lane by lane, the 16
// bits of the source lane's mask spread over 16 lanes.
__device__ __forceinline__ void str_hq_drain32(const StrUnit &un, u32 *hq, int &qn, u64 *wbuf, int &cnt, int b,
                                               const SelGlobal &g, int lane, u32 lt_mask) {
    str_hq_score(un, hq, 32, wbuf, cnt, b, g, lane, lt_mask);
    const u32 rest = hq[32 + lane];
    __syncwarp();
    hq[lane] = rest;
    qn -= 32;
    __syncwarp();
}
__device__ __forceinline__ void str_hq_push(const StrUnit &un, u32 mask, int c, u32 *hq, int &qn, u64 *wbuf,
                                            int &cnt, int b, const SelGlobal &g, int lane, u32 lt_mask) {
    {
        const bool has = mask != 0u;
        const u32 bal = __ballot_sync(0xffffffffu, has);
        if (has) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            hq[qn + __popc(bal & lt_mask)] = ((u32)(c + (k >> 2)) << 7) | ((u32)lane << 2) | (u32)(k & 3);
        }
        qn += __popc(bal);
        __syncwarp();
        if (qn >= 32) str_hq_drain32(un, hq, qn, wbuf, cnt, b, g, lane, lt_mask);
    }
    u32 left = __ballot_sync(0xffffffffu, mask != 0u);
    while (left) {
        const int src = __ffs(left) - 1;
        left &= left - 1;
        const u32 m = __shfl_sync(0xffffffffu, mask, src);          // <= 16 bits (STR_UN planes x 4 positions)
        if ((m >> lane) & 1u)
            hq[qn + __popc(m & lt_mask)] = ((u32)(c + (lane >> 2)) << 7) | ((u32)src << 2) | (u32)(lane & 3);
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) str_hq_drain32(un, hq, qn, wbuf, cnt, b, g, lane, lt_mask);
    }
}

// consume: the prologue (groups 0 .. STR_NG-1) has been issued by the caller.  The full groups run without
// any per-plane predicate, with a running ring slot and a running refill pointer (the loop is issue-sensitive:
// every instruction saved here is a memory request issued earlier); a partial last group takes the general path.
__device__ __forceinline__ void str_unit_async(const StrUnit &un, float4 *ring_lane, u32 *hq, u64 *wbuf, int &cnt,
                                               int b, const SelGlobal &g, int lane, u32 lt_mask) {
    int qn = 0;                                            // entries waiting in hq (warp-uniform)
    const int nplanes = un.c1 - un.c0;
    const int nfull = nplanes / STR_UN;                    // groups with all STR_UN planes
    const int ngroups = (nplanes + STR_UN - 1) / STR_UN;
    const float *q0 = un.pc + (size_t)un.c0 * un.HW;
    const size_t plane_bytes = un.HW * sizeof(float);
    const char *refill = (const char *)q0 + (size_t)(STR_NG * STR_UN) * plane_bytes;      // first plane of group gi + STR_NG
    int slot = 0;                                          // (gi % STR_NG) * STR_UN
    int gi = 0;
    for (; gi < nfull; ++gi) {
        cp_async_wait<STR_NG - 1>();
        float4 *rs = ring_lane + slot * 32;
        float t[STR_UN][4];
#pragma unroll
        for (int u = 0; u < STR_UN; ++u) {
            const float4 w = rs[u * 32];
            t[u][0] = w.x; t[u][1] = w.y; t[u][2] = w.z; t[u][3] = w.w;
        }
        u32 mask = 0;
#pragma unroll
        for (int u = 0; u < STR_UN; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) mask |= (t[u][v] >= un.tcmin[v]) ? (1u << (u * 4 + v)) : 0u;
        // the slot is free as soon as it has been tested: refill first, then look after the hits
        if (gi + STR_NG < nfull) {
#pragma unroll
            for (int u = 0; u < STR_UN; ++u) cp_async16(rs + u * 32, refill + (size_t)u * plane_bytes);
            cp_async_commit();
        } else {
            str_issue(q0, un.HW, nplanes, gi + STR_NG, ring_lane);
        }
        if (__any_sync(0xffffffffu, mask != 0u)) str_hq_push(un, mask, un.c0 + gi * STR_UN, hq, qn, wbuf, cnt, b, g, lane, lt_mask);
        refill += (size_t)STR_UN * plane_bytes;
        slot = slot + STR_UN == STR_RING ? 0 : slot + STR_UN;
    }
    for (; gi < ngroups; ++gi) {                           // at most one partial group
        cp_async_wait<STR_NG - 1>();
        float4 *rs = ring_lane + slot * 32;
        float t[STR_UN][4];
#pragma unroll
        for (int u = 0; u < STR_UN; ++u) {
            if (gi * STR_UN + u < nplanes) {
                const float4 w = rs[u * 32];
                t[u][0] = w.x; t[u][1] = w.y; t[u][2] = w.z; t[u][3] = w.w;
            } else {
                t[u][0] = t[u][1] = t[u][2] = t[u][3] = CUDART_NAN_F;      // NaN >= x is false for every x
            }
        }
        u32 mask = 0;
#pragma unroll
        for (int u = 0; u < STR_UN; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) mask |= (t[u][v] >= un.tcmin[v]) ? (1u << (u * 4 + v)) : 0u;
        if (__any_sync(0xffffffffu, mask != 0u)) str_hq_push(un, mask, un.c0 + gi * STR_UN, hq, qn, wbuf, cnt, b, g, lane, lt_mask);
        cp_async_commit();                                 // keeps the group count uniform
        slot = slot + STR_UN == STR_RING ? 0 : slot + STR_UN;
    }
    if (qn > 0) str_hq_score(un, hq, qn, wbuf, cnt, b, g, lane, lt_mask);
}

__global__ void __launch_bounds__(STR_NT, STR_CTAS_PER_SM)
vy_decode_stream_kernel(const __grid_constant__ VyHeads hd, const __grid_constant__ SelPlan pl, SelGlobal g) {
    __shared__ u64 wbuf_all[STR_NT / 32][64];
    __shared__ u32 hq_all[STR_NT / 32][STR_HQ];
    extern __shared__ __align__(16) unsigned char str_dyn[];          // [STR_NT/32][STR_RING][32] float4
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u64 *wbuf = wbuf_all[wid];
    float4 *ring_lane = (float4 *)str_dyn + (size_t)wid * STR_RING * 32 + lane;
    const u32 lt_mask = (1u << lane) - 1u;
    bool dep_waited = false;                            // the sample kernel's bounds are first read below
    const long long n_warps = (long long)gridDim.x * (STR_NT / 32);
    for (long long unit = (long long)blockIdx.x * (STR_NT / 32) + wid; unit < pl.n_units; unit += n_warps) {
        const int b = (int)(unit / pl.units_per_image);
        int r = (int)(unit - (long long)b * pl.units_per_image);
        int s = 0;
        while (s + 1 < hd.n_scales && r >= pl.unit_begin[s + 1]) ++s;
        r -= pl.unit_begin[s];
        const VyScale &sc = hd.sc[s];
        const int chunks = pl.chunks[s];
        const int q1 = r / chunks, chunk = r - q1 * chunks;
        const int a = q1 / pl.n_groups, grp = q1 - a * pl.n_groups;
        StrUnit un;
        un.c0 = grp * pl.PU;
        un.c1 = min(hd.C, un.c0 + pl.PU);
        un.HW = (size_t)sc.HW;
        int pos0 = (chunk * 32 + lane) * 4;
        int nv = min(4, sc.HW - pos0);
        if (nv <= 0) { nv = 0; pos0 = 0; }                 // idle lane: reads position 0, can never hit
        un.o1 = min(1, max(nv - 1, 0)); un.o2 = min(2, max(nv - 1, 0)); un.o3 = min(3, max(nv - 1, 0));
        const float *p = sc.head + ((size_t)(b * hd.A + a) * hd.P) * un.HW + pos0;
        un.pc = p + 5 * un.HW;
        const bool vec = sc.vec == 4;
        if (vec) {
            // the class planes do not depend on the bound: get them moving first
            const float *q0 = un.pc + (size_t)un.c0 * un.HW;
#pragma unroll
            for (int gi = 0; gi < STR_NG; ++gi) str_issue(q0, un.HW, un.c1 - un.c0, gi, ring_lane);
        }
        un.row0 = (u32)(sc.row_off + (long long)pos0 * hd.A + a);
        un.n_s = (u32)sc.n_s; un.A = (u32)hd.A;
        un.valid_thresh = pl.valid_thresh;
        if (!dep_waited) { vy_grid_dep_wait(); dep_waited = true; }
        un.thr = ~stream_bound_compl(g, b);
        const float smin = fmaxf(un.thr ? vy_key_score(un.thr) : pl.valid_thresh, pl.valid_thresh);
        {
            float to[4];
            if (vec) str_load<true>(un, p + 4 * un.HW, to); else str_load<false>(un, p + 4 * un.HW, to);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                un.conf[v] = vy_sigmoid(to[v]);
                un.tcmin[v] = v < nv ? vy_tcmin(smin, un.conf[v]) : CUDART_INF_F;
            }
        }
        int cnt = 0;                                       // keys waiting in wbuf (warp-uniform)
        if (vec) str_unit_async(un, ring_lane, hq_all[wid], wbuf, cnt, b, g, lane, lt_mask);
        else str_unit<false>(un, wbuf, cnt, b, g, lane, lt_mask);
        if (cnt > 0) {
            int base = 0;
            if (lane == 0) base = atomicAdd(g.scount + b, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lane < cnt && base + lane < g.slist_cap) g.slist[(size_t)b * g.slist_cap + base + lane] = wbuf[lane];
            __syncwarp();
        }
    }
    vy_grid_dep_trigger();                              // (a trigger at the start lets the dependents crowd the tail: measured slower)
}

#ifdef VY_STREAM_ALT
#include "vy_stream_alt.cuh"
#endif

// ------------------------------------------------------------------------------------------------
// source 2: materialised detection rows (generic box_nms).  Iteration = SEL_NT consecutive rows.
// ------------------------------------------------------------------------------------------------
struct RowSrc {
    int n_iters;
    const float *img;            // data + b*R*W
    long long row_begin, row_end;
    int W, score_index, id_index, background_id;
    float valid_thresh, smin;

    __device__ __forceinline__ void refresh(SelBuf &, u64 thr) {
        smin = thr ? vy_key_score(thr) : -CUDART_INF_F;
    }
    __device__ __forceinline__ void run(SelBuf &S, int it0, int it1) {
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
                    const int slot = atomicAdd(&S.qcount, 1);
                    if (slot < SEL_QCAP) S.queue[slot] = (u32)(r[u] - row_begin);
                }
            }
        }
    }
    __device__ __forceinline__ void eval(SelBuf &S, int q0, int q1, u64 thr) const {
        for (int q = q0 + threadIdx.x; q < q1; q += SEL_NT) {
            const long long r = row_begin + S.queue[q];
            const float s = img[r * W + score_index];
            if (id_index >= 0 && background_id >= 0 && (int)img[r * W + id_index] == background_id) continue;
            const u64 key = vy_make_key(s, (u32)r);
            if (key >= thr) sel_push(S, key);
        }
    }
};

__global__ void __launch_bounds__(SEL_NT, SEL_CTAS_PER_SM)
vy_rows_select_kernel(RowParams rp, SelPlan pl, SelGlobal g) {
    __shared__ SelBuf S;
    const int tid = threadIdx.x;
    for (int job = blockIdx.x; job < pl.n_jobs; job += gridDim.x) {
        const int b = job / pl.G;
        SelJob jb;
        jb.G = pl.G; jb.g = job % pl.G; jb.K = pl.K; jb.Kq = pl.Kq;
        jb.g_thr_b = g.thr + b;
        jb.slots_b = g.slots + (size_t)b * pl.G;
        __syncthreads();
        if (tid == 0) { S.count = 0; S.qcount = 0; S.slot_pub = 0; S.thr = ld_relaxed_u64(jb.g_thr_b); }
        __syncthreads();
        RowSrc src;
        src.img = rp.data + (size_t)b * (size_t)rp.R * rp.W;
        src.row_begin = jb.g * pl.rows_per_job;
        src.row_end = min(rp.R, src.row_begin + pl.rows_per_job);
        src.W = rp.W; src.score_index = rp.score_index; src.id_index = rp.id_index;
        src.background_id = rp.background_id; src.valid_thresh = rp.valid_thresh;
        src.n_iters = src.row_end > src.row_begin ? (int)((src.row_end - src.row_begin + SEL_NT - 1) / SEL_NT) : 0;
        int cnt = 0, fresh = 0;
        sel_stream(S, src, jb, cnt, fresh);
        sel_flush(S, jb, g.count + b, g.list + (size_t)b * pl.list_cap, pl.list_cap);
    }
}

// ------------------------------------------------------------------------------------------------
// finalize: one CTA per image
// ------------------------------------------------------------------------------------------------
constexpr int FIN_NT_MAX = 1024;     // the kernel runs with 512 or 1024 threads (blockDim.x)
constexpr int FIN_SLACK = 512;       // candidates beyond K that may reach the ranking sort
constexpr int FIN_CMAX = 256;       // head-map classes up to this count are regrouped by counting instead of sorting
constexpr int FIN_LCAP = 2048;       // candidate lists up to this length are staged in shared memory (general front)
// static shared memory of the finalize kernel (the selection kernels' SelBuf carries 24 KB it has no use for: the
// smaller the CTA's footprint, the more streaming CTAs of a neighbouring stream stay resident beside it)
struct FinBuf {
    u64 keys[FIN_NT_MAX];        // the K best by rank
    u32 hist[256];
    u64 mm[2];                   // min / max key of a list
    u64 thr;                     // inclusive lower bound of the general front
    int count, flag;
    int sel_digit, sel_above, sel_in;
};

struct FinParams {
    int K, post_rows;            // rows written per image
    long long out_stride_rows;   // rows per image in `out` (== post_rows)
    float overlap_thresh;
    float thr_lo, thr_hi;        // nms_suppresses_fast: where the reciprocal estimate decides (set by launch_finalize)
    int force_suppress, in_format, out_format;
    int W;                       // output row width (6 for heads)
    int fill_rest;               // 1: this kernel writes the -1 padding rows itself
    int lcap;                    // candidate lists up to this length are staged in shared memory (0: never)
    int force_rescue;            // (tools/rescue_time.py, VY_FORCE_RESCUE) treat every streamed list as unusable
    float *out;
    int *kept_rows;
};

// BoxArea / Intersect arithmetic: vy_nms_math.cuh

// source row -> class id (and, for head maps, where its box logits live)
template <int SRC>
__device__ __forceinline__ int fin_class(const VyHeads &hd, const RowParams &rp, int b, u32 row) {
    if (SRC == 0) {
        if (hd.agnostic) return 0;
        int s = 0;
        while (s + 1 < hd.n_scales && (long long)row >= hd.sc[s + 1].row_off) ++s;
        return (int)((row - (u32)hd.sc[s].row_off) / (u32)hd.sc[s].n_s);
    } else {
        if (rp.id_index < 0) return 0;
        return (int)rp.data[((size_t)b * (size_t)rp.R + row) * rp.W + rp.id_index];
    }
}

template <int SRC>
__device__ __forceinline__ float4 fin_box(const VyHeads &hd, const RowParams &rp, int b, u32 row) {
    if (SRC == 0) {
        int s = 0;
        while (s + 1 < hd.n_scales && (long long)row >= hd.sc[s + 1].row_off) ++s;
        const VyScale &sc = hd.sc[s];
        const u32 rem = (row - (u32)sc.row_off) % (u32)sc.n_s;
        const int pos = (int)(rem / (u32)hd.A), a = (int)(rem % (u32)hd.A);
        const int y = pos / sc.W, x = pos % sc.W;
        const size_t HW = (size_t)sc.HW;
        const float *p = sc.head + ((size_t)(b * hd.A + a) * hd.P) * HW + pos;
        return vy_box(p[0], p[HW], p[2 * HW], p[3 * HW], x, y, sc.stride, sc.aw[a], sc.ah[a]);
    } else {
        const float *p = rp.data + ((size_t)b * (size_t)rp.R + row) * rp.W + rp.coord_start;
        return make_float4(p[0], p[1], p[2], p[3]);
    }
}

// Lower bound p of the K-th largest key of a list, with K <= #{keys >= p} <= K + slack (MSB-first radix
// select, one sweep of the list per 8-bit digit).  n >= K, slack >= 0.  The keys of a candidate list agree in
// their leading bits (scores of one narrow range), so a first sweep takes the list's minimum and maximum and
// the digits start right below the common prefix: one counting sweep usually settles the bound.
static __device__ __noinline__ u64 fin_list_bound(FinBuf &S, const u64 *list, int n, int K, int slack) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    u64 *mm = S.mm;                                     // [0] = min, [1] = max
    if (tid == 0) { mm[0] = ~0ull; mm[1] = 0ull; }
    __syncthreads();
    {
        u64 lo = ~0ull, hi = 0ull;
        for (int i = tid; i < n; i += nt) { const u64 k = list[i]; lo = k < lo ? k : lo; hi = k > hi ? k : hi; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const u64 l2 = sel_shfl_xor_u64(lo, off), h2 = sel_shfl_xor_u64(hi, off);
            lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
        }
        if (lane == 0) { atomicMin((unsigned long long *)&mm[0], (unsigned long long)lo); atomicMax((unsigned long long *)&mm[1], (unsigned long long)hi); }
    }
    __syncthreads();
    const u64 kmin = mm[0], kmax = mm[1];
    if (kmin == kmax) return kmin;
    const int hb = 63 - __clzll((long long)(kmin ^ kmax));      // highest bit in which two keys differ
    int shift = hb >= 7 ? hb - 7 : 0;
    u64 prefix = shift + 8 >= 64 ? 0ull : (kmax >> (shift + 8)) << (shift + 8);
    int kk = K;
    for (;;) {
        if (tid < 256) S.hist[tid] = 0;
        __syncthreads();
        // run-length aggregation: neighbouring keys of a thread often share the digit, and same-address
        // shared-memory atomics would serialise
        u32 cur = 0xffffffffu, run = 0;
        for (int i = tid; i < n; i += nt) {
            const u64 k = list[i];
            if (shift + 8 >= 64 || ((k ^ prefix) >> (shift + 8)) == 0ull) {
                const u32 d = (u32)(k >> shift) & 255u;
                if (d != cur) { if (run) atomicAdd(&S.hist[cur], run); cur = d; run = 0; }
                ++run;
            }
        }
        if (run) atomicAdd(&S.hist[cur], run);
        __syncthreads();
        if (tid < 32) {
            u32 c[8], sum = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) { c[t] = S.hist[255 - 8 * tid - t]; sum += c[t]; }
            u32 inc = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const u32 v = __shfl_up_sync(0xffffffffu, inc, off);
                if (tid >= off) inc += v;
            }
            const u32 exc = inc - sum;
            if (exc < (u32)kk && (u32)kk <= inc) {
                u32 run2 = exc;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (run2 + c[t] >= (u32)kk) { S.sel_digit = 255 - 8 * tid - t; S.sel_above = (int)run2; S.sel_in = (int)c[t]; break; }
                    run2 += c[t];
                }
            }
        }
        __syncthreads();
        const int d = S.sel_digit, above = S.sel_above, inb = S.sel_in;
        __syncthreads();
        prefix |= (u64)d << shift;                      // (a digit window that overlaps the previous one repeats its bits)
        kk -= above;
        if (inb - kk <= slack || shift == 0) break;
        shift = shift >= 8 ? shift - 8 : 0;
    }
    return prefix;
}

// Front end of the finalize kernel for lists of <= FIN_BK_KPT keys per thread: the exact top-K, sorted, in three light
// passes instead of bound + compaction + sort.  The keys of a list agree in their leading bits (scores of one
// narrow range), so FIN_BK_BINS counting bins laid right below the common prefix hold about one key each:
// histogram, descending scan (= where every bin starts in the sorted order, and which bin holds the K-th
// key), scatter into the bins, and a rank inside each bin (usually of one or two keys).  Every thread keeps
// its <= FIN_BK_KPT keys in registers throughout; the list is read once.
// Returns m1 = #{keys in the bins down to the K-th key's} (>= min(n, K)) with keyr[0 .. min(m1, K)) sorted
// descending, or -1 (CTA-uniform, nothing written) when that many keys would not fit the CTA.
#ifdef VY_FIN_TIMING
__device__ long long vy_fin_front_clk[16];
#define FIN_TB(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) vy_fin_front_clk[k] = clock64(); } while (0)
extern "C" int vy_debug_fin_front_clocks(long long *out) {
    return cudaMemcpyFromSymbol(out, vy_fin_front_clk, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
#else
#define FIN_TB(k) do { } while (0)
#endif
constexpr int FIN_BK_BINS = 2048;
constexpr int FIN_BK_KPT_MAX = 16;    // keys per thread: 8 (lists <= 4096 at 512 threads) or 16
template <int FIN_BK_KPT>
static __device__ __noinline__ int fin_front_buckets(FinBuf &S, const u64 *list, int n, int K,
                                                     u32 *hist, u32 *excl, u64 *out, u64 *keyr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = (int)blockDim.x;   // nt = 512 or 1024
    u64 *mm = S.mm;                                     // [0] = min, [1] = max
    FIN_TB(0);
    u64 k[FIN_BK_KPT];
#pragma unroll
    for (int q = 0; q < FIN_BK_KPT; ++q) { const int i = tid + q * nt; k[q] = i < n ? list[i] : 0ull; }
    for (int i = tid; i < FIN_BK_BINS; i += nt) hist[i] = 0u;
    if (tid == 0) { mm[0] = ~0ull; mm[1] = 0ull; S.sel_digit = 0; S.sel_in = n; }
    __syncthreads();
    FIN_TB(1);
    {
        u64 lo = ~0ull, hi = 0ull;
#pragma unroll
        for (int q = 0; q < FIN_BK_KPT; ++q) if (k[q]) { lo = k[q] < lo ? k[q] : lo; hi = k[q] > hi ? k[q] : hi; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const u64 l2 = sel_shfl_xor_u64(lo, off), h2 = sel_shfl_xor_u64(hi, off);
            lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
        }
        if (lane == 0 && hi) { atomicMin((unsigned long long *)&mm[0], (unsigned long long)lo); atomicMax((unsigned long long *)&mm[1], (unsigned long long)hi); }
    }
    __syncthreads();
    FIN_TB(2);
    const u64 kx = mm[0] ^ mm[1];
    const int hb = kx ? 63 - __clzll((long long)kx) : 0;     // highest bit in which two keys differ
    const int shift = hb >= 10 ? hb - 10 : 0;
#pragma unroll
    for (int q = 0; q < FIN_BK_KPT; ++q) if (k[q]) atomicAdd(&hist[(u32)(k[q] >> shift) & (FIN_BK_BINS - 1)], 1u);
    __syncthreads();
    FIN_TB(3);
    // descending scan: thread t owns the `per` bins from 2047 - per * t downwards (per = 2 or 4)
    const int per = FIN_BK_BINS / nt;
    const int d0 = FIN_BK_BINS - 1 - per * tid;
    u32 h[4], hsum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { h[j] = j < per ? hist[d0 - j] : 0u; hsum += h[j]; }
    u32 incl = hsum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    u32 *wsum = S.hist;                                 // [32]
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const u32 v = lane < (nt >> 5) ? wsum[lane] : 0u;
        u32 inc2 = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 y = __shfl_up_sync(0xffffffffu, inc2, off);
            if (lane >= off) inc2 += y;
        }
        wsum[lane] = inc2 - v;
    }
    __syncthreads();
    {
        u32 run = wsum[warp] + incl - hsum;             // keys in the bins above d0
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < per) {
                excl[d0 - j] = run;
                if (run < (u32)K && (u32)K <= run + h[j]) {     // the K-th key's bin (none: fewer than K keys, the defaults stand)
                    S.sel_digit = d0 - j;
                    S.sel_in = (int)(run + h[j]);
                }
                run += h[j];
            }
        }
    }
    __syncthreads();
    FIN_TB(4);
    const int kbin = S.sel_digit, m1 = S.sel_in;
    if (m1 > nt) return -1;
#pragma unroll
    for (int q = 0; q < FIN_BK_KPT; ++q) {
        if (k[q]) {
            const u32 bin = (u32)(k[q] >> shift) & (FIN_BK_BINS - 1);
            if ((int)bin >= kbin) out[excl[bin] + atomicSub(&hist[bin], 1u) - 1u] = k[q];
        }
    }
    __syncthreads();
    FIN_TB(5);
    if (tid < m1) {
        const u64 key = out[tid];
        const u32 bin = (u32)(key >> shift) & (FIN_BK_BINS - 1);
        const u32 a = excl[bin], e = bin ? excl[bin - 1] : (u32)n;
        u32 rank = a;
        for (u32 q = a; q < e; ++q) rank += out[q] > key ? 1u : 0u;
        if (rank < (u32)K) keyr[rank] = key;
    }
    __syncthreads();
    FIN_TB(6);
    return m1;
}

// -DVY_FIN_TIMING (tools/fin_phases.py only): CTA 0 records clock64() at the phase boundaries
#ifdef VY_FIN_TIMING
#define FIN_T(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) vy_fin_clk[k] = clock64(); } while (0)
__device__ long long vy_fin_cta[1024][4];          // per CTA: cycles, SM id, list length, m1
extern "C" int vy_debug_fin_clocks(long long *out) {
    return cudaMemcpyFromSymbol(out, vy_fin_clk, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
extern "C" int vy_debug_fin_ctas(long long *out, int n) {
    return cudaMemcpyFromSymbol(out, vy_fin_cta, sizeof(long long) * 4 * n) == cudaSuccess ? 0 : -1;
}
#else
#define FIN_T(k) do { } while (0)
#endif

template <int SRC>
__device__ __forceinline__ void fin_box_prefetch(const VyHeads &hd, const RowParams &rp, int b, u32 row) {
    if (SRC == 0) {
        int s = 0;
        while (s + 1 < hd.n_scales && (long long)row >= hd.sc[s + 1].row_off) ++s;
        const VyScale &sc = hd.sc[s];
        const u32 rem = (row - (u32)sc.row_off) % (u32)sc.n_s;
        const int pos = (int)(rem / (u32)hd.A), a = (int)(rem % (u32)hd.A);
        const size_t HW = (size_t)sc.HW;
        const float *p = sc.head + ((size_t)(b * hd.A + a) * hd.P) * HW + pos;
#pragma unroll
        for (int k = 0; k < 4; ++k) asm volatile("prefetch.global.L2 [%0];" :: "l"(p + k * HW));
    } else {
        const float *p = rp.data + ((size_t)b * (size_t)rp.R + row) * rp.W + rp.coord_start;
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
    }
}

// unique keys (the callers pad with distinct non-zero values): rank merges when the CTA has a thread per key
__device__ __forceinline__ void fin_sort(u64 *keys, int npow2) {
    if (npow2 <= (int)blockDim.x) sel_sort_desc_merge(keys, npow2);
    else sel_sort_desc_fast(keys, npow2);
}

// General front end: bound the K-th largest key (lists longer than K + slack), compact what is at or above the
// bound into cand, sort.  Returns m1 <= K + FIN_SLACK with cand[0 .. m1) sorted descending.
static __device__ __noinline__ int fin_front_general(FinBuf &S, const u64 *list, int n_list, int K, u64 *cand,
                                                     u64 *lbuf, int lcap) {
    const int tid = threadIdx.x, lane = tid & 31, FIN_NT = (int)blockDim.x;
    const u32 lt_mask = (1u << lane) - 1u;
    // the ranking sort works on a power-of-two buffer: let through what fills the one K + 64 needs anyway
    int slack;
    { int t2 = 64; while (t2 < K + 64) t2 <<= 1; slack = min(min(t2, FIN_NT_MAX) - K, FIN_SLACK); }
    if (n_list > K + slack && n_list <= lcap) {
        // the radix sweeps below then never leave the SM
        for (int i = tid; i < n_list; i += FIN_NT) lbuf[i] = list[i];
        list = lbuf;
    }
    __syncthreads();
    if (n_list > K + slack) {
        // long list: bound its K-th largest key first, so that one sweep leaves <= K + slack keys
        const u64 p = fin_list_bound(S, list, n_list, K, slack);
        const u64 cur = S.thr;
        __syncthreads();
        if (tid == 0 && p > cur) S.thr = p;
        __syncthreads();
    }
    {
        const u64 thr = S.thr;
        for (int i0 = 0; i0 < n_list; i0 += FIN_NT) {   // CTA-uniform trip count: one shared-memory atomic per warp
            const int i = i0 + tid;
            const u64 key = i < n_list ? list[i] : 0ull;
            const bool in = i < n_list && key >= thr;
            const u32 bal = __ballot_sync(0xffffffffu, in);
            if (bal) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&S.count, __popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                const int slot = base + __popc(bal & lt_mask);
                if (in && slot < K + FIN_SLACK) cand[slot] = key;
            }
        }
    }
    __syncthreads();
    const int m1 = min(S.count, K + FIN_SLACK);
    // rank: keys are unique (the row is part of the key), so a descending sort IS the operator's
    // stable order; the buffer is padded with distinct values below every real key
    int np2 = 32;
    while (np2 < m1) np2 <<= 1;
    for (int i = m1 + tid; i < np2; i += FIN_NT) cand[i] = (u64)(np2 - i);
    __syncthreads();
    fin_sort(cand, np2);
    return m1;
}

// One CTA of FIN_NT_MAX >= K threads per image.  Positions: "rank" = place in the global score order (what
// the operator's output order is); "slot" = place after a stable regrouping by class, in which every class
// is one contiguous segment still ordered by rank.  Suppression only ever happens inside a segment, so the
// IoU tests run over (segment length)^2 pairs instead of K^2, and segments that share no 32-slot block are
// resolved by different warps.
// ---- exact rescue of an image whose streamed candidate list is unusable (overflow, or fewer than K keys under a
// non-trivial bound: the sample misjudged the score distribution; probability ~1e-7 per image with random-init or
// trained-like logits, certain for adversarial inputs such as all-equal scores).  The image's CTA redoes the selection
// alone: MSB-first radix select over the 64-bit keys of EVERY (box, class) of the image, keys recomputed from the head
// maps in each pass (same arithmetic as the streaming pass: vy_score(t_c, sigmoid(t_obj)), row = the reference's), until
// the keys at or above the prefix fit the list; one more pass collects them.  Slow (a pass reads the whole image with
// one CTA) and rare; it replaces a separate rescue launch that sat on the critical path of every call.
template <class F>
__device__ __forceinline__ void fin_for_each_key(const VyHeads &hd, const SelPlan &pl, int b, F f) {
    for (int idx = threadIdx.x; idx < pl.items_per_frame; idx += blockDim.x) {
        int s = 0;
        while (s + 1 < hd.n_scales && idx >= pl.item_begin[s + 1]) ++s;
        const VyScale &sc = hd.sc[s];
        const int rel = idx - pl.item_begin[s];
        const int a = rel / pl.items_per_plane[s];
        const int pos0 = (rel - a * pl.items_per_plane[s]) * 4;
        const int nv = min(4, sc.HW - pos0);
        const bool vec = sc.vec == 4;                      // (then nv == 4: HW is a multiple of 4)
        const int o1 = min(1, nv - 1), o2 = min(2, nv - 1), o3 = min(3, nv - 1);
        const float *p = sc.head + ((size_t)(b * hd.A + a) * hd.P) * (size_t)sc.HW + pos0;
        const u32 row0 = (u32)(sc.row_off + (long long)pos0 * hd.A + a);
        auto load4 = [&](const float *q, float (&t)[4]) {
            if (vec) { const float4 w = vy_ldg128(q); t[0] = w.x; t[1] = w.y; t[2] = w.z; t[3] = w.w; }
            else { t[0] = vy_ldg32(q); t[1] = vy_ldg32(q + o1); t[2] = vy_ldg32(q + o2); t[3] = vy_ldg32(q + o3); }
        };
        float conf[4];
        load4(p + 4 * (size_t)sc.HW, conf);
#pragma unroll
        for (int v = 0; v < 4; ++v) conf[v] = v < nv ? vy_sigmoid(conf[v]) : 0.0f;
        // four class planes at a time: the loads first (the pass is a chain of L2 round trips otherwise)
        for (int c0 = 0; c0 < hd.C; c0 += 4) {
            float t[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + u < hd.C) load4(p + (size_t)(5 + c0 + u) * (size_t)sc.HW, t[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (c0 + u >= hd.C) break;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (v < nv) {
                        const float sv = vy_score(t[u][v], conf[v]);
                        if (sv > pl.valid_thresh) f(vy_make_key(sv, row0 + (u32)(c0 + u) * (u32)sc.n_s + (u32)v * (u32)hd.A));
                    }
                }
            }
        }
    }
}
// the same walk with the streaming pass's logit-domain test in front: f(key) for every element whose score can exceed
// `smin` (t_c >= vy_tcmin(smin, conf): one compare per element, the sigmoids only for what passes)
template <class F>
__device__ __forceinline__ void fin_for_each_key_above(const VyHeads &hd, const SelPlan &pl, int b, float smin, F f) {
    for (int idx = threadIdx.x; idx < pl.items_per_frame; idx += blockDim.x) {
        int s = 0;
        while (s + 1 < hd.n_scales && idx >= pl.item_begin[s + 1]) ++s;
        const VyScale &sc = hd.sc[s];
        const int rel = idx - pl.item_begin[s];
        const int a = rel / pl.items_per_plane[s];
        const int pos0 = (rel - a * pl.items_per_plane[s]) * 4;
        const int nv = min(4, sc.HW - pos0);
        const bool vec = sc.vec == 4;
        const int o1 = min(1, nv - 1), o2 = min(2, nv - 1), o3 = min(3, nv - 1);
        const float *p = sc.head + ((size_t)(b * hd.A + a) * hd.P) * (size_t)sc.HW + pos0;
        const u32 row0 = (u32)(sc.row_off + (long long)pos0 * hd.A + a);
        auto load4 = [&](const float *q, float (&t)[4]) {
            if (vec) { const float4 w = vy_ldg128(q); t[0] = w.x; t[1] = w.y; t[2] = w.z; t[3] = w.w; }
            else { t[0] = vy_ldg32(q); t[1] = vy_ldg32(q + o1); t[2] = vy_ldg32(q + o2); t[3] = vy_ldg32(q + o3); }
        };
        float conf[4], tcm[4];
        load4(p + 4 * (size_t)sc.HW, conf);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            conf[v] = v < nv ? vy_sigmoid(conf[v]) : 0.0f;
            tcm[v] = v < nv ? vy_tcmin(smin, conf[v]) : CUDART_INF_F;
        }
        for (int c0 = 0; c0 < hd.C; c0 += 4) {
            float t[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (c0 + u < hd.C) load4(p + (size_t)(5 + c0 + u) * (size_t)sc.HW, t[u]);
                else t[u][0] = t[u][1] = t[u][2] = t[u][3] = CUDART_NAN_F;
            }
            if (!vy_any_ge16(t, tcm)) continue;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (t[u][v] >= tcm[v]) {
                        const float sv = vy_score(t[u][v], conf[v]);
                        if (sv > pl.valid_thresh) f(vy_make_key(sv, row0 + (u32)(c0 + u) * (u32)sc.n_s + (u32)v * (u32)hd.A));
                    }
        }
    }
}

// returns the number of keys written to out[0 .. cap); *lower = a lower bound of the image's K-th largest key (0: none).
// priv: 16 x 256 words of shared memory -- a histogram per warp (two warps share one in a 1024-thread CTA): in the
// degenerate case every key of the image matches the prefix in every pass, and 1.8 M atomics on ONE histogram were the
// pass (1 ms); the CTA-wide histogram is S.hist.
static __device__ __noinline__ int fin_rescue_heads(FinBuf &S, const VyHeads &hd, const SelPlan &pl, int b, int K, u64 *out,
                                                    int cap, u32 *priv, u64 *lower) {
    const int tid = threadIdx.x;
    u32 *mine = priv + (((tid >> 5) & 15) << 8);
    const int limit = cap < 4096 ? cap : 4096;           // (<= 4096 keys keep the finalize on its bucket-sort front end)
    // First try: every VALID key of the image (score > valid_thresh) in one pass, found with the streaming pass's cheap
    // logit-domain test.  With trained-like logits -- where a bound that was too high is met in practice -- that is a few
    // thousand keys and the rescue ends here; random-init logits overflow the list at once and take the radix passes.
    if (tid == 0) S.count = 0;
    __syncthreads();
    fin_for_each_key_above(hd, pl, b, pl.valid_thresh, [&](u64 key) {
        if (*(volatile int *)&S.count <= cap) {
            const int at = atomicAdd(&S.count, 1);
            if (at < cap) out[at] = key;
        }
    });
    __syncthreads();
    {
        const int n_valid = S.count;
        __syncthreads();
        if (n_valid <= cap) { *lower = 0ull; return n_valid; }
    }
    u64 prefix = 0;
    int kk = K;
    int shift = 56;
    for (;; shift -= 8) {
        for (int i = tid; i < 16 * 256; i += blockDim.x) priv[i] = 0u;
        __syncthreads();
        u64 kmin = ~0ull, kmax = 0ull;                     // of the keys that match the prefix
        {
            // run-length aggregation: neighbouring keys of a thread mostly share the leading digits
            u32 cur = 0xffffffffu, run = 0;
            const int sh = shift;
            const u64 pf = prefix;
            fin_for_each_key(hd, pl, b, [&](u64 key) {
                if (sh == 56 || ((key ^ pf) >> (sh + 8)) == 0ull) {
                    const u32 d = (u32)(key >> sh) & 255u;
                    if (d != cur) { if (run) atomicAdd(&mine[cur], run); cur = d; run = 0; }
                    ++run;
                    kmin = key < kmin ? key : kmin;
                    kmax = key > kmax ? key : kmax;
                }
            });
            if (run) atomicAdd(&mine[cur], run);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const u64 lo = sel_shfl_xor_u64(kmin, off), hi = sel_shfl_xor_u64(kmax, off);
            kmin = lo < kmin ? lo : kmin;
            kmax = hi > kmax ? hi : kmax;
        }
        if (tid == 0) { S.mm[0] = ~0ull; S.mm[1] = 0ull; }
        __syncthreads();
        if ((tid & 31) == 0) { atomicMin((unsigned long long *)&S.mm[0], (unsigned long long)kmin); atomicMax((unsigned long long *)&S.mm[1], (unsigned long long)kmax); }
        __syncthreads();
        if (tid < 256) {
            u32 sum = 0;
#pragma unroll
            for (int h = 0; h < 16; ++h) sum += priv[(h << 8) + tid];
            S.hist[tid] = sum;
        }
        __syncthreads();
        if (tid == 0) {
            u32 above = 0;
            int dig = -1;
            for (int d = 255; d >= 0; --d) {
                if (above + S.hist[d] >= (u32)kk) { dig = d; break; }
                above += S.hist[d];
            }
            S.sel_digit = dig; S.sel_above = (int)above; S.sel_in = dig >= 0 ? (int)S.hist[dig] : 0;
        }
        __syncthreads();
        const int dig = S.sel_digit;
        if (dig < 0) { prefix = 0; break; }              // (first pass only) fewer than K valid keys in the image: take them all
        const u32 above = (u32)S.sel_above, inb = (u32)S.sel_in;
        prefix |= (u64)dig << shift;
        kk -= (int)above;
        const long long ge = (long long)(K - kk) + inb;  // keys at or above the prefix (its lower bits zero)
        // every matching key in this one bin (all-equal scores: the keys differ only in their row bits): the bytes the
        // smallest and the largest of them share tell nothing either -- go straight to the first byte that differs
        const u64 diff = S.mm[0] ^ S.mm[1];
        __syncthreads();
        if (ge <= limit || shift == 0) break;
        if (above == 0 && diff != 0ull && S.hist[dig] == inb) {
            u32 total = 0;                                // (CTA-uniform: every thread reads the same histogram)
            for (int d = 0; d < 256; ++d) total += S.hist[d];
            const int top = 63 - __clzll((long long)diff);                 // highest differing bit
            const int next = (top >> 3) << 3;
            if (total == inb && next < shift - 8) {
                const u64 keep = next + 8 >= 64 ? 0ull : ~((1ull << (next + 8)) - 1ull);
                prefix = S.mm[0] & keep;
                shift = next + 8;                         // (the loop's decrement lands on `next`)
            }
        }
        __syncthreads();
    }
    // collect
    if (tid == 0) S.count = 0;
    __syncthreads();
    {
        const u64 pf = prefix;
        fin_for_each_key(hd, pl, b, [&](u64 key) {
            if (key >= pf) {
                const u32 at = (u32)atomicAdd(&S.count, 1);
                if (at < (u32)cap) out[at] = key;
            }
        });
    }
    __syncthreads();
    const u32 n = (u32)S.count;
    __syncthreads();
    *lower = prefix;
    return (int)(n < (u32)cap ? n : (u32)cap);
}

template <int SRC>   // 0: head maps, 1: rows
__global__ void __launch_bounds__(FIN_NT_MAX)
vy_nms_finalize_kernel(const __grid_constant__ VyHeads hd, const __grid_constant__ RowParams rp, const __grid_constant__ SelPlan pl,
                       const __grid_constant__ SelGlobal g, const __grid_constant__ FinParams fp) {
    __shared__ FinBuf S;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (int)blockDim.x / 32;
    const int FIN_NT = (int)blockDim.x;
    const int b = blockIdx.x;
    const int K = pl.K;
    const int nwK = (K + 31) >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    FIN_T(0);
#ifdef VY_FIN_TIMING
    const long long t_start = clock64();
#endif
    int cp2 = 32;
    while (cp2 < K + FIN_SLACK) cp2 <<= 1;
    u64 *cand = (u64 *)dyn;                            // cp2 >= K + FIN_SLACK candidates (sorted in place)
    float4 *box = (float4 *)(cand + cp2);              // K      by slot
    float *area = (float *)(box + K);                  // K      by slot
    int *cls = (int *)(area + K);                      // K      by slot
    int *seg_end = cls + K;                            // K      by slot
    int *slot_of_rank = seg_end + K;                   // K
    int *rank_of_slot = slot_of_rank + K;              // K
    u32 *begw = (u32 *)(rank_of_slot + K);             // 32     by slot: first slot of a segment
    u32 *keeps = begw + 32;                            // 32     by slot: survivors
    u32 *keepw = keeps + 32;                           // 32     by rank: survivors
    int *rb0 = (int *)(keepw + 32);                    // 32 (+4 pad): first slot block whose segments reach block cb
    // one region, two lives: the front end's scratch (FIN_LCAP keys: bin starts + scattered keys, or a staged list),
    // then -- the front end is over by then -- the bit matrix of the tiled suppression path
    u64 *lbuf = (u64 *)(((uintptr_t)(rb0 + 36) + 15) & ~(uintptr_t)15);
    u32 *supby = (u32 *)lbuf;                          // K * nwK: [j][rb] = rows of slot block rb that would suppress slot j

    // ---- 1. exact top-K of the image's candidate list, sorted descending
    // streaming path: the streamed list, unless it was unusable and the rescue pass rebuilt g.list
    vy_grid_dep_wait();
    const bool use_s = g.scount != nullptr;
    int n_list = use_s ? g.scount[b] : min(g.count[b], pl.list_cap);
    const u64 *list = use_s ? g.slist + (size_t)b * g.slist_cap : g.list + (size_t)b * pl.list_cap;
    u64 thr0 = use_s ? ~stream_bound_compl(g, b) : g.thr[b];
    if (SRC == 0 && use_s && (fp.force_rescue || !stream_list_ok(g, b, K))) {
        // unusable streamed list: this CTA redoes the image's selection exactly (fin_rescue_heads), into the same list
        n_list = fin_rescue_heads(S, hd, pl, b, K, g.slist + (size_t)b * g.slist_cap, g.slist_cap, (u32 *)lbuf, &thr0);
        __syncthreads();
    }
    if (tid == 0) { S.count = 0; S.flag = 0; S.thr = thr0; }
    if (tid < 32) keeps[tid] = 0u;
    u64 *keyr = S.keys;                                 // the K best by rank (K <= FIN_NT_MAX)
    int m1 = -1;
    if (fp.lcap >= FIN_BK_BINS) {
        if (n_list <= 8 * FIN_NT) m1 = fin_front_buckets<8>(S, list, n_list, K, (u32 *)cand, (u32 *)lbuf, lbuf + FIN_BK_BINS / 2, keyr);
        else if (n_list <= FIN_BK_KPT_MAX * FIN_NT)      // trained-like logits: lists of 4-8 K keys are common
            m1 = fin_front_buckets<16>(S, list, n_list, K, (u32 *)cand, (u32 *)lbuf, lbuf + FIN_BK_BINS / 2, keyr);
    }
    FIN_T(1);
    u64 mykey = 0ull;
    if (m1 < 0) {
        // general path (lists beyond the bucket front end's reach): bound, compact, sort
        m1 = fin_front_general(S, list, n_list, K, cand, lbuf, fp.lcap);
        if (tid < min(m1, K)) { mykey = cand[tid]; keyr[tid] = mykey; }
    } else {
        if (tid < min(m1, K)) mykey = keyr[tid];
    }
    FIN_T(2); FIN_T(3);
    const int m = min(m1, K);                           // <= K candidates take part
    const int nw = (m + 31) >> 5;
    // thread i < m owns rank i from here on (FIN_NT >= K)
    if (tid < m) fin_box_prefetch<SRC>(hd, rp, b, vy_key_row(mykey));    // the box logits travel while the classes are sorted
    FIN_T(4);

    // ---- 2. regroup by class, stable in rank
    const bool all_pairs = fp.force_suppress || (SRC == 1 && rp.id_index < 0) || (SRC == 0 && hd.agnostic);
    // head maps with a bounded class count: counting sort (per-warp class counts, match_any for the place
    // inside the warp); anything else: ascending sort of (class, rank) pairs
    const bool counted = !all_pairs && SRC == 0 && hd.C <= FIN_CMAX;
    if (counted) {
        const int C = hd.C;
        unsigned short *cntw = (unsigned short *)cand;  // [nw][C]; the ranked keys live in keyr / registers by now
        u32 *cstart = S.hist;                           // [C]
        __syncthreads();                                // every thread has read its cand[tid]
        for (int i = tid; i < nw * C; i += FIN_NT) cntw[i] = 0;
        __syncthreads();
        int c = 0;
        if (tid < m) c = fin_class<SRC>(hd, rp, b, vy_key_row(mykey));
        const u32 peers = __match_any_sync(0xffffffffu, tid < m ? c : -1 - lane);
        const int intra = __popc(peers & lt_mask);
        if (tid < m && intra == 0) cntw[warp * C + c] = (unsigned short)__popc(peers);
        __syncthreads();
        if (tid < C) {                                  // per class: exclusive prefix over the warps, total
            int run = 0;
            for (int w = 0; w < nw; ++w) { const int v = cntw[w * C + tid]; cntw[w * C + tid] = (unsigned short)run; run += v; }
            cstart[tid] = (u32)run;
        }
        __syncthreads();
        if (warp == 0) {                                // exclusive scan of the class totals (C <= 256: 8 per lane)
            u32 v[FIN_CMAX / 32], sum = 0;
#pragma unroll
            for (int t = 0; t < FIN_CMAX / 32; ++t) { const int idx = lane * (FIN_CMAX / 32) + t; v[t] = idx < C ? cstart[idx] : 0u; sum += v[t]; }
            u32 incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const u32 y = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += y;
            }
            u32 run = incl - sum;
#pragma unroll
            for (int t = 0; t < FIN_CMAX / 32; ++t) { const int idx = lane * (FIN_CMAX / 32) + t; if (idx < C) { cstart[idx] = run; run += v[t]; } }
        }
        __syncthreads();
        if (tid < m) {
            const int slot = (int)cstart[c] + (int)cntw[warp * C + c] + intra;
            slot_of_rank[tid] = slot;
            rank_of_slot[slot] = tid;
            cls[slot] = c;
        }
        __syncthreads();
    } else if (!all_pairs) {
        int mp2 = 32;
        while (mp2 < m) mp2 <<= 1;
        // (the read of cand[tid] above and this write are by the same thread)
        for (int i = tid; i < mp2; i += FIN_NT) {
            u64 k = (u64)(mp2 - i);                     // padding: distinct, below every real key = last in ascending order
            if (i < m) k = ~(((u64)((u32)fin_class<SRC>(hd, rp, b, vy_key_row(mykey)) ^ 0x80000000u) << 32) | (u64)i);
            cand[i] = k;
        }
        __syncthreads();
        fin_sort(cand, mp2);                            // descending in ~key = ascending in (class, rank)
    } else {
        __syncthreads();                                // keyr
    }
    FIN_T(5);
    // thread j < m owns slot j: its box, and whether it opens a segment
    bool beg = false;
    {
        if (tid < m) {
            int rank = tid, c = 0;
            beg = tid == 0;
            if (counted) {
                rank = rank_of_slot[tid];
                c = cls[tid];
                if (tid > 0) beg = cls[tid - 1] != c;
            } else if (!all_pairs) {
                const u64 k = ~cand[tid];
                rank = (int)(u32)(k & 0xffffffffull);
                c = (int)((u32)(k >> 32) ^ 0x80000000u);
                if (tid > 0) beg = (u32)(~cand[tid - 1] >> 32) != (u32)(k >> 32);
                cls[tid] = c;
                rank_of_slot[tid] = rank;
                slot_of_rank[rank] = tid;
            } else {
                cls[tid] = 0;
                rank_of_slot[tid] = tid;
                slot_of_rank[tid] = tid;
            }
            const float4 bx = fin_box<SRC>(hd, rp, b, vy_key_row(keyr[rank]));
            box[tid] = bx;
            area[tid] = nms_area(bx, fp.in_format);
        }
        const u32 bw = __ballot_sync(0xffffffffu, beg);
        if (lane == 0) begw[warp] = bw;
    }
    __syncthreads();
    FIN_T(6);
    if (tid < m) {                                      // segment end = next opening slot
        int w = warp;
        u32 bits = lane == 31 ? 0u : (begw[w] >> (lane + 1)) << (lane + 1);
        int e = m;
        for (;;) {
            if (bits) { e = (w << 5) + __ffs(bits) - 1; break; }
            if (++w >= nw) break;
            bits = begw[w];
        }
        seg_end[tid] = e;
        if (beg) atomicMax(&S.flag, e - tid);           // longest segment
    }
    if (tid < nw) {                                     // block of the segment start of slot 32 * tid (slot 0 opens one)
        int w = tid;
        if (!(begw[w] & 1u)) { do { --w; } while (begw[w] == 0u); }
        rb0[tid] = w;
    }
    __syncthreads();
    FIN_T(7);

    const bool short_segs = S.flag <= 64;
    if (short_segs) {
        // ---- 3s/4s. segments of <= 64 slots (the per-class case): a lane per reference slot i tests its followers
        // i+1 .. in step, the hits of a slot are a 64-bit word relative to it; then the first slot of every segment
        // runs the greedy pass over its segment alone
        u64 *rowrel = cand;                             // [m] (cand is free from here on)
        u64 rel = 0ull;
        int e = 0;
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
        float ai = 0.f;
        if (tid < m) { e = seg_end[tid]; bi = box[tid]; ai = area[tid]; }
        const int trip = __reduce_max_sync(0xffffffffu, tid < m ? e - tid - 1 : 0);
        for (int k = 1; k <= trip; ++k) {
            const int j = tid + k;
            if (j < e && nms_suppresses_fast(bi, ai, box[j], area[j], fp.overlap_thresh, fp.thr_lo, fp.thr_hi, fp.in_format))
                rel |= 1ull << (k - 1);
        }
        if (tid < m) rowrel[tid] = rel;
        __syncthreads();
        FIN_T(8);
        if (tid < m && beg) {
            const int len = e - tid;
            u64 removed = 0ull, kept = 0ull;
            for (int k = 0; k < len; ++k) {
                if (!((removed >> k) & 1ull)) { kept |= 1ull << k; removed |= (rowrel[tid + k] << 1) << k; }
            }
            const int sh = tid & 31, w0 = tid >> 5;
            const u64 lo = kept << sh, hi = sh ? kept >> (64 - sh) : 0ull;
            if ((u32)lo) atomicOr(&keeps[w0], (u32)lo);
            if ((u32)(lo >> 32)) atomicOr(&keeps[w0 + 1], (u32)(lo >> 32));
            if ((u32)hi) atomicOr(&keeps[w0 + 2], (u32)hi);
        }
    } else {
    // ---- 3. suppression tests in 32 x 32 tiles: tile (rb, cb) exists when a segment reaches from slot block rb
    // into slot block cb; lane = candidate slot j of block cb, rows i of block rb in turn (a row no segment of
    // which reaches the block is skipped by the whole warp)
    {
        const int r0 = lane < nw ? rb0[lane] : 0;
        const int cnt = lane < nw ? lane - r0 + 1 : 0;
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int t = warp; t < total; t += nwarps) {
            const int cb = __popc(__ballot_sync(0xffffffffu, lane < nw && incl <= t));
            const int rb = __shfl_sync(0xffffffffu, r0, cb) + (t - __shfl_sync(0xffffffffu, incl - cnt, cb));
            const int j = (cb << 5) + lane;
            float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
            float aj = 0.f;
            if (j < m) { bj = box[j]; aj = area[j]; }
            u32 d = 0;
            const int i0 = rb << 5, c0 = cb << 5;
            // rows of the block with a follower in this column block (lane = row), then two rows per step
            const int myrow = i0 + lane;
            const int e_l = myrow < m ? seg_end[myrow] : 0;
            u32 rows = __ballot_sync(0xffffffffu, e_l > max(myrow + 1, c0));
            while (rows) {
                const int ia = __ffs(rows) - 1;
                rows &= rows - 1;
                const int ib = rows ? __ffs(rows) - 1 : ia;
                rows &= rows - 1;                       // (0 & anything stays 0)
                const int ea = __shfl_sync(0xffffffffu, e_l, ia), eb = __shfl_sync(0xffffffffu, e_l, ib);
                const float4 ba = box[i0 + ia], bb = box[i0 + ib];
                const float aa = area[i0 + ia], ab = area[i0 + ib];
                // both tests unconditionally (side by side in the pipeline); lanes outside a row's segment drop theirs
                const bool ha = nms_suppresses_fast(ba, aa, bj, aj, fp.overlap_thresh, fp.thr_lo, fp.thr_hi, fp.in_format);
                const bool hb = nms_suppresses_fast(bb, ab, bj, aj, fp.overlap_thresh, fp.thr_lo, fp.thr_hi, fp.in_format);
                if (ha && j > i0 + ia && j < ea) d |= 1u << ia;
                if (hb && j > i0 + ib && j < eb) d |= 1u << ib;     // (ib == ia repeats the same bit)
            }
            if (j < m) supby[(size_t)j * nwK + rb] = d;
        }
    }
    __syncthreads();
    FIN_T(8);

    // ---- 4. greedy resolution, one warp per chain of slot blocks linked by a straddling segment (block cb
    // starts a chain when rb0[cb] == cb).  Lane = slot; keeps[rb] of the earlier blocks of a chain were written
    // by this warp.
    if (warp < nw && rb0[warp] == warp) {
        for (int cb = warp; cb < nw; ++cb) {
            const int r0 = rb0[cb];
            if (cb > warp && r0 == cb) break;
            const int j = (cb << 5) + lane;
            bool rem = false;
            u32 d = 0;
            if (j < m) {
                for (int rb = r0; rb < cb; ++rb) rem |= (supby[(size_t)j * nwK + rb] & keeps[rb]) != 0u;
                d = supby[(size_t)j * nwK + cb];
            }
            const bool live = j < m && !rem;
            u32 keepbits = __ballot_sync(0xffffffffu, live && d == 0u);
            u32 pend = __ballot_sync(0xffffffffu, live && d != 0u);
            while (pend) {                              // in slot order: everything below bit i is final
                const int i = __ffs(pend) - 1;
                pend &= pend - 1;
                if (__shfl_sync(0xffffffffu, (int)((d & keepbits) == 0u), i)) keepbits |= 1u << i;
            }
            if (lane == 0) keeps[cb] = keepbits;
            __syncwarp();
        }
    }
    }
    __syncthreads();
    FIN_T(9);
    // survivors by rank
    u32 kb;
    {
        bool kept = false;
        if (tid < m) { const int slot = slot_of_rank[tid]; kept = (keeps[slot >> 5] >> (slot & 31)) & 1u; }
        kb = __ballot_sync(0xffffffffu, kept);
        if (lane == 0) keepw[warp] = kb;
    }
    __syncthreads();
    int n_keep, p_base;
    {
        const int c = lane < nw ? __popc(keepw[lane]) : 0;
        int incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        n_keep = __shfl_sync(0xffffffffu, incl, 31);
        p_base = __shfl_sync(0xffffffffu, incl - c, warp);
    }
    FIN_T(10);

    // ---- 5. survivors to the front in score order; the rest is -1
    const int W = fp.W;
    float *out_b = fp.out + (size_t)b * (size_t)fp.out_stride_rows * W;
    int *kept_b = fp.kept_rows ? fp.kept_rows + (size_t)b * (size_t)fp.out_stride_rows : nullptr;
    if (tid < m && ((kb >> lane) & 1u)) {
        const int i = tid;
        const int p = p_base + __popc(kb & lt_mask);
        if (p < fp.post_rows) {
            const u64 key = mykey;
            const u32 row = vy_key_row(key);
            float *o = out_b + (size_t)p * W;
            if (SRC == 0) {
                const int j = slot_of_rank[i];
                const float4 bx = box[j];
                o[0] = (float)(all_pairs && !hd.agnostic ? fin_class<SRC>(hd, rp, b, row) : cls[j]);
                o[1] = vy_key_score(key);
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
    }
    if (fp.fill_rest) {
        const int first = min(n_keep, fp.post_rows);
        const long long total = (long long)(fp.post_rows - first) * W;
        float *o = out_b + (size_t)first * W;
        for (long long i = tid; i < total; i += FIN_NT) o[i] = -1.0f;
        if (kept_b) for (int i = first + tid; i < fp.post_rows; i += FIN_NT) kept_b[i] = -1;
    }
    FIN_T(11);
#ifdef VY_FIN_TIMING
    if (tid == 0 && b < 1024) {
        u32 smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        vy_fin_cta[b][0] = clock64() - t_start; vy_fin_cta[b][1] = smid; vy_fin_cta[b][2] = n_list; vy_fin_cta[b][3] = m1;
    }
#endif
}


__global__ void vy_fill_kernel(float *out, int *kept, size_t n_out, size_t n_kept) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) out[i] = -1.0f;
    if (kept) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_kept; i += stride) kept[i] = -1;
}

// ------------------------------------------------------------------------------------------------
// host: planning + launches
// ------------------------------------------------------------------------------------------------
// shared memory of the finalize kernel; *lcap = how long a staged candidate list may be
static size_t fin_dyn_smem(int K, int *lcap) {
    const int nwK = (K + 31) / 32;
    size_t cp2 = 32;
    while (cp2 < (size_t)(K + FIN_SLACK)) cp2 <<= 1;
    const size_t base = cp2 * 8 + (size_t)K * (16 + 4 + 4 + 4 + 4 + 4) + 3 * 32 * 4 + 36 * 4 + 16;
    // front-end scratch and suppression bit matrix share one region (K <= 1024: at most 128 KB + 53 KB, within the CTA limit)
    const size_t mat = (size_t)K * nwK * 4, scratch = (size_t)FIN_LCAP * 8;
    if (lcap) *lcap = FIN_LCAP;
    return base + (mat > scratch ? mat : scratch);
}

// CTAs that can be resident at once (SEL_CTAS_PER_SM per SM by __launch_bounds__)
static int resident_ctas() { return SEL_CTAS_PER_SM * vy_sm_count(); }

static void plan_jobs(int B, long long units_per_image, SelPlan *pl) {
    // G CTAs per image: fill the machine in one wave, at least ~SEL_NT units per job, <= 32 slots
    long long G = resident_ctas() / B;
    const long long gmax_units = (units_per_image + SEL_NT - 1) / SEL_NT;
    if (G > gmax_units) G = gmax_units;
    if (G > SEL_GMAX) G = SEL_GMAX;
    if (G < 1) G = 1;
    pl->G = (int)G;
    pl->Kq = (pl->K + pl->G - 1) / pl->G;
    pl->n_jobs = B * pl->G;
    const long long cap = (long long)pl->G * (pl->K + (pl->K >> 2));
    pl->list_cap = (int)(cap < 64 ? 64 : cap);
}

static void plan_stream(const VyHeads &hd, SelPlan *pl);

static int plan_heads(const VyHeads &hd, int topk, float valid_thresh, SelPlan *pl) {
    memset(pl, 0, sizeof(*pl));
    long long K = topk < 0 ? hd.R : (topk < hd.R ? topk : hd.R);
    if (K < 1 || K > SEL_KMAX) return VY_EUNSUPPORTED;
    pl->K = (int)K;
    pl->valid_thresh = valid_thresh;
    int items = 0;
    for (int s = 0; s < hd.n_scales; ++s) {
        pl->items_per_plane[s] = (hd.sc[s].HW + 3) / 4;
        pl->item_begin[s] = items;
        items += pl->items_per_plane[s] * hd.A;
    }
    for (int s = hd.n_scales; s <= VY_MAX_SCALES; ++s) pl->item_begin[s] = items;
    pl->items_per_frame = items;
    plan_jobs(hd.B, items, pl);
    plan_stream(hd, pl);
    return VY_OK;
}

// streaming path: worthwhile once an image has enough rows for a sample to be cheap AND sharp
static void plan_stream(const VyHeads &hd, SelPlan *pl) {
    pl->stream = 0;
    if (hd.agnostic || hd.R < 131072 || hd.C < 1) return;
#ifdef VY_STREAM_ALT
    if (hd.B > 0xffff || hd.A > 8 || hd.n_scales > 4) return;             // fields of the tile-streaming pass's metadata word
#endif
    // sampled fraction 1/S
    long long S = hd.R / (40LL * pl->K);
    if (S < 4) return;
    if (S > 32) S = 32;
    const long long resident = 2LL * vy_sm_count();                     // __launch_bounds__(SAMP_NT, 2)
    for (;; S *= 2) {
        const long long runs = ((long long)pl->items_per_frame + SAMP_RUN * S - 1) / (SAMP_RUN * S);
        const long long items = runs * SAMP_RUN;
        // jobs: item blocks of <= SAMP_NT items x plane blocks; a thread owns <= SAMP_MAXK planes
        const long long ib = (items + SAMP_NT - 1) / SAMP_NT;
        const long long ipj = (items + ib - 1) / ib;
        const long long pls = SAMP_NT / ipj;                            // >= 1
        const long long pbk_min = (hd.C + pls * SAMP_MAXK - 1) / (pls * SAMP_MAXK);
        long long pbk_max = (hd.C + pls * 8 - 1) / (pls * 8);           // no fewer than ~8 planes per thread
        if (pbk_max < pbk_min) pbk_max = pbk_min;
        // as many plane blocks as still fit the whole grid in ONE wave of resident CTAs
        long long pbk = resident / ((long long)hd.B * ib);
        if (pbk > pbk_max) pbk = pbk_max;
        if (pbk < pbk_min) pbk = pbk_min;
        if (ib * pbk > SEL_GMAX) pbk = SEL_GMAX / ib;
        if (pbk < pbk_min || pbk < 1) continue;                         // too many item blocks: sample less
        const long long ppj = (hd.C + pbk - 1) / pbk;
        pbk = (hd.C + ppj - 1) / ppj;
        if (ppj > pls * SAMP_MAXK) continue;                            // too many planes per thread
        if (ib * pbk > pl->K) continue;
        pl->samp_stride = (int)S;
        pl->samp_items = (int)items;
        pl->samp_ib = (int)ib; pl->samp_ipj = (int)ipj;
        pl->samp_ppj = (int)ppj; pl->samp_pls = (int)pls;
        pl->Gs = (int)(ib * pbk);
        break;
    }
    // The bound handed to the streaming pass is the minimum over the Gs jobs of each job's Ksq-th largest
    // sampled score.  With Ksq = ceil(K/Gs) that is a guaranteed lower bound of the image's K-th largest
    // score (disjoint subsets hold >= K scores above it) but lets ~K*S candidates through.  Aiming the
    // sample at rank ~4K instead (j = 4K/S sampled scores) cuts the candidate lists ~S/4-fold; the estimate
    // is below the K-th largest score except with negligible probability (rank std ~ S*sqrt(j) << 3K), and
    // an image where it is not (fewer than K candidates found under a non-trivial bound) is redone exactly
    // by the rescue pass (stream_list_ok).
    // How far above K?  The estimate's rank has a relative spread ~ 1/sqrt(j) (j sampled scores above the bound, times a
    // burstiness factor for clustered logits), so the margin in standard deviations is (1 - 1/aim) * sqrt(aim * K / S):
    // 5.3 = aim 4K at S = 32, K = 400.  Where the sampling rate is higher the same margin needs a lower aim -- 2.4K at VID
    // 320^2 (S = 11): 43 % shorter lists, the call 6 % faster (tools/aim_test.py).
    // Clustered (trained-like) logits make the sample burstier than Poisson, the more so the fewer 128-byte runs of a
    // plane it holds: 416^2 images (11 sampled runs; 608^2: 23, VID 320^2: 18) x 80 classes x 128 sent 4 of 3 072 images
    // to the rescue at 5.3 (0.4 ms each; no other measured configuration rescued any) and none at 5.8 (aim 4.5K) -- which
    // costs every configuration 1-2 % (longer lists, more hits), so only images with fewer than 16 sampled runs pay it.
    // VY_SAMP_SIGMA sets the margin for all, VY_SAMP_AIM the aim itself (A/B).
    static const double aim_env = getenv("VY_SAMP_AIM") ? atof(getenv("VY_SAMP_AIM")) : 0.0;
    static const double sigma_env = getenv("VY_SAMP_SIGMA") ? atof(getenv("VY_SAMP_SIGMA")) : 0.0;
    const double sigma = sigma_env > 0.0 ? sigma_env : (pl->samp_items / SAMP_RUN < 16 ? 5.8 : 5.3);
    double aim = 6.0;
    for (double a = 1.5; a < 6.0; a += 0.1)
        if ((1.0 - 1.0 / a) * sqrt(a * pl->K / (double)pl->samp_stride) >= sigma) { aim = a; break; }
    if (aim_env > 0.0) aim = aim_env;
    long long j = (long long)((aim * pl->K + pl->samp_stride - 1) / pl->samp_stride);
    if (j > pl->K) j = pl->K;
    long long ksq = (j + pl->Gs - 1) / pl->Gs;
    if (ksq < 8) ksq = 8;
    const long long ksq_sure = ((long long)pl->K + pl->Gs - 1) / pl->Gs;
    if (ksq > ksq_sure) ksq = ksq_sure;
    pl->Ksq = (int)ksq;
    // A unit = 128 positions x one group of class planes, streamed by one warp.  The class planes are cut
    // into as many groups as make the unit count a near-multiple of the resident warps: every warp walks
    // ceil(units / warps) units, so 2.47 "waves" cost as much as 3 (each extra group costs a unit prologue).
    int chunks_total = 0;
    for (int s = 0; s < hd.n_scales; ++s) {
        pl->chunks[s] = (hd.sc[s].HW + 127) / 128;
        chunks_total += pl->chunks[s];
    }
    {
        const double warps = (double)STR_CTAS_PER_SM * vy_sm_count() * (256 / 32);
        const int g_min = (hd.C + 95) / 96;
        int best_g = g_min;
        double best = -1.0;
        for (int ng = g_min; ng <= g_min + 3; ++ng) {
            if (ng > g_min && (hd.C + ng - 1) / ng < 16) break;           // groups of fewer than 16 planes are all prologue
            const double waves = (double)chunks_total * hd.A * ng * hd.B / warps;
            const double full = waves <= 1.0 ? 1.0 : (double)(long long)(waves + 0.999999);
            const double score = (waves <= 1.0 ? 1.0 : waves / full) * (1.0 - 0.02 * (ng - g_min));
            if (score > best) { best = score; best_g = ng; }
        }
        pl->n_groups = best_g;
        // Small batches leave most warp slots empty (VID 416^2 x 8 windows: 720 units for 4736 slots) and the pass is as
        // long as its busiest unit -- all hits of an (anchor, class) plane whose logits sit above the others, a
        // confident region -- so the planes are cut further, down to groups of 4 (one round of the ring), as long as
        // (nearly) every unit still gets a warp of its own: the prologue (objectness + bounds) is repeated per group, in
        // parallel.  Up to 1.3 units per warp slot: VID 416^2 x 32 with two groups (1.22) 29.3 -> 22.2 us trained-like,
        // 19.0 -> 18.9 us random-init; 2.0 units per slot: 23.0 / 21.1 us.
        static const int min_pu = getenv("VY_STR_MIN_PU") ? atoi(getenv("VY_STR_MIN_PU")) : 4;     // 0: off (A/B)
        static const double split_waves = getenv("VY_STR_SPLIT_WAVES") ? atof(getenv("VY_STR_SPLIT_WAVES")) : 1.3;
        if (min_pu > 0 && (double)chunks_total * hd.A * g_min * hd.B < warps) {
            for (int pu = min_pu; pu < hd.C; pu += 4) {
                const int ng = (hd.C + pu - 1) / pu;
                if (ng <= g_min) break;
                if ((double)chunks_total * hd.A * ng * hd.B <= warps * split_waves) { pl->n_groups = ng; break; }
            }
        }
    }
    pl->PU = (hd.C + pl->n_groups - 1) / pl->n_groups;
    int units = 0;
    for (int s = 0; s < hd.n_scales; ++s) {
        pl->unit_begin[s] = units;
        units += pl->chunks[s] * pl->n_groups * hd.A;
    }
    for (int s = hd.n_scales; s <= VY_MAX_SCALES; ++s) pl->unit_begin[s] = units;
    pl->units_per_image = units;
    pl->n_units = (long long)units * hd.B;
#ifdef VY_STREAM_ALT
    // tile streaming: tiles per block, table segments
    int tiles = 0, tf = 0;
    pl->tab_max = 0;
    for (int s = 0; s < hd.n_scales; ++s) {
        // a tile never starts in the middle of a plane unless the plane is longer than a tile (then at multiples of
        // the tile size): the table index of a lane's elements is then the same for every tile of a block
        const int TF = S2_TILE / 4, HWs = hd.sc[s].HW;
        if (HWs >= TF) {
            pl->tile_tpp[s] = (HWs + TF - 1) / TF; pl->tile_ppt[s] = 0;
            pl->tiles_blk[s] = hd.C * pl->tile_tpp[s];
        } else {
            pl->tile_tpp[s] = 0; pl->tile_ppt[s] = TF / HWs;
            pl->tiles_blk[s] = (hd.C + pl->tile_ppt[s] - 1) / pl->tile_ppt[s];
        }
        pl->tile_begin[s] = tiles;
        tiles += pl->tiles_blk[s] * hd.A;
        pl->tab_hwp[s] = (hd.sc[s].HW + 3 + 3) & ~3;
        pl->tab_off[s] = tf;
        tf += pl->tab_hwp[s] * hd.A;
        if (pl->tab_hwp[s] > pl->tab_max) pl->tab_max = pl->tab_hwp[s];
    }
    for (int s = hd.n_scales; s <= VY_MAX_SCALES; ++s) { pl->tile_begin[s] = tiles; pl->tab_off[s] = tf; }
    for (int s = 0; s < hd.n_scales; ++s) {               // as many groups as tables of this scale fit beside each other
        static const int divs[] = {12, 6, 4, 3, 2, 1};
        int ng = 1;
        for (int d : divs) if (S3_WARPS % d == 0 && (long long)d * pl->tab_hwp[s] <= pl->tab_max) { ng = d; break; }
        pl->s3_groups[s] = ng;
    }
    pl->tiles_per_image = tiles;
    pl->n_tiles = (long long)tiles * hd.B;
#endif
    pl->stream = 1;
}

#ifdef VY_STREAM_ALT
// ring depth and dynamic shared memory of the tile-streaming kernel: the ring gets what the two table buffers leave of
// the budget (VY_S2_SMEM_KB, default 150 KB: a finalize / sample CTA of a neighbouring stream still fits on the SM)
static int stream2_stages(const SelPlan &pl, size_t *dyn_bytes) {
    static const int budget_kb = getenv("VY_S2_SMEM_KB") ? atoi(getenv("VY_S2_SMEM_KB")) : 0;
    static size_t static_bytes = 0;
    if (!static_bytes) {
        cudaFuncAttributes fa;
        static_bytes = cudaFuncGetAttributes(&fa, (const void *)vy_decode_stream2_kernel) == cudaSuccess ? fa.sharedSizeBytes : 48 << 10;
    }
    const size_t tabs = 2 * (size_t)pl.tab_max * sizeof(float);
    size_t budget = (size_t)(227 << 10) - static_bytes - 1024;         // what a CTA may have beside the static arrays
    if (budget_kb >= 48 && ((size_t)budget_kb << 10) < budget) budget = (size_t)budget_kb << 10;
    long long n = budget > tabs ? (long long)((budget - tabs) / S2_STAGE_BYTES) : 0;
    if (n > S2_MAX_STAGES) n = S2_MAX_STAGES;
    if (n < 2) n = 2;
    if (dyn_bytes) *dyn_bytes = (size_t)n * S2_STAGE_BYTES + tabs;
    return (int)n;
}
#endif

// keys per image in the streamed list: 4x the expected K*S, never more than the image has rows
static int stream_list_cap(const VyHeads &hd, const SelPlan &pl) {
    long long cap = 4LL * pl.K * pl.samp_stride;
    if (cap < 4096) cap = 4096;
    if (cap > hd.R) cap = hd.R;
    return (int)cap;
}

static int plan_rows(int B, long long R, int topk, float valid_thresh, SelPlan *pl) {
    memset(pl, 0, sizeof(*pl));
    long long K = topk < 0 ? R : (topk < R ? topk : R);
    if (K < 1 || K > SEL_KMAX) return VY_EUNSUPPORTED;
    pl->K = (int)K;
    pl->valid_thresh = valid_thresh;
    plan_jobs(B, (R + 15) / 16, pl);          // >= 16 iterations of SEL_NT rows per job
    long long rpj = (R + pl->G - 1) / pl->G;
    rpj = (rpj + SEL_NT - 1) / SEL_NT * SEL_NT;
    pl->rows_per_job = rpj;
    return VY_OK;
}

static int select_grid(int n_jobs) {
    const int resident = resident_ctas();
    return n_jobs < resident ? n_jobs : resident;
}

template <int SRC>
static int launch_finalize(const VyHeads &hd, const RowParams &rp, const SelPlan &pl, const SelGlobal &g,
                           FinParams fp, int B, cudaStream_t st) {
    const size_t dyn = fin_dyn_smem(pl.K, &fp.lcap);
    if (pl.K > FIN_NT_MAX) VY_FAIL(VY_EINVAL, "finalize: K exceeds the CTA size");      // a thread per rank / slot
    // a thread per rank / slot: 512 threads serve K <= 512 (the reference's nms_topk = 400).  Measured after the
    // shared-memory diet (53 KB per CTA): against 1024 threads the one-stream step is 3 % slower at COCO 608 x 64 and
    // faster everywhere else (one frame, 416^2 x 128, 320^2 x 256: two CTAs per SM), and independent batches on
    // several streams gain 1-9 % (a small finalize CTA leaves the neighbouring stream's streaming CTAs resident)
    const int fin_nt = pl.K <= FIN_NT_MAX / 2 ? FIN_NT_MAX / 2 : FIN_NT_MAX;
    if (fp.overlap_thresh > 0.0f && fp.overlap_thresh < 1e30f) {
        fp.thr_lo = fp.overlap_thresh * (1.0f - 9.5367431640625e-07f);      // 2^-20 (vy_nms_math.cuh)
        fp.thr_hi = fp.overlap_thresh * (1.0f + 9.5367431640625e-07f);
    } else {            // thr <= 0 (or absurd): always the exact division
        fp.thr_lo = -INFINITY;
        fp.thr_hi = INFINITY;
    }
    VY_CUDA_CHECK(vy_ensure_dyn_smem((const void *)vy_nms_finalize_kernel<SRC>, dyn));
    VY_KERNEL(VY_K_FINALIZE, st, (vy_launch(vy_nms_finalize_kernel<SRC>, dim3(B), dim3(fin_nt), dyn, st, true, hd, rp, pl, g, fp)));
    VY_LAUNCH_CHECK("vy_nms_finalize_kernel");
    return VY_OK;
}

// ------------------------------------------------------------------------------------------------
// plan handle of the fused path: everything that depends only on the shapes and arguments -- head description, job /
// tile / table plan, workspace layout -- is resolved once; a launch fills in the pointers and enqueues the kernels.
// ------------------------------------------------------------------------------------------------
struct vy_decode_nms_plan {
    VyHeads hd;                  // head pointers null; vec decided per launch from the real pointers
    SelPlan pl;
    int slist_cap;
    size_t ws_bytes, header;
    float overlap_thresh;
    int force_suppress, post_nms, topk;
};

static int plan_init(vy_decode_nms_plan *P, const int *H, const int *W, const float *stride, const float *anchors,
                     int n_scales, int B, int A, int C, int agnostic, float overlap_thresh, float valid_thresh, int topk,
                     int force_suppress, int post_nms) {
    const float *fake[VY_MAX_SCALES] = {nullptr, nullptr, nullptr, nullptr};
    int rc = vy_fill_heads(&P->hd, fake, H, W, stride, anchors, n_scales, B, A, C, agnostic);
    if (rc != VY_OK) return rc;
    if (post_nms < 1) VY_FAIL(VY_EINVAL, "vy_decode_nms: post_nms must be >= 1");
    rc = plan_heads(P->hd, topk, valid_thresh, &P->pl);
    if (rc != VY_OK) VY_FAIL(rc, "vy_decode_nms: min(topk,R)=%d outside [1,%d]; use vy_decode_f32 + vy_box_nms_f32",
                             topk, SEL_KMAX);
    P->slist_cap = P->pl.stream ? stream_list_cap(P->hd, P->pl) : 0;
    P->ws_bytes = sel_workspace_layout(B, P->pl.G, P->pl.list_cap, nullptr, nullptr, &P->header, P->pl.stream ? P->pl.Gs : 0,
                                       P->slist_cap);
    P->overlap_thresh = overlap_thresh; P->force_suppress = force_suppress; P->post_nms = post_nms; P->topk = topk;
    return VY_OK;
}

static int plan_launch(const vy_decode_nms_plan *P, const float *const *head, float *out, int32_t *kept_rows,
                       void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (!out) VY_FAIL(VY_EINVAL, "vy_decode_nms: out is null");
    if (!workspace || workspace_bytes < P->ws_bytes)
        VY_FAIL(VY_EWORKSPACE, "vy_decode_nms: workspace %zu < %zu bytes", workspace_bytes, P->ws_bytes);
    if (((uintptr_t)workspace & 255) != 0) VY_FAIL(VY_EALIGN, "workspace must be 256-byte aligned");
    VyHeads hd = P->hd;
    const SelPlan &pl = P->pl;
    const int B = hd.B;
    for (int s = 0; s < hd.n_scales; ++s) {
        if (!head || !head[s]) VY_FAIL(VY_EINVAL, "vy_decode_nms: head[%d] is null", s);
        if (((uintptr_t)head[s] & 3) != 0) VY_FAIL(VY_EALIGN, "head[%d] not 4-byte aligned", s);
        hd.sc[s].head = head[s];
        hd.sc[s].vec = (hd.sc[s].HW % 4 == 0 && (((uintptr_t)head[s]) & 15) == 0) ? 4 : 1;
    }
    SelGlobal g;
    sel_workspace_layout(B, pl.G, pl.list_cap, &g, workspace, nullptr, pl.stream ? pl.Gs : 0, P->slist_cap);
    if (!pl.stream) VY_CUDA_CHECK(cudaMemsetAsync(workspace, 0, P->header, st));     // (the sample kernel zeroes its images' state itself)
    if (pl.stream) {
        VY_KERNEL(VY_K_SAMPLE, st, (vy_decode_sample_kernel<<<B * pl.Gs, SAMP_NT, 0, st>>>(hd, pl, g)));
        VY_LAUNCH_CHECK("vy_decode_sample_kernel");
        long long ctas = (pl.n_units + STR_NT / 32 - 1) / (STR_NT / 32);
        const long long resident = (long long)STR_CTAS_PER_SM * vy_sm_count();
        if (ctas > resident) ctas = resident;
        // PDL on this launch only in the latency regime (grid below one wave): on a full machine the early
        // CTAs crowd the sample kernel's tail (measured: -4 % at COCO 608 x 64, +15 % at VOC 416 x 1)
        const size_t ring_bytes = (size_t)(STR_NT / 32) * STR_RING * 32 * sizeof(float4);
#ifdef VY_STREAM_ALT
        // development build: VY_STREAM_MODE=v2 (tile streaming) / v3 (segment streaming) select the alternates
        static const char *mode_env = getenv("VY_STREAM_MODE");
        static const int mode = !mode_env ? 0 : (!strcmp(mode_env, "v3") ? 3 : (!strcmp(mode_env, "v2") ? 2 : 0));
        if (mode == 3) {
            const size_t dyn = (size_t)S3_WARPS * S3_RING_BYTES + (size_t)pl.tab_max * sizeof(float);
            long long c3 = (long long)S3_CTAS_PER_SM * vy_sm_count();
            if (c3 > pl.n_tiles) c3 = pl.n_tiles;
            VY_CUDA_CHECK(vy_ensure_dyn_smem((const void *)vy_decode_stream3_kernel, dyn));
            VY_KERNEL(VY_K_STREAM, st, (vy_launch(vy_decode_stream3_kernel, dim3((unsigned)c3), dim3(S3_NT), dyn, st, true, hd, pl, g)));
            VY_LAUNCH_CHECK("vy_decode_stream3_kernel");
        } else if (mode == 2) {
            size_t dyn = 0;
            const int nst = stream2_stages(pl, &dyn);
            long long c2 = vy_sm_count();
            if (c2 > pl.n_tiles) c2 = pl.n_tiles;
            VY_CUDA_CHECK(vy_ensure_dyn_smem((const void *)vy_decode_stream2_kernel, dyn));
            VY_KERNEL(VY_K_STREAM, st, (vy_launch(vy_decode_stream2_kernel, dim3((unsigned)c2), dim3(S2_NT), dyn, st, true, hd, pl, g, nst)));
            VY_LAUNCH_CHECK("vy_decode_stream2_kernel");
        } else
#endif
        {
            VY_CUDA_CHECK(vy_ensure_dyn_smem((const void *)vy_decode_stream_kernel, ring_bytes));
            VY_KERNEL(VY_K_STREAM, st, (vy_launch(vy_decode_stream_kernel, dim3((unsigned)ctas), dim3(STR_NT), ring_bytes, st, ctas < resident, hd, pl, g)));
            VY_LAUNCH_CHECK("vy_decode_stream_kernel");
        }
    }
    if (!pl.stream) {        // small / class-agnostic inputs: the adaptive select IS the selection
        VY_KERNEL(VY_K_SELECT_HEADS, st, (vy_launch(vy_decode_select_kernel, dim3(select_grid(pl.n_jobs)), dim3(SEL_NT), 0, st, true, hd, pl, g)));
        VY_LAUNCH_CHECK("vy_decode_select_kernel");
    }
    FinParams fp;
    fp.K = pl.K; fp.post_rows = P->post_nms; fp.out_stride_rows = P->post_nms;
    fp.overlap_thresh = P->overlap_thresh; fp.force_suppress = P->force_suppress;
    fp.in_format = VY_FMT_CORNER; fp.out_format = VY_FMT_CORNER; fp.W = 6; fp.fill_rest = 1;
    fp.out = out; fp.kept_rows = kept_rows;
    static const bool force_rescue = getenv("VY_FORCE_RESCUE") != nullptr;
    fp.force_rescue = force_rescue ? 1 : 0;
    RowParams rp;
    memset(&rp, 0, sizeof(rp));
    int rc = launch_finalize<0>(hd, rp, pl, g, fp, B, st);
    // VY_DEBUG_LISTS=1: (debugging aid, synchronises) print the fill of the streamed candidate lists
    static const bool dbg = getenv("VY_DEBUG_LISTS") != nullptr;
    if (dbg && pl.stream && rc == VY_OK) {
        std::vector<int> cnt(B);
        std::vector<u64> thr(B), slots((size_t)B * pl.Gs);
        cudaStreamSynchronize(st);
        cudaMemcpy(cnt.data(), g.scount, sizeof(int) * B, cudaMemcpyDeviceToHost);
        cudaMemcpy(slots.data(), g.sslots, sizeof(u64) * slots.size(), cudaMemcpyDeviceToHost);
        for (int i = 0; i < B; ++i) {
            thr[i] = 0;
            for (int j = 0; j < pl.Gs; ++j) thr[i] = slots[(size_t)i * pl.Gs + j] > thr[i] ? slots[(size_t)i * pl.Gs + j] : thr[i];
        }
        long long sum = 0; int mn = 1 << 30, mx = 0, bad = 0;
        for (int i = 0; i < B; ++i) {
            sum += cnt[i]; mn = cnt[i] < mn ? cnt[i] : mn; mx = cnt[i] > mx ? cnt[i] : mx;
            bad += cnt[i] > g.slist_cap || (cnt[i] < pl.K && ~thr[i] != 0ull);
        }
        fprintf(stderr, "[vyolo] streamed lists: K=%d S=%d Gs=%d Ksq=%d cap=%d | fill mean %.0f min %d max %d | rescued %d of %d\n",
                pl.K, pl.samp_stride, pl.Gs, pl.Ksq, g.slist_cap, (double)sum / B, mn, mx, bad, B);
    }
    return rc;
}

extern "C" int vy_decode_nms_plan_create(const int *H, const int *W, const float *stride, const float *anchors,
                                         int n_scales, int B, int A, int C, int agnostic, float overlap_thresh,
                                         float valid_thresh, int topk, int force_suppress, int post_nms,
                                         vy_decode_nms_plan_t **plan) {
    if (!plan) VY_FAIL(VY_EINVAL, "vy_decode_nms_plan_create: plan is null");
    *plan = nullptr;
    vy_decode_nms_plan *P = (vy_decode_nms_plan *)calloc(1, sizeof(vy_decode_nms_plan));
    if (!P) VY_FAIL(VY_EINVAL, "vy_decode_nms_plan_create: out of host memory");
    const int rc = plan_init(P, H, W, stride, anchors, n_scales, B, A, C, agnostic, overlap_thresh, valid_thresh, topk,
                             force_suppress, post_nms);
    if (rc != VY_OK) { free(P); return rc; }
    *plan = P;
    return VY_OK;
}
extern "C" size_t vy_decode_nms_plan_workspace_bytes(const vy_decode_nms_plan_t *plan) { return plan ? plan->ws_bytes : 0; }
extern "C" int vy_decode_nms_plan_launch(const vy_decode_nms_plan_t *plan, const float *const *head, float *out,
                                         int32_t *kept_rows, void *workspace, size_t workspace_bytes, vy_stream_t stream) {
    if (!plan) VY_FAIL(VY_EINVAL, "vy_decode_nms_plan_launch: plan is null");
    return plan_launch(plan, head, out, kept_rows, workspace, workspace_bytes, (cudaStream_t)stream);
}
extern "C" void vy_decode_nms_plan_destroy(vy_decode_nms_plan_t *plan) { free(plan); }

extern "C" size_t vy_decode_nms_workspace_bytes(const int *H, const int *W, int n_scales, int B, int A,
                                                int C, int agnostic, int topk) {
    vy_decode_nms_plan P;
    float st[VY_MAX_SCALES] = {1, 1, 1, 1};
    float an[VY_MAX_SCALES * VY_MAX_ANCHORS * 2] = {0};
    if (plan_init(&P, H, W, st, an, n_scales, B, A, C, agnostic, 0.5f, 0.0f, topk, 0, 1) != VY_OK) return 0;
    return P.ws_bytes;
}

extern "C" int vy_decode_nms_f32(const float *const *head, const int *H, const int *W, const float *stride,
                                 const float *anchors, int n_scales, int B, int A, int C, int agnostic,
                                 float overlap_thresh, float valid_thresh, int topk, int force_suppress,
                                 int post_nms, float *out, int32_t *kept_rows, void *workspace,
                                 size_t workspace_bytes, vy_stream_t stream) {
    vy_decode_nms_plan P;
    const int rc = plan_init(&P, H, W, stride, anchors, n_scales, B, A, C, agnostic, overlap_thresh, valid_thresh, topk,
                             force_suppress, post_nms);
    if (rc != VY_OK) return rc;
    return plan_launch(&P, head, out, kept_rows, workspace, workspace_bytes, (cudaStream_t)stream);
}

// implemented in vy_nms_large.cu: topk < 0 or min(topk,R) > SEL_KMAX
size_t vy_box_nms_large_workspace_bytes(int B, long long R, int W_elem);
int vy_box_nms_large(const RowParams &rp, int B, long long K, float overlap_thresh, int force_suppress,
                     int in_format, int out_format, long long out_rows, float *out, int32_t *kept_rows,
                     void *workspace, size_t workspace_bytes, cudaStream_t st);

extern "C" size_t vy_box_nms_workspace_bytes(int B, long R, int W_elem, int topk) {
    SelPlan pl;
    if (B < 1 || R < 1) return 0;
    if (plan_rows(B, R, topk, 0.0f, &pl) == VY_OK)
        return sel_workspace_layout(B, pl.G, pl.list_cap, nullptr, nullptr, nullptr);
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
    const int force_flags = force_suppress;      // bit 0x100: exhaustive large-path kernel (tests)
    force_suppress = (force_suppress & 0xff) ? 1 : 0;
    RowParams rp;
    rp.data = data; rp.R = R; rp.W = W_elem; rp.coord_start = coord_start; rp.score_index = score_index;
    rp.id_index = id_index; rp.background_id = background_id; rp.valid_thresh = valid_thresh;
    SelPlan pl;
    if (plan_rows(B, R, topk, valid_thresh, &pl) != VY_OK) {
        const long long K = topk < 0 ? R : (topk < R ? topk : R);
        if (K < 1) {      // topk == 0: nothing takes part
            VY_KERNEL(VY_K_FILL, st, (vy_fill_kernel<<<vy_sm_count() * 4, 256, 0, st>>>(
                out, kept_rows, (size_t)B * out_rows * W_elem, (size_t)B * out_rows)));
            VY_LAUNCH_CHECK("vy_fill_kernel");
            return VY_OK;
        }
        return vy_box_nms_large(rp, B, K, overlap_thresh, force_suppress | (force_flags & 0x100), in_format, out_format, out_rows,
                                out, kept_rows, workspace, workspace_bytes, st);
    }
    SelGlobal g;
    size_t header = 0;
    const size_t need = sel_workspace_layout(B, pl.G, pl.list_cap, &g, workspace, &header);
    if (!workspace || workspace_bytes < need)
        VY_FAIL(VY_EWORKSPACE, "vy_box_nms_f32: workspace %zu < %zu bytes", workspace_bytes, need);
    if (((uintptr_t)workspace & 255) != 0) VY_FAIL(VY_EALIGN, "workspace must be 256-byte aligned");
    VY_CUDA_CHECK(cudaMemsetAsync(workspace, 0, header, st));
    VY_KERNEL(VY_K_SELECT_ROWS, st, (vy_rows_select_kernel<<<select_grid(pl.n_jobs), SEL_NT, 0, st>>>(rp, pl, g)));
    VY_LAUNCH_CHECK("vy_rows_select_kernel");
    FinParams fp;
    fp.force_rescue = 0;
    fp.K = pl.K; fp.post_rows = (int)(out_rows < pl.K ? out_rows : pl.K); fp.out_stride_rows = out_rows;
    fp.overlap_thresh = overlap_thresh; fp.force_suppress = force_suppress;
    fp.in_format = in_format; fp.out_format = out_format; fp.W = W_elem;
    fp.out = out; fp.kept_rows = kept_rows;
    if (out_rows > pl.K) {
        // at most K rows can survive: pad everything first, survivors overwrite the front
        VY_KERNEL(VY_K_FILL, st, (vy_fill_kernel<<<vy_sm_count() * 8, 256, 0, st>>>(
            out, kept_rows, (size_t)B * out_rows * W_elem, (size_t)B * out_rows)));
        VY_LAUNCH_CHECK("vy_fill_kernel");
        fp.fill_rest = 0;
    } else {
        fp.fill_rest = 1;
    }
    VyHeads hd;
    memset(&hd, 0, sizeof(hd));
    return launch_finalize<1>(hd, rp, pl, g, fp, B, st);
}
