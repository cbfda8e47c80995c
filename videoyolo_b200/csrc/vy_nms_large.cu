// vy_nms_large.cu -- box_nms when more than SEL_KMAX candidates take part (topk < 0 or topk > 1024):
// MXNet's own default (topk = -1) and BASELINE config 4 (80 classes, valid_thresh 0.001, topk -1,
// force_suppress on/off, up to 851 760 participating rows per image).
//
// The small path (vy_nms.cu) keeps K <= 1024 candidates of an image in one CTA's shared memory.  Here
// every valid row can take part, so the work is organised around global-memory lists:
//
//   1. lg_key_kernel      key = (image, orderable score) per row, value = row; counts valid rows
//   2. radix sort #1      stable, descending  -> per image: rows in MXNet's order (score desc, ties by
//                         ascending source row), rank = position in that order
//   3. lg_key2_kernel     (class-aware only) key = (image, takes-part, class id), value = rank
//      radix sort #2      stable, ascending   -> every (image, class) is one contiguous segment that
//                         is still in rank order;  lg_seg_kernel lists the segment heads
//   4. lg_nms_kernel      one CTA per segment (per image when force_suppress / no ids): greedy
//                         suppression in tiles of 1024 candidates.  A tile is first tested against the
//                         boxes KEPT by earlier tiles (only survivors can suppress, so the cost is
//                         candidates x survivors, not candidates^2), then resolved internally with a
//                         1024 x 1024 suppression bitmask built by ballots and one warp's greedy scan;
//                         its survivors are appended to the segment's kept list.
//   5. lg_count / lg_write  survivors compacted to the front of the image in rank order.
//
// The two sorts are cub::DeviceRadixSort (CUDA toolkit header library, stable LSD radix sort): a
// commodity step, not the product; everything that defines the operator's result is in this file.
// IoU arithmetic: vy_nms_math.cuh (bit-exact with the oracle).
#include "vy_select.cuh"
#include "vy_nms_math.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <string.h>

namespace {

#ifndef LG_SH
#define LG_SH 1                   // tiled kernel: cells of 2^(L - LG_SH); measured at configs[3] force_suppress: 0 -> 336 ms, 1 -> 332 ms, 2 -> 490 ms, 3 -> 1303 ms (the lockstep cell loop pays per probed cell)
#endif
constexpr int LG_NT = 1024;          // threads per NMS CTA
constexpr int LG_TILE = 1024;        // candidates per tile (== LG_NT)
constexpr int LG_WORDS = LG_TILE / 32;
constexpr int LG_CHUNK = 2048;       // ranks per compaction chunk
constexpr int LG_CNT = 256;          // threads per compaction CTA (8 flags each)
constexpr int LA_NT = 256;           // threads per CTA of the adjacency kernel
#ifndef LA_SH
#define LA_SH 1                   // adjacency kernel: cells of 2^(L - LA_SH); measured at configs[3] class-aware: 0 -> 289 ms, 1 -> 239 ms, 2 -> 270 ms, 3 -> 494 ms
#endif
#ifndef LA_CTAS
#define LA_CTAS 8                  // CTAs per SM (= 32 registers per thread): the kernel is latency-bound, measured at configs[3]
#endif                           // class-aware: 4 -> 231 ms, 5 -> 217 ms, 6 -> 199 ms, 8 -> 183 ms; the grid is exactly the resident CTAs
constexpr int LA_E = 4;              // edge slots per candidate (average over a segment)

struct LgLayout {
    size_t hdr, keys_a, keys_b, vals_a, vals_b, vals_c, kept_box, kept_area, keep, seg, csum, cub, edges, pending, status, sbox, sap, total;
    size_t hdr_bytes, cub_bytes;
};

static size_t lg_align(size_t v) { return (v + 255) / 256 * 256; }

static int lg_bits(int B) { int b = 0; while ((1LL << b) < B) ++b; return b; }

// temp storage of the two sorts (the larger one); 0 on failure
static size_t lg_cub_bytes(long long N, int bitsB) {
    size_t t1 = 0, t2 = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(nullptr, t1, (const u64 *)nullptr, (u64 *)nullptr,
                                                               (const u32 *)nullptr, (u32 *)nullptr, N, 0, 32 + bitsB);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    e = cub::DeviceRadixSort::SortPairs(nullptr, t2, (const u64 *)nullptr, (u64 *)nullptr, (const u32 *)nullptr,
                                        (u32 *)nullptr, N, 0, 33 + bitsB);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return (t1 > t2 ? t1 : t2) + 256;
}

static bool lg_layout(int B, long long R, LgLayout *L) {
    const long long N = (long long)B * R;
    const int nchunk = (int)((R + LG_CHUNK - 1) / LG_CHUNK);
    size_t off = 0;
    L->hdr = off;       L->hdr_bytes = lg_align(sizeof(int) * ((size_t)B + 8)); off += L->hdr_bytes;   // nvalid[B], n_seg
    L->keys_a = off;    off = lg_align(off + sizeof(u64) * (size_t)N);
    L->keys_b = off;    off = lg_align(off + sizeof(u64) * (size_t)N);
    L->vals_a = off;    off = lg_align(off + sizeof(u32) * (size_t)N);
    L->vals_b = off;    off = lg_align(off + sizeof(u32) * (size_t)N);
    L->vals_c = off;    off = lg_align(off + sizeof(u32) * (size_t)N);
    L->kept_box = off;  off = lg_align(off + sizeof(float4) * (size_t)N);
    L->kept_area = off; off = lg_align(off + sizeof(float) * (size_t)N);
    L->keep = off;      off = lg_align(off + (size_t)N);
    L->seg = off;       off = lg_align(off + sizeof(u32) * (size_t)N);
    L->csum = off;      off = lg_align(off + sizeof(int) * (size_t)B * nchunk);
    L->cub_bytes = lg_cub_bytes(N, lg_bits(B));
    if (L->cub_bytes == 0) return false;
    L->cub = off;       off = lg_align(off + L->cub_bytes);
    L->edges = off;     off = lg_align(off + sizeof(u64) * (size_t)LA_E * (size_t)N);       // class-aware adjacency path
    L->pending = off;   off = lg_align(off + sizeof(int) * (size_t)N);
    L->status = off;    off = lg_align(off + (size_t)N);
    L->sbox = off;      off = lg_align(off + sizeof(float4) * (size_t)N);
    L->sap = off;       off = lg_align(off + sizeof(uint2) * (size_t)N);
    L->total = off;
    return true;
}

// ------------------------------------------------------------------------------------------------
// 1. keys of sort #1.  Descending order of ((B-1-b) << 32 | ord(score)) = images ascending, scores
//    descending; invalid rows carry ord 0 (no valid score maps to 0) and sink to the image's end.
// ------------------------------------------------------------------------------------------------
__global__ void lg_key_kernel(RowParams rp, int B, u64 *keys, u32 *vals, int *nvalid) {
    const long long N = (long long)B * rp.R;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long Nr = (N + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < Nr; i += stride) {
        const bool in = i < N;
        const int b = in ? (int)(i / rp.R) : -1;
        bool valid = false;
        float s = 0.0f;
        if (in) {
            const float *row = rp.data + (size_t)i * rp.W;
            s = vy_ldg32(row + rp.score_index);
            valid = s > rp.valid_thresh;                                       // strict; NaN fails
            if (valid && rp.id_index >= 0 && rp.background_id >= 0 && (int)row[rp.id_index] == rp.background_id)
                valid = false;
            keys[i] = ((u64)(u32)(B - 1 - b) << 32) | (u64)(valid ? vy_f2ord(s) : 0u);
            vals[i] = (u32)(i - (long long)b * rp.R);
        }
        const u32 peers = __match_any_sync(0xffffffffu, valid ? b : -1);
        if (valid && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(nvalid + b, __popc(peers));
    }
}

// ------------------------------------------------------------------------------------------------
// 3. keys of sort #2: (image, does-not-take-part, class id + 2^31), value = rank
// ------------------------------------------------------------------------------------------------
__global__ void lg_key2_kernel(RowParams rp, int B, long long K, const u32 *row_of, const int *nvalid,
                               u64 *keys2, u32 *vals2) {
    const long long N = (long long)B * rp.R;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const int b = (int)(i / rp.R);
        const long long rank = i - (long long)b * rp.R;
        const long long nb = min((long long)nvalid[b], K);
        const bool part = rank < nb;
        int cls = 0;
        if (part) cls = (int)rp.data[((size_t)b * (size_t)rp.R + row_of[i]) * rp.W + rp.id_index];
        keys2[i] = ((u64)(u32)b << 33) | ((u64)(part ? 0u : 1u) << 32) | (u64)((u32)cls ^ 0x80000000u);
        vals2[i] = (u32)rank;
    }
}

__global__ void lg_seg_kernel(const u64 *keys2, long long N, u32 *seg_starts, int *n_seg) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const u64 k = keys2[i];
        if ((k >> 32) & 1ull) continue;
        if (i == 0 || keys2[i - 1] != k) seg_starts[atomicAdd(n_seg, 1)] = (u32)i;
    }
}

// ------------------------------------------------------------------------------------------------
// 4. tiled greedy suppression
// ------------------------------------------------------------------------------------------------
struct LgNms {
    RowParams rp;
    int B;
    long long K;
    float thr, thr_lo, thr_hi;
    int all_pairs;               // one segment per image (force_suppress or no ids)
    const u32 *row_of;           // [N] source row by (image, rank)            (values of sort #1)
    const u64 *keys2;            // [N] sorted (image, part, class)            (class-aware only)
    const u32 *rank_of;          // [N] rank by segment position               (values of sort #2)
    const u32 *seg_starts;
    const int *n_seg;
    const int *nvalid;
    float4 *kept_box;            // [N] kept boxes of a segment, from the segment's first position
    float *kept_area;
    unsigned char *keep;         // [N] by (image, rank)
    u32 *hash_head;              // [2N] spatial hash of the kept boxes, a power-of-two region per segment (GRID)
    u32 *next;                   // [N]  chain links of the kept boxes                                    (GRID)
    // adjacency path (class-aware, lg_adj_kernel)
    u32 *seg_starts_rw;          // == seg_starts; bit 31 of an entry = "this segment goes to the tiled kernel"
    u64 *edges;                  // [LA_E * N] (suppressor position << 32 | suppressed position), a region per segment
    int *pending;                // [N] by position: undetermined suppressors (bit 30: one of them is kept)
    unsigned char *status;       // [N] by position: 0 undetermined, 1 kept, 2 suppressed
    float4 *sbox; uint2 *sap;                   // [N] the regular boxes of a segment sorted by cell: box, (area, position)
    int only_flagged;            // lg_nms_kernel: serve only the flagged segments
};

// ---- spatial index over the kept boxes (GRID variant) ------------------------------------------------------
// A box with IoU > t against a kept box k must overlap it (the fp32 predicate's intersection is positive only if
// the real intervals overlap: the sign of an fp32 difference is exact) and, from iw*ih > t*w_k*h_k with ih <= h_k
// and iw <= w_c, has w_k < w_c/t and w_c < w_k/t (same for h): sizes within a factor 1/t per axis.  Kept boxes
// are therefore registered under (level, cell of the centre), level L = the power of two with
// 2^(L-1) < max(w,h) <= 2^L and cells of size 2^L, and a candidate only probes the levels its size allows and the
// cells its extent (+ one cell of the level, the largest a registered box can be) reaches.  Every bound carries
// a 1e-3 relative slack, orders of magnitude above the 2^-24 roundings of the quantities involved, and the test
// applied to whatever the probes find is the exact predicate, so the result is the exhaustive kernel's.
// Boxes that cannot interact (w <= 0, h <= 0 or NaN: zero intersection with everything) are "inert"; boxes the
// index cannot place (non-finite or absurd magnitudes) are "irregular": kept ones sit on a chain every candidate
// walks, irregular candidates scan the whole kept list.
constexpr u32 LG_EMPTY = 0xffffffffu;
struct LgGeom { float cx, cy, w, h; int kind; };      // kind 0 regular, 1 inert, 2 irregular
template <int FMT>
__device__ __forceinline__ LgGeom lg_geom(float4 b) {
    LgGeom g;
    if (FMT == VY_FMT_CORNER) {
        g.w = __fsub_rn(b.z, b.x); g.h = __fsub_rn(b.w, b.y);
        g.cx = 0.5f * __fadd_rn(b.x, b.z); g.cy = 0.5f * __fadd_rn(b.y, b.w);
    } else { g.cx = b.x; g.cy = b.y; g.w = b.z; g.h = b.w; }
    if (!(g.w > 0.0f) || !(g.h > 0.0f)) { g.kind = 1; return g; }
    const float m = fmaxf(g.w, g.h), mn = fminf(g.w, g.h);
    const bool ok = m < 1.0e15f && mn > 1.0e-15f && fabsf(g.cx) < 1.0e15f && fabsf(g.cy) < 1.0e15f;   // NaN fails
    g.kind = ok ? 0 : 2;
    return g;
}
// level of a positive normal float: 2^(L-1) <= m < 2^L
__device__ __forceinline__ int lg_level(float m) { return (int)((__float_as_uint(m) >> 23) & 0xffu) - 126; }
__device__ __forceinline__ float lg_pow2(int e) { return __uint_as_float((u32)(e + 127) << 23); }      // e in [-126, 127]
__device__ __forceinline__ u32 lg_hash(int L, int ix, int iy) {
    u32 h = (u32)ix * 0x9E3779B1u ^ (u32)iy * 0x85EBCA77u ^ (u32)L * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}

template <int FMT, bool GRID>
__global__ void __launch_bounds__(LG_NT, 1) lg_nms_kernel(const __grid_constant__ LgNms p) {
    extern __shared__ __align__(16) unsigned char dyn[];
    float4 *tb = (float4 *)dyn;                     // [LG_TILE] live boxes of the tile, in order
    float4 *sb = tb + LG_TILE;                      // [LG_TILE] staged chunk of the kept list
    float *ta = (float *)(sb + LG_TILE);            // [LG_TILE]
    float *sa = ta + LG_TILE;                       // [LG_TILE]
    u32 *trank = (u32 *)(sa + LG_TILE);             // [LG_TILE]
    u32 *mask = trank + LG_TILE;                    // [LG_TILE][LG_WORDS]
    __shared__ u32 rowany[LG_WORDS], keepw[LG_WORDS];
    __shared__ int wcount[LG_WORDS], woff[LG_WORDS + 1], kpre[LG_WORDS + 1];
    __shared__ long long sh_len;
    __shared__ u32 sh_irr;                           // head of the segment's chain of irregular kept boxes
    __shared__ int sh_next;                          // next candidate of the tile to be probed (GRID)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    const long long R = p.rp.R, N = (long long)p.B * R;
    const int n_seg = p.all_pairs ? p.B : *p.n_seg;
    const float reach = fmaxf(1.0f - p.thr, 0.5f / p.thr - 0.5f) * 1.01f;       // (GRID: thr >= 0.05)
    const int sh = p.thr >= 0.3f ? LG_SH : (p.thr >= 0.15f ? 1 : 0);

    for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
        long long seg0, len = 0;
        u64 segkey = 0;
        int b;
        if (p.all_pairs) { b = seg; seg0 = (long long)b * R; len = min((long long)p.nvalid[b], p.K); }
        else {
            const u32 s0 = p.seg_starts[seg];
            if (p.only_flagged && !(s0 >> 31)) continue;     // (CTA-uniform) served by the adjacency kernel
            seg0 = s0 & 0x7fffffffu; segkey = p.keys2[seg0]; b = (int)(segkey >> 33);
        }
        const size_t img = (size_t)b * (size_t)R;
        u32 hmask = 0;
        u32 *hhead = nullptr, *htail = nullptr;
        if (GRID) {
            // the segment's hash region: [2*seg0, 2*seg0 + hsize), hsize the largest power of two <= 2*len
            if (!p.all_pairs) {
                if (tid == 0) {                     // end of the segment: first position with another key (sorted)
                    long long lo = seg0 + 1, hi = N;
                    while (lo < hi) {
                        const long long mid = (lo + hi) >> 1;
                        if (p.keys2[mid] == segkey) lo = mid + 1; else hi = mid;
                    }
                    sh_len = lo - seg0;
                }
                __syncthreads();
                len = sh_len;
            }
            if (tid == 0) sh_irr = LG_EMPTY;
            // hsize = the largest power of two <= len: heads in the first half of the region, TAILS in the second (a chain is
            // walked oldest = highest-ranked kept box first: that is the likely suppressor, and a candidate stops at the
            // first one it finds; with the newest box at the head a candidate walked most of a chain before it met its
            // suppressor -- 2.8 k warp instructions per candidate, most of them here)
            u32 hsize = 1;
            while ((long long)hsize * 2 <= len) hsize *= 2;
            hmask = hsize - 1;
            hhead = p.hash_head + 2 * (size_t)seg0;
            htail = hhead + hsize;
            __syncthreads();
        }
        int m = 0;                                  // boxes kept so far (CTA-uniform)
        for (long long t0 = 0;; t0 += LG_TILE) {
            // ---- this thread's candidate
            bool have;
            u32 rank = 0;
            if (p.all_pairs) { have = t0 + tid < len; rank = (u32)(t0 + tid); }
            else {
                const long long gi = seg0 + t0 + tid;
                have = gi < N && p.keys2[gi] == segkey;
                if (have) rank = p.rank_of[gi];
            }
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            float ar = 0.0f;
            if (have) {
                const float *q = p.rp.data + (img + p.row_of[img + rank]) * p.rp.W + p.rp.coord_start;
                bx = make_float4(q[0], q[1], q[2], q[3]);
                ar = nms_area(bx, FMT);
            }
            const int nv = __syncthreads_count(have);
            if (nv == 0) break;
            bool alive = have;
            // ---- a. against the boxes kept by earlier tiles (a suppressed box never suppresses)
            for (int c0 = 0; !GRID && c0 < m; c0 += LG_TILE) {
                const int cnt = min(LG_TILE, m - c0);
                __syncthreads();
                if (tid < cnt) { sb[tid] = p.kept_box[seg0 + c0 + tid]; sa[tid] = p.kept_area[seg0 + c0 + tid]; }
                __syncthreads();
                if (alive) {
                    int j = 0;
                    for (; j + 4 <= cnt; j += 4) {
                        bool h = nms_suppresses_fast(sb[j], sa[j], bx, ar, p.thr, p.thr_lo, p.thr_hi, FMT);
                        h |= nms_suppresses_fast(sb[j + 1], sa[j + 1], bx, ar, p.thr, p.thr_lo, p.thr_hi, FMT);
                        h |= nms_suppresses_fast(sb[j + 2], sa[j + 2], bx, ar, p.thr, p.thr_lo, p.thr_hi, FMT);
                        h |= nms_suppresses_fast(sb[j + 3], sa[j + 3], bx, ar, p.thr, p.thr_lo, p.thr_hi, FMT);
                        if (h) { alive = false; break; }
                    }
                    if (alive)
                        for (; j < cnt; ++j)
                            if (nms_suppresses_fast(sb[j], sa[j], bx, ar, p.thr, p.thr_lo, p.thr_hi, FMT)) { alive = false; break; }
                }
            }
            if (GRID && m > 0) {                         // CTA-uniform
                // One candidate at a time per warp, the 32 lanes side by side on ITS probes: every lane takes a cell of
                // the candidate's windows and walks that cell's chain (oldest = highest-ranked kept box first), and the
                // warp stops at the first suppressor any lane finds.  (With a candidate per lane the warp ran to the
                // longest of its 32 walks -- a chain of dependent L2 round trips -- although 99 % of the candidates
                // of a force_suppress tile are suppressed by one of the first kept boxes they meet.)  The warps take
                // the tile's candidates from a shared counter: a survivor walks all its chains, and a warp that
                // draws several of them would keep the other 31 waiting at the barrier.
                __syncthreads();
                sb[tid] = bx; sa[tid] = ar;                                            // (the staging buffers of the other path)
                trank[tid] = have ? 1u : 0u;                                           // alive flags (trank is filled below)
                if (tid == 0) sh_next = 0;
                __syncthreads();
                for (;;) {
                    int c = 0;
                    u32 alive_c = 0u;
                    if (lane == 0) { c = atomicAdd(&sh_next, 1); alive_c = c < LG_TILE ? trank[c] : 0u; }   // only lane 0 touches trank[c] in this loop
                    c = __shfl_sync(0xffffffffu, c, 0);
                    alive_c = __shfl_sync(0xffffffffu, alive_c, 0);
                    if (c >= LG_TILE) break;
                    if (!alive_c) continue;                                            // (warp-uniform)
                    const float4 cb = sb[c];
                    const float ca = sa[c];
                    const LgGeom gc = lg_geom<FMT>(cb);
                    if (gc.kind == 1) continue;                                        // inert: cannot be suppressed
                    bool dead = false;
                    bool scan_all = gc.kind == 2;                                      // irregular candidate: the whole kept list
                    int my_x0 = 0, my_y0 = 0, my_nx = 1, my_n = 0, l_lo = 0;
                    if (!scan_all) {
                        const float mc = fmaxf(gc.w, gc.h);
                        // sizes a suppressor can have: (t*mc, mc/t), with slack; thr >= 0.05 on this path
                        l_lo = lg_level(p.thr * mc * 0.999f);
                        const int nlev = lg_level(mc / p.thr * 1.001f) - l_lo + 1;
                        bool bad = false;
                        if (lane < nlev && nlev <= 8) {                                // lane i: the window at level l_lo + i
                            const int L = l_lo + lane;
                            // centres of two boxes with IoU > t are closer than reach * (own extent) per axis (derivation
                            // at lg_adj_kernel); kept boxes are registered by centre in cells of 2^(L - sh)
                            const float cs = lg_pow2(L - sh), inv = lg_pow2(sh - L);
                            const float rx = gc.w * reach, ry = gc.h * reach;
                            const float mx = rx + cs * 1e-3f + (fabsf(gc.cx) + rx) * 1e-6f;
                            const float my = ry + cs * 1e-3f + (fabsf(gc.cy) + ry) * 1e-6f;
                            const float fx0 = floorf((gc.cx - mx) * inv), fx1 = floorf((gc.cx + mx) * inv);
                            const float fy0 = floorf((gc.cy - my) * inv), fy1 = floorf((gc.cy + my) * inv);
                            if (!(fabsf(fx0) < 1.0e9f && fabsf(fx1) < 1.0e9f && fabsf(fy0) < 1.0e9f && fabsf(fy1) < 1.0e9f) ||
                                fx1 - fx0 > 31.0f || fy1 - fy0 > 31.0f) bad = true;
                            else {
                                my_x0 = (int)fx0; my_y0 = (int)fy0; my_nx = (int)fx1 - (int)fx0 + 1;
                                my_n = my_nx * ((int)fy1 - (int)fy0 + 1);
                            }
                        }
                        scan_all = nlev > 8 || __any_sync(0xffffffffu, bad);
                    }
                    if (scan_all) {                                                    // awkward geometry: every kept box, 32 at a time
                        for (int j0 = 0; j0 < m && !dead; j0 += 32) {
                            const int j = j0 + lane;
                            const bool hit = j < m && nms_suppresses_fast(p.kept_box[seg0 + j], p.kept_area[seg0 + j], cb, ca,
                                                                          p.thr, p.thr_lo, p.thr_hi, FMT);
                            dead = __any_sync(0xffffffffu, hit);
                        }
                    } else {
                        int my_end = my_n;                                             // inclusive prefix over the levels (lanes 0..7)
#pragma unroll
                        for (int off = 1; off < 8; off <<= 1) {
                            const int v = __shfl_up_sync(0xffffffffu, my_end, off);
                            if (lane >= off) my_end += v;
                        }
                        const int total = __shfl_sync(0xffffffffu, my_end, 7);
                        // cell `total` stands for the chain of irregular kept boxes: everyone checks them
                        for (int q0 = 0; q0 <= total && !dead; q0 += 32) {
                            const int q = q0 + lane;
                            int i = 0;
#pragma unroll
                            for (int j = 0; j < 7; ++j) {
                                const int e = __shfl_sync(0xffffffffu, my_end, j);
                                if (q >= e) i = j + 1;
                            }
                            const int lv_end = __shfl_sync(0xffffffffu, my_end, i), lv_n = __shfl_sync(0xffffffffu, my_n, i);
                            const int nx = __shfl_sync(0xffffffffu, my_nx, i), x0 = __shfl_sync(0xffffffffu, my_x0, i);
                            const int y0 = __shfl_sync(0xffffffffu, my_y0, i);
                            u32 cur = LG_EMPTY;
                            if (q < total) {
                                const int qq = q - (lv_end - lv_n);
                                const int iy = qq / nx, ix = qq - iy * nx;
                                cur = hhead[lg_hash(l_lo + i, x0 + ix, y0 + iy) & hmask];
                            } else if (q == total) {
                                cur = sh_irr;
                            }
                            for (int budget = m + 1; budget > 0; --budget) {           // (a corrupted table cannot hang the GPU)
                                bool hit = false;
                                if (cur != LG_EMPTY) {
                                    const u32 nx_ = p.next[cur];
                                    hit = nms_suppresses_fast(p.kept_box[cur], p.kept_area[cur], cb, ca, p.thr, p.thr_lo, p.thr_hi, FMT);
                                    cur = nx_;
                                }
                                if (__any_sync(0xffffffffu, hit)) { dead = true; break; }
                                if (!__any_sync(0xffffffffu, cur != LG_EMPTY)) break;
                            }
                        }
                    }
                    if (dead && lane == 0) trank[c] = 0u;
                }
                __syncthreads();
                alive = trank[tid] != 0u;
                __syncthreads();
            }
            // ---- b. the tile's live candidates, compacted in order
            const u32 bal = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) wcount[warp] = __popc(bal);
            if (tid < LG_WORDS) rowany[tid] = 0u;
            __syncthreads();
            if (warp == 0) {
                const int c = wcount[lane];
                int inc = c;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, off);
                    if (lane >= off) inc += v;
                }
                woff[lane + 1] = inc;
                if (lane == 0) woff[0] = 0;
            }
            __syncthreads();
            const int na = woff[LG_WORDS];
            if (alive) {
                const int idx = woff[warp] + __popc(bal & lt_mask);
                tb[idx] = bx; ta[idx] = ar; trank[idx] = rank;
            }
            __syncthreads();
            if (na > 0) {
                // suppression bitmask among the live candidates: one warp per reference, 32 later
                // candidates per ballot; every word from the diagonal on is written
                const int nw = (na + 31) >> 5;
                for (int i = warp; i < na; i += LG_NT / 32) {
                    const float4 bi = tb[i];
                    const float ai = ta[i];
                    u32 any = 0;
                    for (int w = i >> 5; w < nw; ++w) {
                        const int jx = (w << 5) + lane;
                        bool sup = false;
                        if (jx > i && jx < na) sup = nms_suppresses_fast(bi, ai, tb[jx], ta[jx], p.thr, p.thr_lo, p.thr_hi, FMT);
                        const u32 bits = __ballot_sync(0xffffffffu, sup);
                        if (lane == 0) mask[(size_t)i * LG_WORDS + w] = bits;
                        any |= bits;
                    }
                    if (lane == 0 && any) atomicOr(&rowany[i >> 5], 1u << (i & 31));
                }
                __syncthreads();
                if (warp == 0) {
                    nms_greedy_scan_warp(mask, LG_WORDS, rowany, na, keepw, lane);
                    __syncwarp();
                    const int c = lane < nw ? __popc(keepw[lane]) : 0;
                    int inc = c;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, inc, off);
                        if (lane >= off) inc += v;
                    }
                    kpre[lane] = inc - c;
                    if (lane == 31) kpre[LG_WORDS] = inc;
                }
                __syncthreads();
                if (tid < na) {
                    const u32 kw = keepw[tid >> 5];
                    if ((kw >> (tid & 31)) & 1u) {
                        const int pos = m + kpre[tid >> 5] + __popc(kw & lt_mask);
                        p.kept_box[seg0 + pos] = tb[tid];
                        p.kept_area[seg0 + pos] = ta[tid];
                        p.keep[img + trank[tid]] = 1;
                        if (GRID) {
                            const LgGeom gk = lg_geom<FMT>(tb[tid]);
                            const u32 idx = (u32)(seg0 + pos);
                            bool placed = false;
                            if (gk.kind == 0) {
                                const int L = lg_level(fmaxf(gk.w, gk.h));
                                const float inv = lg_pow2(sh - L);
                                const float fx = floorf(gk.cx * inv), fy = floorf(gk.cy * inv);
                                if (fabsf(fx) < 1.0e9f && fabsf(fy) < 1.0e9f) {
                                    // append at the tail (the exchanges on the tail word order concurrent appends; the
                                    // link of the previous tail is written by its successor)
                                    const u32 hb_ = lg_hash(L, (int)fx, (int)fy) & hmask;
                                    p.next[idx] = LG_EMPTY;
                                    __threadfence_block();
                                    const u32 prev = atomicExch(&htail[hb_], idx);
                                    if (prev == LG_EMPTY) hhead[hb_] = idx; else p.next[prev] = idx;
                                    placed = true;
                                }
                            }
                            if (!placed && gk.kind != 1) p.next[idx] = atomicExch(&sh_irr, idx);
                        }
                    }
                }
                m += kpre[LG_WORDS];
            }
            if (nv < LG_TILE) break;
            __syncthreads();
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// 4b. class-aware segments: static adjacency + wavefront resolution (lg_adj_kernel)
//
// Greedy suppression keeps candidate i iff no KEPT candidate of higher rank has IoU > t with it.  The tiled kernel above
// tests candidates against the survivors so far: candidates x survivors pair tests (3.9e7 per class of 10 647 random
// boxes, 69 % of which survive).  But which pairs CAN interact does not depend on the order: with every candidate of the
// segment registered in the spatial hash (same levels / cells / slack as above) the pairs with IoU > t are found by
// probing ~40 cells per candidate (~1e2 exact tests instead of ~4e3), as edges (suppressor -> suppressed, higher rank
// -> lower rank).  The keep decisions then follow from the edges alone, exactly:
//     kept(i)  <=>  every suppressor of i is suppressed;   suppressed(i)  <=>  some suppressor of i is kept
// resolved in rounds over the edge list: an edge is consumed once its source is decided; the consumer that takes a
// node's last open edge decides the node (one atomic word per node: open-edge count + "has a kept suppressor" bit).
// The lowest undecided position always has all its suppressors decided, so every round makes progress; random boxes
// need ~10-20 rounds.  One CTA per (image, class) segment, everything in rank order of the segment (= position).
// Segments the scheme does not fit -- irregular geometry, more than LA_E edges per candidate on average (thousands of
// near-identical boxes) -- are flagged and served by the tiled kernel afterwards.
// ------------------------------------------------------------------------------------------------
#ifdef LA_STATS
__device__ unsigned long long la_stats[8];     // visits, lower positions, area-compatible, edges, probes, rounds, segments
#define LA_COUNT(i, v) atomicAdd(&la_stats[i], (unsigned long long)(v))
#else
#define LA_COUNT(i, v) do { } while (0)
#endif
template <int FMT>
__global__ void __launch_bounds__(LA_NT, LA_CTAS) lg_adj_kernel(const __grid_constant__ LgNms p) {
    __shared__ long long sh_len;
    __shared__ int sh_flag, sh_ne, sh_done, sh_nreg;
    __shared__ u32 sh_scan[LA_NT];
    const int tid = threadIdx.x;
    const long long R = p.rp.R, N = (long long)p.B * R;
    const int n_seg = *p.n_seg;
    // How far apart can the centres of two boxes with IoU > t be?  inter > t * union >= t * max(A, A') and ih <= min(h, h')
    // give iw > t * max(w, w'), and iw <= (w + w') / 2 - |dcx|, so |dcx| < (w + w') / 2 - t * max(w, w'); over the widths a
    // suppressor can have, w' in (t w, w / t), that is at most reach * w with reach = max(1 - t, 1 / (2 t) - 1 / 2)
    // (0.61 at t = 0.45) -- a bound in terms of the candidate's own extent, whatever the partner's level.  (Real
    // arithmetic; the fp32 predicate's roundings are relative 2^-23, the probes carry a 1 % slack.)  Boxes are registered by
    // centre in cells of 2^(L - sh), finer where the reach is short, so a probe window is a few cells wide.
    const float reach = fmaxf(1.0f - p.thr, 0.5f / p.thr - 0.5f) * 1.01f;
    const int sh = p.thr >= 0.3f ? LA_SH : (p.thr >= 0.15f ? 1 : 0);
    const float atl = p.thr * 0.99f;
    for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
        const long long seg0 = p.seg_starts[seg];
        const u64 segkey = p.keys2[seg0];
        const int b = (int)(segkey >> 33);
        const size_t img = (size_t)b * (size_t)R;
        __syncthreads();
        if (tid == 0) {                             // end of the segment: first position with another key (sorted)
            long long lo = seg0 + 1, hi = N;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (p.keys2[mid] == segkey) lo = mid + 1; else hi = mid;
            }
            sh_len = lo - seg0;
            sh_flag = 0; sh_ne = 0; sh_done = 0;
        }
        __syncthreads();
        const long long len = sh_len;
        u32 hsize = 2;
        while ((long long)hsize * 2 <= 2 * len) hsize *= 2;
        const u32 hmask = hsize - 1;
        u32 *cell = p.hash_head + 2 * (size_t)seg0;                 // [hsize] members per cell, then where each cell ENDS in the sorted order
        float4 *box = p.kept_box + seg0;                             // by position
        float *area = p.kept_area + seg0;
        u32 *cell_of = p.next + seg0;                                // by position: the box's cell (LG_EMPTY: inert)
        float4 *sbox = p.sbox + seg0;                                // by slot: the regular boxes sorted by cell
        uint2 *sap = p.sap + seg0;                                   // by slot: (area, position)
        int *pending = p.pending + seg0;
        volatile unsigned char *status = p.status + seg0;
        u64 *edges = p.edges + (size_t)LA_E * (size_t)seg0;
        const long long cap = (long long)LA_E * len;
        const int cap_i = (int)(cap < 0x3fffffffLL ? cap : 0x3fffffffLL);       // edge slots of the segment (the counter is an int)
        // ---- 1. boxes by position; members per (level, cell) bucket
        for (u32 i = tid; i < hsize; i += LA_NT) cell[i] = 0u;
        __syncthreads();
        for (long long pos = tid; pos < len; pos += LA_NT) {
            const u32 rank = p.rank_of[seg0 + pos];
            const float *q = p.rp.data + (img + p.row_of[img + rank]) * p.rp.W + p.rp.coord_start;
            const float4 bx = make_float4(q[0], q[1], q[2], q[3]);
            box[pos] = bx;
            area[pos] = nms_area(bx, FMT);
            const LgGeom gk = lg_geom<FMT>(bx);
            u32 c = LG_EMPTY;
            bool placed = gk.kind == 1;             // inert boxes interact with nothing: kept
            if (gk.kind == 0) {
                const int L = lg_level(fmaxf(gk.w, gk.h));
                const float inv = lg_pow2(sh - L);                  // cells of 2^(L - sh): a fraction of the size class
                const float fx = floorf(gk.cx * inv), fy = floorf(gk.cy * inv);
                if (fabsf(fx) < 1.0e9f && fabsf(fy) < 1.0e9f) {
                    c = lg_hash(L, (int)fx, (int)fy) & hmask;
                    atomicAdd(&cell[c], 1u);
                    placed = true;
                }
            }
            cell_of[pos] = c;
            status[pos] = c == LG_EMPTY ? 1 : 0;
            pending[pos] = 0;
            if (!placed) sh_flag = 1;               // irregular geometry: the tiled kernel takes the segment
        }
        __syncthreads();
        if (sh_flag) { if (tid == 0) p.seg_starts_rw[seg] = (u32)seg0 | 0x80000000u; continue; }
        // ---- exclusive scan of the member counts (contiguous chunk per thread), then a counting sort by cell: members
        // of a cell are neighbours in the sorted order, so the lanes of a warp probe the same cells
        {
            const u32 per = (hsize + LA_NT - 1) / LA_NT;
            const u32 i0 = tid * per, i1 = min(hsize, i0 + per);
            u32 sum = 0;
            for (u32 i = i0; i < i1; ++i) sum += cell[i];
            sh_scan[tid] = sum;
            __syncthreads();
            if (tid == 0) {
                u32 run = 0;
                for (int t = 0; t < LA_NT; ++t) { const u32 v = sh_scan[t]; sh_scan[t] = run; run += v; }
                sh_nreg = (int)run;
            }
            __syncthreads();
            u32 run = sh_scan[tid];
            for (u32 i = i0; i < i1; ++i) { const u32 v = cell[i]; cell[i] = run; run += v; }
        }
        __syncthreads();
        // (chunks of LA_NT positions, one after the other: the members of a cell end up in ascending order of position up to
        // the order inside a chunk, so a scan for LOWER positions can stop at the first member of a later chunk)
        for (long long p0 = 0; p0 < len; p0 += LA_NT) {
            const long long pos = p0 + tid;
            const u32 c = pos < len ? cell_of[pos] : LG_EMPTY;
            if (c != LG_EMPTY) {
                const u32 slot = atomicAdd(&cell[c], 1u);            // cell[c] ends up at the END of cell c = the start of c + 1
                sbox[slot] = box[pos];
                sap[slot] = make_uint2(__float_as_uint(area[pos]), (u32)pos);
            }
            __syncthreads();
        }
        const int nreg = sh_nreg;
        // ---- 2. edges: every candidate looks for suppressors (lower positions) among the boxes its extent can reach.
        // One candidate per WARP (a thread per candidate walks its own loop nest and the warp runs a lane or two at a
        // time): the lanes take the cells of the probe windows side by side, trim each cell's member list to the lower
        // positions, and then the members of all 32 cells are tested 32 at a time (ranges flattened with a warp scan).
        {
            const int lane = tid & 31, warp = tid >> 5;
            for (int sl = warp; sl < nreg; sl += LA_NT / 32) {
                if (*(volatile int *)&sh_ne > cap_i) break;          // overflowed: the tiled kernel takes the segment
                const float4 bx = sbox[sl];
                const uint2 me = sap[sl];
                const float ar = __uint_as_float(me.x);
                const u32 pos = me.y;
                const u32 pos_end = (pos | (u32)(LA_NT - 1)) + 1u;   // first position of the next placement chunk
                const LgGeom gc = lg_geom<FMT>(bx);
                int found = 0;
                auto test = [&](int r, uint2 q) {
                    const u32 k = q.y;
                    LA_COUNT(0, 1);
                    if (k >= pos) return;
                    LA_COUNT(1, 1);
                    const float ak = __uint_as_float(q.x);
                    // IoU > t needs both areas within a factor t of each other (inter <= min, union >= max); 1 % slack,
                    // and only where the products cannot overflow or vanish
                    if (ak < 1e30f && ar < 1e30f && ak > 1e-30f && ar > 1e-30f && (ak < atl * ar || ar < atl * ak)) return;
                    LA_COUNT(2, 1);
                    if (nms_suppresses_fast(sbox[r], ak, bx, ar, p.thr, p.thr_lo, p.thr_hi, FMT)) {
                        LA_COUNT(3, 1);
                        // (once the region has overflowed nobody counts on: the counter stays within a CTA's worth of cap)
                        if (*(volatile int *)&sh_ne <= cap_i) {
                            const int at = atomicAdd(&sh_ne, 1);
                            if (at < cap_i) edges[at] = ((u64)k << 32) | (u64)pos;
                        }
                        ++found;
                    }
                };
                const float mc = fmaxf(gc.w, gc.h);
                // sizes a suppressor can have: (t*mc, mc/t), with slack; thr >= 0.05 on this path
                const int l_lo = lg_level(p.thr * mc * 0.999f), l_hi = lg_level(mc / p.thr * 1.001f);
                // the probe windows: lane i works out level l_lo + i (first cell, cells per row, cells); prefix over the levels
                const int nlev = l_hi - l_lo + 1;
                int my_x0 = 0, my_y0 = 0, my_nx = 1, my_n = 0;
                bool bad = false;
                if (lane < nlev && nlev <= 8) {
                    const int L = l_lo + lane;
                    const float cs = lg_pow2(L - sh), inv = lg_pow2(sh - L);
                    const float rx = gc.w * reach, ry = gc.h * reach;
                    const float mx = rx + cs * 1e-3f + (fabsf(gc.cx) + rx) * 1e-6f;
                    const float my = ry + cs * 1e-3f + (fabsf(gc.cy) + ry) * 1e-6f;
                    const float fx0 = floorf((gc.cx - mx) * inv), fx1 = floorf((gc.cx + mx) * inv);
                    const float fy0 = floorf((gc.cy - my) * inv), fy1 = floorf((gc.cy + my) * inv);
                    if (!(fabsf(fx0) < 1.0e9f && fabsf(fx1) < 1.0e9f && fabsf(fy0) < 1.0e9f && fabsf(fy1) < 1.0e9f) ||
                        fx1 - fx0 > 31.0f || fy1 - fy0 > 31.0f) bad = true;
                    else {
                        my_x0 = (int)fx0; my_y0 = (int)fy0; my_nx = (int)fx1 - (int)fx0 + 1;
                        my_n = my_nx * ((int)fy1 - (int)fy0 + 1);
                    }
                }
                const bool scan_all = nlev > 8 || __any_sync(0xffffffffu, bad);
                int my_end = my_n;                                   // inclusive prefix over the levels (lanes 0..7)
#pragma unroll
                for (int off = 1; off < 8; off <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, my_end, off);
                    if (lane >= off) my_end += v;
                }
                const int total = __shfl_sync(0xffffffffu, my_end, 7);
                if (scan_all) {                                      // (awkward geometry: every member, 32 at a time)
                    for (int r = lane; r < nreg; r += 32) test(r, sap[r]);
                } else {
                    for (int q0 = 0; q0 < total; q0 += 32) {
                        // this lane's cell of the round: member range, trimmed to the chunks at or below the candidate's
                        int r0 = 0, n = 0;
                        const int q = q0 + lane;
                        // level of cell q: the first level whose prefix exceeds q (<= 8 levels: a few shuffles, all lanes)
                        int i = 0;
#pragma unroll
                        for (int j = 0; j < 7; ++j) {
                            const int e = __shfl_sync(0xffffffffu, my_end, j);
                            if (q >= e) i = j + 1;
                        }
                        const int lv_end = __shfl_sync(0xffffffffu, my_end, i), lv_n = __shfl_sync(0xffffffffu, my_n, i);
                        const int nx = __shfl_sync(0xffffffffu, my_nx, i), x0 = __shfl_sync(0xffffffffu, my_x0, i), y0 = __shfl_sync(0xffffffffu, my_y0, i);
                        if (q < total) {
                            const int qq = q - (lv_end - lv_n);
                            const int iy = qq / nx, ix = qq - iy * nx;
                            const u32 c = lg_hash(l_lo + i, x0 + ix, y0 + iy) & hmask;
                            LA_COUNT(4, 1);
                            r0 = c ? (int)cell[c - 1] : 0;
                            int r1 = (int)cell[c];
                            if (r1 - r0 > 4) {                       // members ascend by chunk: cut at the first later chunk
                                int lo = r0, hi = r1;
                                while (lo < hi) {
                                    const int mid = (lo + hi) >> 1;
                                    if (sap[mid].y < pos_end) lo = mid + 1; else hi = mid;
                                }
                                r1 = lo;
                            }
                            n = r1 - r0;
                        }
                        const int nmax = __reduce_max_sync(0xffffffffu, n);
                        if (nmax == 0) continue;
                        // flatten: member t of the round belongs to the lane whose running offset covers it
                        int inc = n;
#pragma unroll
                        for (int off = 1; off < 32; off <<= 1) {
                            const int v = __shfl_up_sync(0xffffffffu, inc, off);
                            if (lane >= off) inc += v;
                        }
                        const int exc = inc - n;
                        const int T = __shfl_sync(0xffffffffu, inc, 31);
                        if (((T + 31) >> 5) * 3 >= nmax * 2) {
                            // short lists: lane by lane, member m of every cell at once (no owner search)
                            for (int m = 0; m < nmax; m += 2) {          // (two members per round: their loads overlap)
                                const uint2 q0 = m < n ? sap[r0 + m] : make_uint2(0u, 0xffffffffu);
                                const uint2 q1 = m + 1 < n ? sap[r0 + m + 1] : make_uint2(0u, 0xffffffffu);
                                test(r0 + m, q0);                         // (position 0xffffffff is never below the candidate's)
                                test(r0 + m + 1, q1);
                            }
                        } else {
                            for (int t0 = 0; t0 < T; t0 += 32) {
                                const int t = t0 + lane;
                                // owner = the last lane with exc <= t (binary search over the 32 offsets)
                                int own = 0;
#pragma unroll
                                for (int step = 16; step > 0; step >>= 1) {
                                    const int cand = own + step;
                                    const int e = __shfl_sync(0xffffffffu, exc, cand & 31);
                                    if (cand < 32 && e <= t) own = cand;
                                }
                                const int o_exc = __shfl_sync(0xffffffffu, exc, own);
                                const int o_r0 = __shfl_sync(0xffffffffu, r0, own);
                                if (t < T) {
                                    const int r = o_r0 + (t - o_exc);
                                    test(r, sap[r]);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) found += __shfl_xor_sync(0xffffffffu, found, off);
                if (lane == 0) { pending[pos] = found; if (found == 0) status[pos] = 1; }
            }
        }
        __syncthreads();
        const int ne = sh_ne;
        if (ne > cap_i) { if (tid == 0) p.seg_starts_rw[seg] = (u32)seg0 | 0x80000000u; continue; }
        // ---- 3. resolution: rounds over the open edges
        for (int done = 0; done < ne; ) {
            for (int e = tid; e < ne; e += LA_NT) {
                const u64 ed = edges[e];
                if (ed >> 63) continue;                                   // consumed
                const u32 src = (u32)(ed >> 32), dst = (u32)ed;
                const unsigned char sj = status[src];
                if (sj == 0) continue;
                edges[e] = ed | (1ull << 63);
                atomicAdd(&sh_done, 1);
                if (sj == 1) atomicOr(&pending[dst], 1 << 30);            // a kept suppressor
                const int old = atomicSub(&pending[dst], 1);
                if ((old & 0x3fffffff) == 1) status[dst] = (old >> 30) & 1 ? 2 : 1;      // the node's last open edge
            }
            __syncthreads();
            done = sh_done;
            if (tid == 0) LA_COUNT(5, 1);
            __syncthreads();
        }
        if (tid == 0) LA_COUNT(6, 1);
        for (long long pos = tid; pos < len; pos += LA_NT)
            if (status[pos] == 1) p.keep[img + p.rank_of[seg0 + pos]] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// 5. compaction of the survivors, in rank order, to the front of each image
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LG_CNT) lg_count_kernel(const unsigned char *keep, long long R, int nchunk, int *csum) {
    __shared__ int wsum[LG_CNT / 32];
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const long long r0 = (long long)chunk * LG_CHUNK + (long long)tid * 8;
    const unsigned char *kb = keep + (size_t)b * (size_t)R;
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) if (r0 + k < R) s += kb[r0 + k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((tid & 31) == 0) wsum[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < LG_CNT / 32; ++w) t += wsum[w];
        csum[(size_t)b * nchunk + chunk] = t;
    }
}

__global__ void __launch_bounds__(LG_CNT) lg_write_kernel(RowParams rp, const unsigned char *keep, const u32 *row_of,
                                                          const int *csum, int nchunk, long long out_rows,
                                                          int in_format, int out_format, float *out, int32_t *kept_rows) {
    __shared__ int wsum[LG_CNT / 32], sh_base;
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // survivors in the chunks before this one
    int s = 0;
    for (int c = tid; c < chunk; c += LG_CNT) s += csum[(size_t)b * nchunk + c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < LG_CNT / 32; ++w) t += wsum[w]; sh_base = t; }
    __syncthreads();
    const long long base = sh_base;
    if (base >= out_rows) return;
    __syncthreads();
    const long long R = rp.R;
    const size_t img = (size_t)b * (size_t)R;
    const long long r0 = (long long)chunk * LG_CHUNK + (long long)tid * 8;
    unsigned char f[8];
    int mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { f[k] = (r0 + k < R) ? keep[img + r0 + k] : 0; mine += f[k]; }
    int inc = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    long long pos = base + before + (inc - mine);
    const int W = rp.W;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (!f[k]) continue;
        if (pos < out_rows) {
            const u32 row = row_of[img + r0 + k];
            const float *src = rp.data + (img + row) * W;
            float *o = out + ((size_t)b * (size_t)out_rows + (size_t)pos) * W;
            for (int c = 0; c < W; ++c) o[c] = src[c];
            if (in_format != out_format) {
                float *q = o + rp.coord_start;
                if (!(q[0] < 0)) {
                    if (out_format == VY_FMT_CENTER) {
                        const float l = q[0], t = q[1], r2 = q[2], bt = q[3];
                        q[0] = __fdiv_rn(__fadd_rn(l, r2), 2.0f); q[1] = __fdiv_rn(__fadd_rn(t, bt), 2.0f);
                        q[2] = __fsub_rn(r2, l); q[3] = __fsub_rn(bt, t);
                    } else {
                        const float x = q[0], y = q[1];
                        const float hw = __fdiv_rn(q[2], 2.0f), hh = __fdiv_rn(q[3], 2.0f);
                        q[0] = __fsub_rn(x, hw); q[1] = __fsub_rn(y, hh);
                        q[2] = __fadd_rn(x, hw); q[3] = __fadd_rn(y, hh);
                    }
                }
            }
            if (kept_rows) kept_rows[(size_t)b * (size_t)out_rows + (size_t)pos] = (int32_t)row;
        }
        ++pos;
    }
}

__global__ void lg_fill_kernel(float *out, int *kept, size_t n_out, size_t n_kept) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) out[i] = -1.0f;
    if (kept) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_kept; i += stride) kept[i] = -1;
}

}  // namespace

size_t vy_box_nms_large_workspace_bytes(int B, long long R, int W_elem) {
    (void)W_elem;
    if (B < 1 || R < 1 || (long long)B * R > 0x7fffffffLL) { vy_set_error("box_nms (large): B*R must be < 2^31"); return 0; }
    LgLayout L;
    if (!lg_layout(B, R, &L)) { vy_set_error("box_nms (large): radix-sort workspace query failed (no CUDA device?)"); return 0; }
    return L.total;
}

int vy_box_nms_large(const RowParams &rp, int B, long long K, float overlap_thresh, int force_suppress,
                     int in_format, int out_format, long long out_rows, float *out, int32_t *kept_rows,
                     void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const long long R = rp.R, N = (long long)B * R;
    if (N > 0x7fffffffLL) VY_FAIL(VY_EUNSUPPORTED, "box_nms with topk > %d needs B*R < 2^31 (got %lld)", SEL_KMAX, N);
    LgLayout L;
    if (!lg_layout(B, R, &L)) VY_FAIL(VY_ECUDA, "box_nms (large): radix-sort workspace query failed");
    if (!workspace || workspace_bytes < L.total)
        VY_FAIL(VY_EWORKSPACE, "vy_box_nms_f32: workspace %zu < %zu bytes", workspace_bytes, L.total);
    if (((uintptr_t)workspace & 255) != 0) VY_FAIL(VY_EALIGN, "workspace must be 256-byte aligned");
    char *ws = (char *)workspace;
    int *nvalid = (int *)(ws + L.hdr);
    int *n_seg = nvalid + B;
    u64 *keys_a = (u64 *)(ws + L.keys_a), *keys_b = (u64 *)(ws + L.keys_b);
    u32 *vals_a = (u32 *)(ws + L.vals_a), *vals_b = (u32 *)(ws + L.vals_b), *vals_c = (u32 *)(ws + L.vals_c);
    unsigned char *keep = (unsigned char *)(ws + L.keep);
    const int sms = vy_sm_count();
    const int bitsB = lg_bits(B);
    const bool exhaustive = (force_suppress & 0x100) != 0;
    force_suppress &= 0xff;
    const bool all_pairs = force_suppress || rp.id_index < 0;

    VY_CUDA_CHECK(cudaMemsetAsync(ws + L.hdr, 0, L.hdr_bytes, st));
    VY_CUDA_CHECK(cudaMemsetAsync(keep, 0, (size_t)N, st));
    VY_KERNEL(VY_K_NMS_LARGE, st, (lg_fill_kernel<<<sms * 8, 256, 0, st>>>(out, kept_rows, (size_t)B * out_rows * rp.W,
                                                                         (size_t)B * out_rows)));
    VY_LAUNCH_CHECK("lg_fill_kernel");
    VY_KERNEL(VY_K_NMS_LARGE, st, (lg_key_kernel<<<sms * 8, 256, 0, st>>>(rp, B, keys_a, vals_a, nvalid)));
    VY_LAUNCH_CHECK("lg_key_kernel");
    size_t tb = L.cub_bytes;
    VY_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(ws + L.cub, tb, (const u64 *)keys_a, keys_b, (const u32 *)vals_a,
                                                            vals_b, N, 0, 32 + bitsB, st));
    // vals_b = row_of[(image, rank)]
    if (!all_pairs) {
        VY_KERNEL(VY_K_NMS_LARGE, st, (lg_key2_kernel<<<sms * 8, 256, 0, st>>>(rp, B, K, vals_b, nvalid, keys_a, vals_a)));
        VY_LAUNCH_CHECK("lg_key2_kernel");
        tb = L.cub_bytes;
        VY_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + L.cub, tb, (const u64 *)keys_a, keys_b, (const u32 *)vals_a, vals_c,
                                                      N, 0, 33 + bitsB, st));
        VY_KERNEL(VY_K_NMS_LARGE, st, (lg_seg_kernel<<<sms * 8, 256, 0, st>>>(keys_b, N, (u32 *)(ws + L.seg), n_seg)));
        VY_LAUNCH_CHECK("lg_seg_kernel");
    }
    LgNms p;
    memset(&p, 0, sizeof(p));
    p.rp = rp; p.B = B; p.K = K; p.thr = overlap_thresh; p.all_pairs = all_pairs ? 1 : 0;
    if (overlap_thresh > 0.0f && overlap_thresh < 1e30f) {
        p.thr_lo = overlap_thresh * (1.0f - 9.5367431640625e-07f);      // 2^-20
        p.thr_hi = overlap_thresh * (1.0f + 9.5367431640625e-07f);
    } else {            // thr <= 0 (or absurd): always the exact division
        p.thr_lo = -INFINITY;
        p.thr_hi = INFINITY;
    }
    p.row_of = vals_b; p.keys2 = keys_b; p.rank_of = vals_c; p.seg_starts = (const u32 *)(ws + L.seg);
    p.n_seg = n_seg; p.nvalid = nvalid;
    p.kept_box = (float4 *)(ws + L.kept_box); p.kept_area = (float *)(ws + L.kept_area); p.keep = keep;
    p.seg_starts_rw = (u32 *)(ws + L.seg); p.edges = (u64 *)(ws + L.edges); p.pending = (int *)(ws + L.pending);
    p.status = (unsigned char *)(ws + L.status); p.only_flagged = 0;
    p.sbox = (float4 *)(ws + L.sbox); p.sap = (uint2 *)(ws + L.sap);
    const size_t dyn = (size_t)LG_TILE * (16 + 16 + 4 + 4 + 4) + (size_t)LG_TILE * LG_WORDS * 4;
    const int grid = all_pairs ? (B < sms ? B : sms) : sms;
    // spatial index over the kept boxes unless the threshold is too small for the size bound to prune
    // (or the caller asks for the exhaustive kernel: bit 0x100 of force_suppress, used by the tests)
    const bool use_grid = overlap_thresh >= 0.05f && overlap_thresh < 1e30f && !exhaustive;
    if (use_grid) {
        p.hash_head = (u32 *)keys_a;             // both sorts are done: their input buffers are free
        p.next = vals_a;
        if (all_pairs) VY_CUDA_CHECK(cudaMemsetAsync(keys_a, 0xff, sizeof(u64) * (size_t)N, st));     // (the adjacency kernel zeroes its cells itself)
    }
    // class-aware segments: adjacency + wavefront resolution first; what it flags goes to the tiled kernel (exhaustive
    // variant: the hash arrays hold the adjacency kernel's chains)
    const bool use_adj = use_grid && !all_pairs;
    if (use_adj) {
        if (in_format == VY_FMT_CORNER) VY_KERNEL(VY_K_NMS_LARGE, st, (lg_adj_kernel<VY_FMT_CORNER><<<sms * LA_CTAS, LA_NT, 0, st>>>(p)));
        else VY_KERNEL(VY_K_NMS_LARGE, st, (lg_adj_kernel<VY_FMT_CENTER><<<sms * LA_CTAS, LA_NT, 0, st>>>(p)));
        VY_LAUNCH_CHECK("lg_adj_kernel");
        p.only_flagged = 1;
    }
#define LG_LAUNCH(FMT, GRID) do { \
        VY_CUDA_CHECK(cudaFuncSetAttribute(lg_nms_kernel<FMT, GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
        VY_KERNEL(VY_K_NMS_LARGE, st, (lg_nms_kernel<FMT, GRID><<<grid, LG_NT, dyn, st>>>(p))); } while (0)
    if (in_format == VY_FMT_CORNER) { if (use_grid && !use_adj) LG_LAUNCH(VY_FMT_CORNER, true); else LG_LAUNCH(VY_FMT_CORNER, false); }
    else { if (use_grid && !use_adj) LG_LAUNCH(VY_FMT_CENTER, true); else LG_LAUNCH(VY_FMT_CENTER, false); }
#undef LG_LAUNCH
    VY_LAUNCH_CHECK("lg_nms_kernel");
#ifdef LA_STATS
    if (use_adj) {
        unsigned long long h[8];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, la_stats, sizeof(h));
        fprintf(stderr, "[la_stats] segments %llu | per segment: probes %.0f visits %.0f lower %.0f area-ok %.0f edges %.0f rounds %.1f\n", h[6],
                (double)h[4] / h[6], (double)h[0] / h[6], (double)h[1] / h[6], (double)h[2] / h[6], (double)h[3] / h[6], (double)h[5] / h[6]);
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(la_stats, h, sizeof(h));
    }
#endif
    const int nchunk = (int)((R + LG_CHUNK - 1) / LG_CHUNK);
    int *csum = (int *)(ws + L.csum);
    VY_KERNEL(VY_K_NMS_LARGE, st, (lg_count_kernel<<<dim3(nchunk, B), LG_CNT, 0, st>>>(keep, R, nchunk, csum)));
    VY_LAUNCH_CHECK("lg_count_kernel");
    VY_KERNEL(VY_K_NMS_LARGE, st, (lg_write_kernel<<<dim3(nchunk, B), LG_CNT, 0, st>>>(rp, keep, vals_b, csum, nchunk, out_rows,
                                                                                      in_format, out_format, out, kept_rows)));
    VY_LAUNCH_CHECK("lg_write_kernel");
    return VY_OK;
}
