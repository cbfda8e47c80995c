// vy_stream_alt.cuh -- two alternative streaming passes for the fused decode + box_nms path, kept for A/B measurement.
// NOT part of the product build: compiled only with -DVY_STREAM_ALT (python tools/build_variants.py alt=-DVY_STREAM_ALT)
// and then selected with VY_STREAM_MODE=v2 / v3.  Both pass every parity test of the fused path; both are slower than
// the unit-streaming pass (vy_decode_stream_kernel) on B200 -- DESIGN.md section 4.2 has the measurements and why.
// Included by vy_nms.cu (uses its SelPlan / SelGlobal, vy_tcmin, cp_async16 ...).
#pragma once

// ------------------------------------------------------------------------------------------------
// Tile streaming (the bandwidth pass, second generation): vy_decode_stream2_kernel.
//
//   Every (b, scale, anchor) block of class planes is ONE contiguous array of C*HW floats (channel a*P+5+c,
//   yolo3.py:158-160).  It is cut into tiles of S2_TILE bytes; the global tile sequence is dealt to the CTAs in equal
//   contiguous ranges (each CTA streams a few MB of contiguous memory).  One CTA per SM, four kinds of warps:
//     producer (1)    brings tiles into a shared-memory ring with 1-D bulk copies (cp.async.bulk, one elected lane,
//                     mbarrier complete_tx).  It reads only the head maps, so it starts BEFORE griddepcontrol.wait:
//                     the ring fills while the sample kernel drains.
//     table (S2_TABW) per block the per-position logit bound t_c >= logit(s_min / sigma(t_obj)) (vy_tcmin) for the image's
//                     bound s_min, HW floats in shared memory (double-buffered: the next block's table is built while
//                     this one streams).  The old pass recomputed these bounds in the prologue of every unit.
//     consumers (24)  S2_WPS warps per ring stage test float4 of the tile against float4 of the table (position =
//                     element index mod HW): two LDS.128 and four FSETP per 16 bytes, nothing else in the loop.  A
//                     flagged float4 is QUEUED as one 32-bit entry (first element + validity mask) in a warp-private
//                     ring -- a ballot and a store per row that holds a hit -- and the stage goes back to the producer;
//                     full batches of 32 entries are then scored one per lane: logits and objectness come back from L2
//                     (the bulk copy left them there), the score is the decode kernel's, the key is tested against the
//                     image's bound and survivors go to the warp's key buffer (one atomic + one 256-byte store per 32).
//   Planes whose size is not a multiple of 4 floats (13^2, 19^2 grids) go through the same ring: the copy starts at the
//   16-byte boundary below the tile and the consumers shift.
// ------------------------------------------------------------------------------------------------
#ifndef VY_STREAM_DEFAULT
#define VY_STREAM_DEFAULT 3                          // 0: unit streaming (round 1), 2: tile streaming, 3: segment streaming
#endif
#ifndef S2_TILE
#define S2_TILE 24576                               // bytes per tile
#endif
#ifndef S2_WARPS
#define S2_WARPS 24                                 // consumer warps: S2_WPS per ring stage
#endif
#ifndef S2_WPS
#define S2_WPS 4                                    // warps sharing a tile: warp h of a stage tests the row groups g = h (mod S2_WPS)
#endif
#ifndef S2_TABW
#define S2_TABW 2                                   // table warps
#endif
constexpr int S2_MAX_STAGES = S2_WARPS / S2_WPS;    // ring depth is chosen at launch from the shared-memory budget
constexpr int S2_NT = (S2_WARPS + 1 + S2_TABW) * 32;
constexpr int S2_STAGE_BYTES = S2_TILE + 256;       // a tile lands at the same offset inside a 128-byte line as its source (+ <= 12 bytes of shift)
constexpr int S2_K = S2_TILE / 16 / 32;             // float4 rows (of 32 lanes) per tile
constexpr int S2_HQ = 128;                          // queue entries per warp (a ring; a batch leaves at 32)
static_assert(S2_K % 4 == 0 && S2_K < 64, "tile size");
static_assert((S2_HQ & (S2_HQ - 1)) == 0, "queue size");

// what the producer tells the consumers about a tile: one 16-byte word
//   x = e0      first element of the tile inside its (b, s, a) block (the hit path turns element indices into rows)
//   y = n       floats in the tile
//   z = zpos    position (inside its plane) of the tile's first element: 0 unless the plane is longer than a tile
//   w = b:16 | s:2 | a:3 | tab:1 | tpar:1 | aligned:1 | shift:2 | tail:2 | doff:3
//       tab / tpar = table buffer of the block and the parity of its completion (the block is the CTA's ord-th: tab =
//       ord & 1, tpar = (ord >> 1) & 1); shift = floats in front of the tile in the stage (the copy starts at the 16-byte
//       boundary below it); tail = floats of the tile that are NOT in the stage (the copy must not run past the end of
//       the tensor: the last <= 3 floats of a tensor whose end is not 16-byte aligned); doff = 16-byte units between the
//       stage base and the copy: source and destination of a bulk copy sit at the same offset inside their 128-byte lines
constexpr u32 S2_BLK_MASK = 0x7fffffu;             // b, s, a, tab, tpar: equal for the tiles of one block
struct S2Tile { int b, s, a, tab, tpar, aligned, shift, tail, doff; };
__device__ __forceinline__ u32 s2_pack(const S2Tile &t) {
    return (u32)t.b | ((u32)t.s << 16) | ((u32)t.a << 18) | ((u32)t.tab << 21) | ((u32)t.tpar << 22) |
           ((u32)t.aligned << 23) | ((u32)t.shift << 24) | ((u32)t.tail << 26) | ((u32)t.doff << 28);
}

__device__ __forceinline__ u32 s2_smem(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s2_mbar_init(u64 *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s2_smem(bar)), "r"(count) : "memory");
}
// (a wait that has not come true after ~2^24 polls -- seconds -- is a protocol error: trap instead of hanging the device)
__device__ __forceinline__ void s2_mbar_wait(u64 *bar, u32 parity) {
    u32 ok = 0, spins = 0;
    const u32 a = s2_smem(bar);
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ bool s2_mbar_test(u64 *bar, u32 parity) {
    u32 ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(s2_smem(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void s2_mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s2_smem(bar)) : "memory");
}
__device__ __forceinline__ void s2_mbar_expect(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s2_smem(bar)), "r"(bytes) : "memory");
}

// tile r of an image -> (scale, anchor, tile inside the block)
__device__ __forceinline__ void s2_tile_block(const VyHeads &hd, const SelPlan &pl, int r, int &s, int &a, int &t) {
    s = 0;
    while (s + 1 < hd.n_scales && r >= pl.tile_begin[s + 1]) ++s;
    r -= pl.tile_begin[s];
    a = r / pl.tiles_blk[s];
    t = r - a * pl.tiles_blk[s];
}

// The table of block (b, s, a) -- tab_hwp[s] floats: the bound of every position, positions 0 .. 2 again behind the
// last one (a shifted float4 may wrap), +inf padding.  The objectness plane comes into the table buffer itself with one
// bulk copy (s2_table_fetch, one thread; planes that are not 16-byte aligned are read with plain loads instead) and is
// turned into bounds in place by threads t0, t0 + nthr, ... (whole warps).  The image's bound is the minimum over its
// sample jobs (kept as complements).
__device__ __forceinline__ bool s2_table_vec(const VyHeads &hd, int s) { return hd.sc[s].vec == 4; }
__device__ __forceinline__ void s2_table_fetch(const VyHeads &hd, int b, int s, int a, float *dst, u64 *bar) {
    const VyScale &sc = hd.sc[s];
    const float *obj = sc.head + ((size_t)(b * hd.A + a) * hd.P + 4) * (size_t)sc.HW;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the buffer's last table was written with plain stores
    if (s2_table_vec(hd, s)) {
        s2_mbar_expect(bar, (u32)sc.HW * 4u);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(s2_smem(dst)), "l"((const void *)obj), "r"((u32)sc.HW * 4u), "r"(s2_smem(bar)) : "memory");
    } else {
        s2_mbar_arrive(bar);                              // (keeps the barrier's phases in step with the blocks)
    }
}
__device__ __forceinline__ void s2_build_table(const VyHeads &hd, const SelPlan &pl, const SelGlobal &g, int b, int s,
                                               int a, float *dst, u64 *thr_out, int t0, int nthr) {
    const VyScale &sc = hd.sc[s];
    const int HW = sc.HW, hwp = pl.tab_hwp[s];
    const float *obj = sc.head + ((size_t)(b * hd.A + a) * hd.P + 4) * (size_t)HW;
    u64 m = (t0 & 31) < g.Gs ? g.sslots[(size_t)b * g.Gs + (t0 & 31)] : 0ull;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const u64 o = sel_shfl_xor_u64(m, off); m = o > m ? o : m; }
    const u64 thr = ~m;
    const float smin = fmaxf(thr ? vy_key_score(thr) : pl.valid_thresh, pl.valid_thresh);
    const bool vec = s2_table_vec(hd, s);
    for (int p0 = t0; p0 < HW; p0 += 8 * nthr) {
        float to[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int pos = p0 + k * nthr;
            to[k] = pos < HW ? (vec ? dst[pos] : vy_ldg32(obj + pos)) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int pos = p0 + k * nthr;
            if (pos < HW) {
#ifdef S2_DBG_NOTABLE
                const float v = to[k] + 1e30f;
#else
                const float v = vy_tcmin(smin, vy_sigmoid(to[k]));
#endif
                dst[pos] = v;
                if (pos < 3) dst[HW + pos] = v;
            }
        }
    }
    for (int pos = HW + 3 + t0; pos < hwp; pos += nthr) dst[pos] = CUDART_INF_F;
    if (HW < 3 && t0 == 0) for (int pos = HW; pos < HW + 3; ++pos) dst[pos] = CUDART_INF_F;     // (degenerate grids never wrap)
    if (t0 == 0) *thr_out = thr;
}

// ---- the rare path of a consumer warp.  A queue entry names a float4 of the block: bits 0..27 = (first element + 3)
// (a shifted float4 of the block's first tile starts up to 3 elements in front of it), bits 28..31 = which of its four
// elements belong to the tile.  Every lane scores one entry: <= 4 elements, logits and objectness from L2.
__device__ __noinline__ int s2_score(const VyHeads &hd, float valid_thresh, u32 w, u64 thr, const u32 *hq, int qh, int nb,
                                     u64 *wbuf, int cnt, const SelGlobal &g) {
    const int lane = threadIdx.x & 31;
    const u32 lt_mask = (1u << lane) - 1u;
    const int b = (int)(w & 0xffffu), s = (int)((w >> 16) & 3u), a = (int)((w >> 18) & 7u);
    const VyScale &sc = hd.sc[s];
    const int HW = sc.HW;
    const float *pc0 = sc.head + ((size_t)(b * hd.A + a) * hd.P + 5) * (size_t)HW;       // class plane 0 of the block
    const u32 ent = lane < nb ? hq[(qh + lane) & (S2_HQ - 1)] : 0u;
    const u32 mask = ent >> 28;
    const int e0 = (int)(ent & 0x0fffffffu) - 3;
    const int ef = e0 < 0 ? 0 : e0;
    const int pl0 = ef / HW, pos0 = ef - pl0 * HW;
    float tv[4], to[4];
    u32 row[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        tv[v] = 0.0f; to[v] = 0.0f; row[v] = 0u;
        if ((mask >> v) & 1u) {
            const int ev = e0 + v;
            int p = pos0 + (ev - ef), pln = pl0;
            if (p >= HW) { p -= HW; ++pln; }
            tv[v] = vy_ldg32(pc0 + ev);
            to[v] = vy_ldg32(pc0 + p - HW);              // the objectness plane sits right below class plane 0
            row[v] = (u32)(sc.row_off + a) + (u32)pln * (u32)sc.n_s + (u32)p * (u32)hd.A;
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        bool ok = false;
        u64 key = 0;
        if ((mask >> v) & 1u) {
            const float sv = vy_score(tv[v], vy_sigmoid(to[v]));
            if (sv > valid_thresh) {
                key = vy_make_key(sv, row[v]);
                ok = key >= thr;
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
    return cnt;
}
// end of a block: the warp's keys go to the image's list
__device__ __noinline__ void s2_flush_keys(u32 w, const u64 *wbuf, int cnt, const SelGlobal &g) {
    const int lane = threadIdx.x & 31;
    const int b = (int)(w & 0xffffu);
    int base = 0;
    if (lane == 0) base = atomicAdd(g.scount + b, cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < cnt && base + lane < g.slist_cap) g.slist[(size_t)b * g.slist_cap + base + lane] = wbuf[lane];
    __syncwarp();
}

// One CTA per SM.
//   barriers   full[stage]   the tile's bytes have landed (producer: expect_tx; waited for by the stage's warps)
//              empty[stage]  the stage's warps have tested the tile (S2_WPS arrivals)
//              tab_full[buf] the block's table is in buffer buf (S2_TABW arrivals; every consumer warp waits for it once
//                            per block it meets)
//              tab_free[buf] every tile of the buffer's previous block is through (producer -> table warps)
//   Tiles go to WHICHEVER stage is free (a warp that is scoring hits or flushing a block holds its stage for a few
//   microseconds; with tiles bound to stages in order the whole ring would wait behind it).  A stage's warps take what
//   arrives, one round after the other, and stop at a tile of zero floats.
__global__ void __launch_bounds__(S2_NT, 1)
vy_decode_stream2_kernel(const __grid_constant__ VyHeads hd, const __grid_constant__ SelPlan pl,
                         const __grid_constant__ SelGlobal g, int n_stages) {
    extern __shared__ __align__(128) unsigned char s2_dyn[];          // ring [n_stages][S2_STAGE_BYTES], tables [2][tab_max]
    __shared__ __align__(8) u64 bar_full[S2_MAX_STAGES], bar_empty[S2_MAX_STAGES], bar_tab_full[2], bar_tab_free[2], bar_obj;
    __shared__ u64 tab_thr[2];
    __shared__ int4 meta[S2_MAX_STAGES];
    __shared__ int p_busy[S2_MAX_STAGES], p_par[S2_MAX_STAGES], p_buf[S2_MAX_STAGES], p_out[2], p_closed[2], p_next;   // producer's books
    __shared__ u64 wbuf_all[S2_WARPS][64];
    __shared__ u32 hq_all[S2_WARPS][S2_HQ];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned char *ring = s2_dyn;
    float *tabs = (float *)(s2_dyn + (size_t)n_stages * S2_STAGE_BYTES);
    if (tid == 0) {
        for (int i = 0; i < n_stages; ++i) { s2_mbar_init(&bar_full[i], 1); s2_mbar_init(&bar_empty[i], S2_WPS); }
        for (int i = 0; i < 2; ++i) { s2_mbar_init(&bar_tab_full[i], S2_TABW); s2_mbar_init(&bar_tab_free[i], 1); }
        s2_mbar_init(&bar_obj, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < n_stages; ++i) { p_busy[i] = 0; p_par[i] = 0; p_buf[i] = 0; }
        p_out[0] = p_out[1] = 0; p_closed[0] = p_closed[1] = 0; p_next = 0;
    }
    __syncthreads();
    const long long t_begin = pl.n_tiles * (long long)blockIdx.x / gridDim.x;
    const long long t_end = pl.n_tiles * (long long)(blockIdx.x + 1) / gridDim.x;
    const int n_mine = (int)(t_end - t_begin);
    const int b0 = (int)(t_begin / pl.tiles_per_image);
    const int r0 = (int)(t_begin - (long long)b0 * pl.tiles_per_image);

    if (wid != S2_WARPS) {
        // the table of the range's first block is built by everyone but the producer (the ring fills meanwhile)
        int s, a, tt;
        s2_tile_block(hd, pl, r0, s, a, tt);
        if (tid == (S2_WARPS + 1) * 32) s2_table_fetch(hd, b0, s, a, tabs, &bar_obj);
        vy_grid_dep_wait();                               // the sample kernel's bounds, lists and counters
        s2_mbar_wait(&bar_obj, 0u);
        s2_build_table(hd, pl, g, b0, s, a, tabs, &tab_thr[0], tid < S2_WARPS * 32 ? tid : tid - 32, S2_NT - 32);
        asm volatile("bar.sync 1, %0;" :: "n"(S2_NT - 32) : "memory");
    }
    if (wid == S2_WARPS) {
        // ------------------------------------------------------------------ producer warp
        // A batch = 32 consecutive tiles, one per lane.  The lanes work out their tiles side by side (the integer
        // divisions of the decode cost one lane as much as 32) and keep the descriptors in registers; then the tiles are
        // issued in order, each by its own lane: wait for a stage, one 16-byte word of metadata, expect_tx, bulk copy.
        // The ring covers the pause of the next batch's decode.  (No griddepcontrol.wait here: the head maps are inputs.)
        // a stage whose warps have arrived goes back on the books as free; when that was the last tile of a block the
        // range has left, the block's table buffer goes back to the table warps
        auto reap = [&](int st) {
            if (p_busy[st] && s2_mbar_test(&bar_empty[st], (u32)p_par[st])) {
                p_busy[st] = 0; p_par[st] ^= 1;
                const int bf = p_buf[st];
                if (--p_out[bf] == 0 && p_closed[bf]) { p_closed[bf] = 0; s2_mbar_arrive(&bar_tab_free[bf]); }
            }
            return !p_busy[st];
        };
        int blocks_before = 0;                            // blocks this CTA's range has entered so far
        for (int i0 = 0; i0 < n_mine; i0 += 32) {
            const int it = i0 + lane;                     // tile index inside the CTA's range
            const bool act = it < n_mine;
            int b = b0, s = 0, a = 0, t = 0;
            if (act) {
                int r = r0 + it;
                const int db = r / pl.tiles_per_image;
                b += db; r -= db * pl.tiles_per_image;
                s2_tile_block(hd, pl, r, s, a, t);
            }
            const bool first = act && (t == 0 || it == 0);
            const u32 fb = __ballot_sync(0xffffffffu, first);
            const u32 below = fb & ((1u << lane) - 1u);
            const int ord = blocks_before + __popc(below) + (first ? 1 : 0) - 1;      // which block of the range this tile is in
            blocks_before += __popc(fb);
            const VyScale &sc = hd.sc[s];
            int e0, n, zpos;
            if (pl.tile_tpp[s] > 0) {                     // long planes: tiles inside one plane
                const int c = t / pl.tile_tpp[s], part = t - c * pl.tile_tpp[s];
                zpos = part * (S2_TILE / 4);
                e0 = c * sc.HW + zpos;
                n = min(S2_TILE / 4, sc.HW - zpos);
            } else {                                      // whole planes per tile
                const int c0 = t * pl.tile_ppt[s];
                zpos = 0;
                e0 = c0 * sc.HW;
                n = min(pl.tile_ppt[s], hd.C - c0) * sc.HW;
            }
            const float *src = sc.head + ((size_t)(b * hd.A + a) * hd.P + 5) * (size_t)sc.HW + e0;
            const uintptr_t addr = (uintptr_t)src;
            S2Tile ti;
            ti.b = b; ti.s = s; ti.a = a; ti.tab = ord & 1; ti.tpar = (ord >> 1) & 1;
            ti.shift = (int)((addr & 15) >> 2);
            const uintptr_t src_al = addr - (uintptr_t)ti.shift * 4;
            ti.doff = (int)((src_al & 127) >> 4);
            long long bytes = ((long long)(ti.shift + n) * 4 + 15) & ~15LL;
            const uintptr_t end_al = ((uintptr_t)(sc.head + (size_t)hd.B * hd.A * hd.P * (size_t)sc.HW)) & ~(uintptr_t)15;
            if (src_al + (uintptr_t)bytes > end_al) bytes = end_al > src_al ? (long long)(end_al - src_al) : 0;
            int n_smem = (int)(bytes / 4) - ti.shift;
            n_smem = n_smem < 0 ? 0 : (n_smem > n ? n : n_smem);
            ti.tail = n - n_smem;                         // <= 3
            ti.aligned = (ti.shift == 0 && (sc.HW & 3) == 0) ? 1 : 0;
            const int4 word = make_int4(e0, n, zpos, (int)s2_pack(ti));
            const int n_batch = min(32, n_mine - i0);
            for (int i = 0; i < n_batch; ++i) {           // in tile order, each tile by its own lane
                if (lane == i) {
                    if (first) {
                        // at most two blocks are in flight (one per table buffer; the parity waits and the per-buffer
                        // tile counts rely on it): the buffer's previous block, ord - 2, must be through
                        for (u32 spins = 0; p_out[ti.tab] > 0; ) {
                            for (int st = 0; st < n_stages; ++st) reap(st);
                            if (++spins > (1u << 24)) __trap();
                        }
                        if (ord >= 1) {                   // the range leaves block ord - 1
                            const int pb = ti.tab ^ 1;
                            if (p_out[pb] == 0) s2_mbar_arrive(&bar_tab_free[pb]); else p_closed[pb] = 1;
                        }
                    }
                    int st = p_next;
                    for (u32 spins = 0; !reap(st); ) {    // the next free stage, starting behind the last one used
                        if (++st == n_stages) st = 0;
                        if (++spins > (1u << 26)) __trap();
                    }
                    p_next = st + 1 == n_stages ? 0 : st + 1;
                    p_busy[st] = 1; p_buf[st] = ti.tab; p_out[ti.tab] += 1;
                    meta[st] = word;
                    s2_mbar_expect(&bar_full[st], (u32)bytes);
                    if (bytes)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     :: "r"(s2_smem(ring + (size_t)st * S2_STAGE_BYTES + ti.doff * 16)), "l"((const void *)src_al),
                                        "r"((u32)bytes), "r"(s2_smem(&bar_full[st])) : "memory");
                }
                __syncwarp();
            }
        }
        // every stage gets a last message: a tile of zero floats -- once ALL of them are through (the table warps may
        // still be waiting for the buffer of a block whose last tile sits in any stage)
        if (lane == 0) {
            for (u32 spins = 0;;) {
                bool all_free = true;
                for (int st = 0; st < n_stages; ++st) all_free &= reap(st);
                if (all_free) break;
                if (++spins > (1u << 26)) __trap();
            }
            for (int st = 0; st < n_stages; ++st) {
                meta[st] = make_int4(0, 0, 0, -1);
                s2_mbar_arrive(&bar_full[st]);
            }
        }
    } else if (wid > S2_WARPS) {
        // ------------------------------------------------------------------ table warps
        // block ord of the range -> buffer ord & 1, as soon as block ord - 2 is through
        const int t0 = tid - (S2_WARPS + 1) * 32;
        long long t = t_begin;
        for (int ord = 0; t < t_end; ++ord) {
            const int b = (int)(t / pl.tiles_per_image);
            int s, a, tt;
            s2_tile_block(hd, pl, (int)(t - (long long)b * pl.tiles_per_image), s, a, tt);
            const int buf = ord & 1;
            if (ord >= 2) s2_mbar_wait(&bar_tab_free[buf], (u32)(((ord - 2) >> 1) & 1));
            if (ord >= 1) {
                float *dst = tabs + (size_t)buf * pl.tab_max;
                if (t0 == 0) s2_table_fetch(hd, b, s, a, dst, &bar_obj);
                s2_mbar_wait(&bar_obj, (u32)(ord & 1));
                s2_build_table(hd, pl, g, b, s, a, dst, &tab_thr[buf], t0, S2_TABW * 32);
            }
            __syncwarp();
            if (lane == 0) s2_mbar_arrive(&bar_tab_full[buf]);
            t += pl.tiles_blk[s] - tt;
        }
    } else {
        // ------------------------------------------------------------------ consumer warps
        // Warp `wid` serves ring stage wid % n_stages together with S2_WPS - 1 others (a stage's rounds must be waited
        // for in order: a parity wait cannot tell round r from round r - 2): of a tile's float4 rows of 32 lanes, warp
        // `half` takes the groups of four rows g = half (mod S2_WPS).
        // Where those sit in the block's table: tiles start at position 0 of a plane (or -- planes longer than a tile --
        // at zpos), so the first index is the same for every tile of a block and the rest follow by adding 32 (mod the
        // plane size).
        u64 *wbuf = wbuf_all[wid];
        u32 *hq = hq_all[wid];
        int qh = 0, qt = 0, cnt = 0;                      // queue head / tail (running), keys waiting in wbuf: warp-uniform
        const u32 lt_mask = (1u << lane) - 1u;
        const float valid_thresh = pl.valid_thresh;
        const int tab_max = pl.tab_max;
        u32 cur_w = 0xffffffffu;                          // metadata word of the block this warp is in
        u64 cur_thr = 0;                                  // and the image's bound key
        int idx0 = 0, idx0h = 0, plane = 4;               // first table index of this lane; plane size (float4, or floats when shifted)
        const int stage = wid % n_stages, half = wid / n_stages;        // warps beyond S2_WPS * n_stages have no stage
        for (u32 phase = 0; half < S2_WPS; phase ^= 1u) {
            s2_mbar_wait(&bar_full[stage], phase);
            const int4 word = meta[stage];
            const u32 w = (u32)word.w;
            const int n = word.y;
            if (n == 0) break;                            // the producer's last message
            if ((w ^ cur_w) & S2_BLK_MASK) {              // the warp enters another block
                while (qt != qh) {
                    const int nb = min(32, qt - qh);
                    cnt = s2_score(hd, valid_thresh, cur_w, cur_thr, hq, qh, nb, wbuf, cnt, g);
                    qh += nb;
                }
                if (cnt > 0) { s2_flush_keys(cur_w, wbuf, cnt, g); cnt = 0; }
                const u32 buf = (w >> 21) & 1u;
                s2_mbar_wait(&bar_tab_full[buf], (w >> 22) & 1u);
                cur_thr = tab_thr[buf];
                const int HW = hd.sc[(w >> 16) & 3u].HW;
                if ((w >> 23) & 1u) { plane = HW >> 2; idx0 = lane % plane; idx0h = (idx0 + 128 * half) % plane; }
                else { plane = HW; idx0 = (4 * lane) % plane; }
            }
            cur_w = w;
            const float *tab = tabs + ((w >> 21) & 1u) * tab_max;
            const unsigned char *st = ring + (size_t)stage * S2_STAGE_BYTES + (w >> 28) * 16;
            const bool aligned = (w >> 23) & 1u;
            const int shift = aligned ? 0 : (int)((w >> 24) & 3u);
            u64 hits = 0;                                 // bit k: this lane's float4 of row k has an element at or above its bound
#ifndef S2_DBG_NOCOMPARE
            if (aligned) {
                // float4 of the tile against float4 of the table
                const int nf4 = n >> 2;
                const float4 *d4 = (const float4 *)st + lane;
                const float4 *t4 = (const float4 *)tab + (word.z >> 2);
                const int step = 32 % plane, skip = (128 * (S2_WPS - 1)) % plane;
                int idx = idx0h;                          // this warp's first group of four rows starts at row 4 * half
                const int kmax = (nf4 + 31) >> 5;         // float4 rows of 32 lanes in this tile (<= S2_K)
                // four rows at a time: the eight loads first, then branch-free compares.  Rows past the end of the tile
                // are read all the same (stale bytes of the stage, a valid table index) and masked out.
                for (int k0 = 4 * half; k0 < kmax; k0 += 4 * S2_WPS) {
                    float4 v[4], t[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        v[u] = d4[32 * (k0 + u)];
                        t[u] = t4[idx];
                        idx += step;
                        if (idx >= plane) idx -= plane;
                    }
                    idx += skip;                          // over the other warps' groups
                    if (idx >= plane) idx -= plane;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        u32 h;
                        asm("{\n\t.reg .pred p;\n\t"
                            "setp.ge.f32 p, %1, %5;\n\t"
                            "setp.ge.or.f32 p, %2, %6, p;\n\t"
                            "setp.ge.or.f32 p, %3, %7, p;\n\t"
                            "setp.ge.or.f32 p, %4, %8, p;\n\t"
                            "selp.u32 %0, 1, 0, p;\n\t}"
                            : "=r"(h) : "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w), "f"(t[u].x), "f"(t[u].y), "f"(t[u].z), "f"(t[u].w));
                        h &= (lane + 32 * (k0 + u) < nf4) ? 1u : 0u;
                        hits |= (u64)h << (k0 + u);
                    }
                }
            } else {
                // shifted / odd plane size: stage float (shift + el) is element e0 + el of the block
                const int n_smem = n - (int)((w >> 26) & 3u);
                const int nf4 = (shift + n + 3) >> 2;
                const float4 *d4 = (const float4 *)st;
                const float *pc0e = nullptr;
                if (n_smem < n) {
                    const int b = (int)(w & 0xffffu), a = (int)((w >> 18) & 7u);
                    pc0e = hd.sc[(w >> 16) & 3u].head + ((size_t)(b * hd.A + a) * hd.P + 5) * (size_t)plane + word.x;
                }
                int pb = idx0 + word.z - shift;           // position of stage float 4*lane (zpos > 0 only in planes longer than a tile)
                if (pb < 0) pb += plane;
                if (pb >= plane) pb -= plane;
                const int kmax = (nf4 + 31) >> 5;
                const int step_w = (128 * S2_WPS) % plane;
                pb += 128 * half;
                if (pb >= plane) pb %= plane;
                for (int k = half; k < kmax; k += S2_WPS) {
                    const int j = lane + 32 * k;
                    const float4 v4 = d4[j];              // (rows past the end: stale bytes of the stage, masked below)
                    const float t0 = tab[pb], t1 = tab[pb + 1], t2 = tab[pb + 2], t3 = tab[pb + 3];
                    const int el = 4 * j - shift;         // element of v4.x
                    float x0 = v4.x, x1 = v4.y, x2 = v4.z, x3 = v4.w;
                    if (n_smem < n) {                     // the tensor's last <= 3 floats are not in the stage
                        if (el >= n_smem && el < n) x0 = vy_ldg32(pc0e + el);
                        if (el + 1 >= n_smem && el + 1 < n) x1 = vy_ldg32(pc0e + el + 1);
                        if (el + 2 >= n_smem && el + 2 < n) x2 = vy_ldg32(pc0e + el + 2);
                        if (el + 3 >= n_smem && el + 3 < n) x3 = vy_ldg32(pc0e + el + 3);
                    }
                    const u32 h = ((x0 >= t0) & (el >= 0) & (el < n)) | ((x1 >= t1) & (el + 1 >= 0) & (el + 1 < n)) |
                                  ((x2 >= t2) & (el + 2 >= 0) & (el + 2 < n)) | ((x3 >= t3) & (el + 3 >= 0) & (el + 3 < n));
                    hits |= (u64)(h & 1u) << k;
                    pb += step_w;
                    if (pb >= plane) pb -= plane;
                }
            }
#endif
            // rare: the flagged float4 go to the warp's queue, one round per ROW that holds any (a ballot, a store)
#ifndef S2_DBG_NOQUEUE
            u32 rows_lo = __reduce_or_sync(0xffffffffu, (u32)hits), rows_hi = __reduce_or_sync(0xffffffffu, (u32)(hits >> 32));
            while (rows_lo | rows_hi) {
                int k;
                if (rows_lo) { k = __ffs(rows_lo) - 1; rows_lo &= rows_lo - 1; }
                else { k = 32 + __ffs(rows_hi) - 1; rows_hi &= rows_hi - 1; }
                if (qt - qh > S2_HQ - 32) {               // (dense hits: make room for a row)
                    cnt = s2_score(hd, valid_thresh, w, cur_thr, hq, qh, 32, wbuf, cnt, g);
                    qh += 32;
                }
                const bool has = (hits >> k) & 1ull;
                const u32 bal = __ballot_sync(0xffffffffu, has);
                if (has) {
                    const int el = 4 * (lane + 32 * k) - shift;           // element (inside the tile) of the float4's first float
                    u32 m = 0xfu;
                    if (!aligned)
                        m = ((el >= 0 && el < n) ? 1u : 0u) | ((el + 1 >= 0 && el + 1 < n) ? 2u : 0u) |
                            ((el + 2 >= 0 && el + 2 < n) ? 4u : 0u) | ((el + 3 >= 0 && el + 3 < n) ? 8u : 0u);
                    hq[(qt + __popc(bal & lt_mask)) & (S2_HQ - 1)] = (u32)(word.x + el + 3) | (m << 28);
                }
                qt += __popc(bal);
            }
#endif
            // the stage is free as soon as it has been tested and its hits are queued: hand it back before scoring them --
            // that is a round trip to L2 which must not hold up the ring
            __syncwarp();
            if (lane == 0) s2_mbar_arrive(&bar_empty[stage]);
#ifndef S2_DBG_NOSCORE
            while (qt - qh >= 32) {
                cnt = s2_score(hd, valid_thresh, w, cur_thr, hq, qh, 32, wbuf, cnt, g);
                qh += 32;
            }
#else
            while (qt - qh >= 32) qh += 32;
#endif
        }
        while (qt != qh) {
            const int nb = min(32, qt - qh);
            cnt = s2_score(hd, valid_thresh, cur_w, cur_thr, hq, qh, nb, wbuf, cnt, g);
            qh += nb;
        }
        if (cnt > 0) s2_flush_keys(cur_w, wbuf, cnt, g);
    }
    vy_grid_dep_trigger();
}

// ------------------------------------------------------------------------------------------------
// Segment streaming (the bandwidth pass, third generation): vy_decode_stream3_kernel.
//
//   Every (b, scale, anchor) block of class planes is ONE contiguous array of C*HW floats (channel a*P+5+c,
//   yolo3.py:158-160).  The global sequence of tiles (S2_TILE bytes, an accounting unit only) is dealt to the CTAs in
//   equal contiguous ranges; a CTA cuts its range into SEGMENTS -- the part of a range that lies in one block -- and a
//   group of its warps streams a segment side by side, every warp a contiguous run of 512-byte rows:
//     * data movement: warp-private ring of S3_NSLOT slots x S3_G rows in shared memory, refilled with ONE bulk copy
//       (cp.async.bulk, elected lane, mbarrier complete_tx) per slot right after the slot has been tested -- no producer
//       warp, no hand-off between warps, 144 KB in flight per SM;
//     * the per-position logit bound t_c >= logit(s_min / sigma(t_obj)) (vy_tcmin) comes from a table of HW floats in
//       shared memory that the group builds once per segment (the unit pass recomputed the bounds in the prologue of
//       every unit: up to half of its instructions); position = element index mod HW;
//     * the loop: two LDS.128 and four FSETP per 16 bytes; a flagged float4 is queued as one 32-bit entry (first
//       element + validity mask) in a warp-private queue and batches of 32 entries are scored one per lane (logits and
//       objectness come back from L2), exactly the decode kernel's score, tested against the image's bound key.
//   Groups: 12 warps share a 76^2 table (23 KB); smaller planes leave room for several tables, so the CTA splits into
//   3 groups of 4 warps (38^2) or 12 single warps (19^2) that work on different blocks at once.
//   Planes whose size is not a multiple of 4 floats (13^2, 19^2 grids) take the same path: the copy starts at the 16-byte
//   boundary below the segment and the compare shifts (scalar table loads).
// ------------------------------------------------------------------------------------------------
#ifndef S3_WARPS
#define S3_WARPS 12
#endif
#ifndef S3_G
#define S3_G 4                                       // 512-byte rows per bulk copy
#endif
#ifndef S3_NSLOT
#define S3_NSLOT 3
#endif
#ifndef S3_BULK
#define S3_BULK 0                                    // 1: one cp.async.bulk per slot (elected lane, mbarrier); 0: cp.async 16 B per lane
#endif
#ifndef S3_PF_GROUPS
#define S3_PF_GROUPS 0                               // L2 prefetch distance, in bulk copies (0: none)
#endif
#ifndef S3_CTAS_PER_SM
#define S3_CTAS_PER_SM 2
#endif
constexpr int S3_NT = S3_WARPS * 32;
constexpr int S3_SLOT_BYTES = S3_G * 512;
constexpr int S3_RING_BYTES = S3_NSLOT * S3_SLOT_BYTES;      // per warp
constexpr int S3_HQ = 64;                            // queue entries per warp (a ring; a batch leaves at 32, a row adds <= 32)

// ---- the rare path.  A queue entry names a float4 of the block: bits 0..27 = (first element + 3) (a shifted float4 at
// the start of a block begins up to 3 elements in front of it), bits 28..31 = which of its four elements take part.
// Every lane scores one entry: <= 4 elements, logits and objectness from L2.
struct S3Seg {                   // what a group streams: elements [e_lo, e_hi) of block (b, s, a)
    int b, s, a;
    int e_lo, n;                 // first element, element count
    int shift;                   // floats between the 16-byte boundary the copy starts at and e_lo
    int n_smem;                  // elements that reach shared memory (the rest, <= 3: the tensor ends off a 16-byte boundary)
    long long bytes;             // bytes copied
};
__device__ __noinline__ int s3_score(const VyHeads &hd, float valid_thresh, int b, int s, int a, u64 thr, const u32 *hq,
                                     int qh, int nb, u64 *wbuf, int cnt, const SelGlobal &g) {
    const int lane = threadIdx.x & 31;
    const u32 lt_mask = (1u << lane) - 1u;
    const VyScale &sc = hd.sc[s];
    const int HW = sc.HW;
    const float *pc0 = sc.head + ((size_t)(b * hd.A + a) * hd.P + 5) * (size_t)HW;       // class plane 0 of the block
    const u32 ent = lane < nb ? hq[(qh + lane) & (S3_HQ - 1)] : 0u;
    const u32 mask = ent >> 28;
    const int e0 = (int)(ent & 0x0fffffffu) - 3;
    const int ef = e0 < 0 ? 0 : e0;
    const int pl0 = ef / HW, pos0 = ef - pl0 * HW;
    float tv[4], to[4];
    u32 row[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        tv[v] = 0.0f; to[v] = 0.0f; row[v] = 0u;
        if ((mask >> v) & 1u) {
            const int ev = e0 + v;
            int p = pos0 + (ev - ef), pln = pl0;
            if (p >= HW) { p -= HW; ++pln; }
            tv[v] = vy_ldg32(pc0 + ev);
            to[v] = vy_ldg32(pc0 + p - HW);              // the objectness plane sits right below class plane 0
            row[v] = (u32)(sc.row_off + a) + (u32)pln * (u32)sc.n_s + (u32)p * (u32)hd.A;
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        bool ok = false;
        u64 key = 0;
        if ((mask >> v) & 1u) {
            const float sv = vy_score(tv[v], vy_sigmoid(to[v]));
            if (sv > valid_thresh) {
                key = vy_make_key(sv, row[v]);
                ok = key >= thr;
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
    return cnt;
}
// the warp's keys go to the image's list
__device__ __noinline__ void s3_flush_keys(int b, const u64 *wbuf, int cnt, const SelGlobal &g) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(g.scount + b, cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < cnt && base + lane < g.slist_cap) g.slist[(size_t)b * g.slist_cap + base + lane] = wbuf[lane];
    __syncwarp();
}

__device__ __forceinline__ void s3_group_sync(int grp, int k) {
    if (k == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" :: "r"(1 + grp), "r"(k * 32) : "memory");
}

// one group of four rows against the table.  VEC: 16-byte aligned segment of a plane size that is a multiple of 4 (float4
// table loads, no masking unless EDGE); otherwise four scalar table loads per float4 and every element is range-checked.
// Returns the 4-bit row mask of this lane (bit u: its float4 of row u holds an element at or above its bound).
template <bool VEC, bool EDGE>
__device__ __forceinline__ u32 s3_test_group(const float4 *slot, const float *tab, int &tpos, int step, int plane,
                                             int rows, int e_rel0, int n_smem) {
    u32 mask = 0;
    if (VEC) {
        float4 v[S3_G], t[S3_G];
#pragma unroll
        for (int u = 0; u < S3_G; ++u) {
            v[u] = slot[u * 32];
            t[u] = ((const float4 *)tab)[tpos];
            tpos += step;
            if (tpos >= plane) tpos -= plane;
        }
#pragma unroll
        for (int u = 0; u < S3_G; ++u) {
            u32 h;
            asm("{\n\t.reg .pred p;\n\t"
                "setp.ge.f32 p, %1, %5;\n\t"
                "setp.ge.or.f32 p, %2, %6, p;\n\t"
                "setp.ge.or.f32 p, %3, %7, p;\n\t"
                "setp.ge.or.f32 p, %4, %8, p;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(h) : "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w), "f"(t[u].x), "f"(t[u].y), "f"(t[u].z), "f"(t[u].w));
            if (EDGE) h &= (u < rows && e_rel0 + 128 * u < n_smem) ? 1u : 0u;     // (n_smem is a multiple of 4 here)
            mask |= h << u;
        }
    } else {
#pragma unroll
        for (int u = 0; u < S3_G; ++u) {
            const float4 v = slot[u * 32];
            const float t0 = tab[tpos], t1 = tab[tpos + 1], t2 = tab[tpos + 2], t3 = tab[tpos + 3];
            tpos += step;
            if (tpos >= plane) tpos -= plane;
            const int el = e_rel0 + 128 * u;
            const u32 h = ((v.x >= t0) & (el >= 0) & (el < n_smem)) | ((v.y >= t1) & (el + 1 >= 0) & (el + 1 < n_smem)) |
                          ((v.z >= t2) & (el + 2 >= 0) & (el + 2 < n_smem)) | ((v.w >= t3) & (el + 3 >= 0) & (el + 3 < n_smem));
            mask |= (u < rows ? (h & 1u) : 0u) << u;
        }
    }
    return mask;
}

__global__ void __launch_bounds__(S3_NT, S3_CTAS_PER_SM)
vy_decode_stream3_kernel(const __grid_constant__ VyHeads hd, const __grid_constant__ SelPlan pl,
                         const __grid_constant__ SelGlobal g) {
    extern __shared__ __align__(128) unsigned char s3_dyn[];          // rings [S3_WARPS][S3_RING_BYTES], tables [tab_max floats]
    __shared__ __align__(8) u64 bars[S3_WARPS][S3_NSLOT];
    __shared__ u32 hq_all[S3_WARPS][S3_HQ];
    __shared__ u64 wbuf_all[S3_WARPS][64];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned char *ring = s3_dyn + (size_t)wid * S3_RING_BYTES;
    float *tabs = (float *)(s3_dyn + (size_t)S3_WARPS * S3_RING_BYTES);
    u64 *bar = bars[wid];
    u32 *hq = hq_all[wid];
    u64 *wbuf = wbuf_all[wid];
    if (lane == 0) {
        for (int i = 0; i < S3_NSLOT; ++i) s2_mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const u32 lt_mask = (1u << lane) - 1u;
    const float valid_thresh = pl.valid_thresh;
    const long long t_begin = pl.n_tiles * (long long)blockIdx.x / gridDim.x;
    const long long t_end = pl.n_tiles * (long long)(blockIdx.x + 1) / gridDim.x;
    u32 ii = 0, ic = 0;                                   // bulk copies issued / consumed by this warp (running: slot and parity)
    int qh = 0, qt = 0, cnt = 0;                          // queue head / tail (running), keys waiting in wbuf: warp-uniform
    vy_grid_dep_wait();                                   // the sample kernel's bounds, lists and counters

    for (long long t = t_begin; t < t_end; ) {            // phases: the part of the range that lies in one scale of one image
        const int pb = (int)(t / pl.tiles_per_image);
        const int pr = (int)(t - (long long)pb * pl.tiles_per_image);
        int ps = 0;
        while (ps + 1 < hd.n_scales && pr >= pl.tile_begin[ps + 1]) ++ps;
        long long pe = (long long)pb * pl.tiles_per_image + pl.tile_begin[ps] + (long long)pl.tiles_blk[ps] * hd.A;
        if (pe > t_end) pe = t_end;
        const int ngrp = pl.s3_groups[ps], k = S3_WARPS / ngrp;          // k warps per group
        const int grp = wid / k, wg = wid - grp * k;
        const long long g0 = t + (pe - t) * grp / ngrp, g1 = t + (pe - t) * (grp + 1) / ngrp;
        const VyScale &sc = hd.sc[ps];
        const int HW = sc.HW, hwp = pl.tab_hwp[ps];
        float *tab = tabs + (size_t)grp * hwp;
        __syncthreads();                                  // the table region is cut up anew
        for (long long u0 = g0; u0 < g1; ) {              // segments of this group
            int r = (int)(u0 - (long long)pb * pl.tiles_per_image) - pl.tile_begin[ps];
            const int a = r / pl.tiles_blk[ps], tt = r - a * pl.tiles_blk[ps];
            long long u1 = u0 + (pl.tiles_blk[ps] - tt);
            if (u1 > g1) u1 = g1;
            const int tl = tt + (int)(u1 - u0);           // tiles [tt, tl) of the block
            // elements of those tiles
            int e_lo, e_hi;
            if (pl.tile_tpp[ps] > 0) {                    // planes longer than a tile: tiles inside one plane
                const int tpp = pl.tile_tpp[ps];
                const int c0 = tt / tpp, p0 = tt - c0 * tpp, c1 = tl / tpp, p1 = tl - c1 * tpp;
                e_lo = c0 * HW + p0 * (S2_TILE / 4);
                e_hi = c1 * HW + p1 * (S2_TILE / 4);
            } else {
                e_lo = tt * pl.tile_ppt[ps] * HW;
                e_hi = min(tl * pl.tile_ppt[ps], hd.C) * HW;
            }
            const float *pc0 = sc.head + ((size_t)(pb * hd.A + a) * hd.P + 5) * (size_t)HW;
            const uintptr_t addr = (uintptr_t)(pc0 + e_lo);
            const int shift = (int)((addr & 15) >> 2);
            const uintptr_t src_al = addr - (uintptr_t)shift * 4;
            const int n = e_hi - e_lo;
            long long bytes = ((long long)(shift + n) * 4 + 15) & ~15LL;
            const uintptr_t end_al = ((uintptr_t)(sc.head + (size_t)hd.B * hd.A * hd.P * (size_t)HW)) & ~(uintptr_t)15;
            if (src_al + (uintptr_t)bytes > end_al) bytes = end_al > src_al ? (long long)(end_al - src_al) : 0;
            int n_smem = (int)(bytes / 4) - shift;
            n_smem = n_smem < 0 ? 0 : (n_smem > n ? n : n_smem);
            const bool vec = shift == 0 && (HW & 3) == 0;
            // this warp's rows of the segment
            const int rows_total = (int)((bytes + 511) >> 9);
            const int rw_lo = (int)((long long)rows_total * wg / k), rw_hi = (int)((long long)rows_total * (wg + 1) / k);
            const int ng = (rw_hi - rw_lo + S3_G - 1) / S3_G;
            auto issue = [&](int j) {                     // group j of this warp's rows -> slot ii % S3_NSLOT
                const long long off = (long long)(rw_lo + S3_G * j) * 512;
                const u32 slot = ii % S3_NSLOT;
#if S3_BULK
                if (lane == 0) {
                    long long nb = (long long)min(S3_G, rw_hi - (rw_lo + S3_G * j)) * 512;
                    if (nb > bytes - off) nb = bytes - off;
                    s2_mbar_expect(&bar[slot], (u32)nb);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(s2_smem(ring + slot * S3_SLOT_BYTES)), "l"((const void *)(src_al + (uintptr_t)off)),
                                    "r"((u32)nb), "r"(s2_smem(&bar[slot])) : "memory");
                }
#else
                // every lane fetches the 16 bytes it will test (a per-thread cp.async.wait_group is all the synchronisation)
                const int nr = min(S3_G, rw_hi - (rw_lo + S3_G * j));
                const long long lo = off + lane * 16;
#pragma unroll
                for (int u = 0; u < S3_G; ++u)
                    if (u < nr && lo + u * 512 < bytes)
                        cp_async16(ring + slot * S3_SLOT_BYTES + u * 512 + lane * 16, (const void *)(src_al + (uintptr_t)(lo + u * 512)));
                cp_async_commit();
#endif
                ++ii;
            };
            // the class planes do not depend on the table: get them moving first
#if S3_BULK
            for (int j = 0; j < S3_NSLOT && j < ng; ++j) issue(j);
#else
            for (int j = 0; j < S3_NSLOT; ++j) { if (j < ng) issue(j); else { cp_async_commit(); ++ii; } }
#endif
            // ---- the group's table: bound of every position (objectness plane = the plane right below class plane 0),
            // positions 0 .. 2 again behind the last one (a shifted float4 may wrap), +inf padding
            u64 thr;
            {
                u64 m = lane < g.Gs ? g.sslots[(size_t)pb * g.Gs + lane] : 0ull;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) { const u64 o = sel_shfl_xor_u64(m, off); m = o > m ? o : m; }
                thr = ~m;
            }
            const float smin = fmaxf(thr ? vy_key_score(thr) : valid_thresh, valid_thresh);
            s3_group_sync(grp, k);                        // (the group's previous table is no longer in use)
            {
                const int gt = wg * 32 + lane, nthr = k * 32;
                const float *obj = pc0 - HW;
                for (int p0 = gt; p0 < HW; p0 += 8 * nthr) {
                    float to[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) { const int pos = p0 + q * nthr; to[q] = pos < HW ? vy_ldg32(obj + pos) : 0.0f; }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int pos = p0 + q * nthr;
                        if (pos < HW) {
                            const float v = vy_tcmin(smin, vy_sigmoid(to[q]));
                            tab[pos] = v;
                            if (pos < 3) tab[HW + pos] = v;
                        }
                    }
                }
                for (int pos = HW + 3 + gt; pos < hwp; pos += nthr) tab[pos] = CUDART_INF_F;
                if (HW < 3 && gt == 0) for (int pos = HW; pos < HW + 3; ++pos) tab[pos] = CUDART_INF_F;
            }
            s3_group_sync(grp, k);
            // ---- stream this warp's rows
            // table index of this lane's float4 in the warp's first row, and what a row further means
            int plane, step, tpos;
            {
                const long long e_first = (long long)e_lo - shift + 128LL * rw_lo + 4 * lane;      // element of v.x (may be < 0 in row 0)
                if (vec) { plane = HW >> 2; step = 32 % plane; tpos = (int)((e_first >> 2) % plane); }
                else { plane = HW; step = 128 % plane; tpos = (int)(((e_first % plane) + plane) % plane); }
            }
            for (int j = 0; j < ng; ++j) {
                const u32 slot = ic % S3_NSLOT;
#if S3_BULK
                s2_mbar_wait(&bar[slot], (ic / S3_NSLOT) & 1u);
#else
                cp_async_wait<S3_NSLOT - 1>();
#endif
                ++ic;
                const float4 *sl = (const float4 *)(ring + slot * S3_SLOT_BYTES) + lane;
                const int row0 = rw_lo + S3_G * j;                                   // first row of the group (segment-relative)
                const int rows = min(S3_G, rw_hi - row0);
                const int e_rel0 = 128 * row0 + 4 * lane - shift;                      // element (relative to e_lo) of this lane's float4 in row 0 of the group
                const bool edge = rows < S3_G || 128 * (row0 + S3_G) - shift > n_smem;
                u32 mask;
                if (vec) {
                    if (!edge) mask = s3_test_group<true, false>(sl, tab, tpos, step, plane, rows, e_rel0, n_smem);
                    else mask = s3_test_group<true, true>(sl, tab, tpos, step, plane, rows, e_rel0, n_smem);
                } else {
                    mask = s3_test_group<false, true>(sl, tab, tpos, step, plane, rows, e_rel0, n_smem);
                }
                // the slot is free as soon as it has been tested: refill first, then look after the hits
#if S3_BULK
                __syncwarp();
                if (j + S3_NSLOT < ng) issue(j + S3_NSLOT);
#else
                if (j + S3_NSLOT < ng) issue(j + S3_NSLOT); else { cp_async_commit(); ++ii; }      // (keeps the group count uniform)
#endif
                u32 rows_hit = __reduce_or_sync(0xffffffffu, mask);
                while (rows_hit) {                        // rare: one round per ROW that holds a flagged float4
                    const int u = __ffs(rows_hit) - 1;
                    rows_hit &= rows_hit - 1;
                    const bool has = (mask >> u) & 1u;
                    const u32 bal = __ballot_sync(0xffffffffu, has);
                    if (has) {
                        const int el = e_rel0 + 128 * u;
                        u32 m = 0xfu;
                        if (!vec || edge)
                            m = ((el >= 0 && el < n_smem) ? 1u : 0u) | ((el + 1 >= 0 && el + 1 < n_smem) ? 2u : 0u) |
                                ((el + 2 >= 0 && el + 2 < n_smem) ? 4u : 0u) | ((el + 3 >= 0 && el + 3 < n_smem) ? 8u : 0u);
                        hq[(qt + __popc(bal & lt_mask)) & (S3_HQ - 1)] = (u32)(e_lo + el + 3) | (m << 28);
                    }
                    qt += __popc(bal);
                    __syncwarp();
                    if (qt - qh >= 32) {
                        cnt = s3_score(hd, valid_thresh, pb, ps, a, thr, hq, qh, 32, wbuf, cnt, g);
                        qh += 32;
                    }
                }
            }
            // the tensor's last <= 3 floats when it ends off a 16-byte boundary: never in shared memory
            if (n_smem < n && wg == k - 1) {
                const int el = n_smem + lane;
                const bool hit = lane < n - n_smem && vy_ldg32(pc0 + e_lo + el) >= tab[(e_lo + el) % HW];
                const u32 bal = __ballot_sync(0xffffffffu, hit);
                if (hit) hq[(qt + __popc(bal & lt_mask)) & (S3_HQ - 1)] = (u32)(e_lo + el + 3) | (1u << 28);
                qt += __popc(bal);
                __syncwarp();
            }
            // end of the segment: queue entries name elements of THIS block, keys go to THIS image's list
            while (qt != qh) {
                const int nb = min(32, qt - qh);
                cnt = s3_score(hd, valid_thresh, pb, ps, a, thr, hq, qh, nb, wbuf, cnt, g);
                qh += nb;
            }
            if (cnt > 0) { s3_flush_keys(pb, wbuf, cnt, g); cnt = 0; }
            u0 = u1;
        }
        t = pe;
    }
    vy_grid_dep_trigger();
}

