// vy_select.cuh -- CTA-level streaming top-K selection over 64-bit keys.
//
// A CTA streams candidates past a monotonically rising threshold key.  Candidates that beat the
// threshold are pushed into a shared-memory buffer.  From time to time the CTA runs an MSB-first
// radix select over the buffer to (a) bound its own K-th largest key and (b) bound its own
// ceil(K/G)-th largest key, which it publishes in a per-image slot.  G CTAs share one image:
//      min over the G slots      -- each CTA holds >= ceil(K/G) candidates at or above its slot, so
//                                   together they hold >= K candidates at or above the minimum
//      a CTA's own K-th largest  -- trivially
// are both valid lower bounds of the image-wide K-th largest key.  Every key below a valid bound
// can be dropped without losing a single top-K candidate; the best bound known for the image is
// kept in global memory (atomicMax) and picked up by every CTA working on that image.
#pragma once
#include "vy_common.cuh"

constexpr int SEL_NT  = 256;               // threads per selecting CTA
constexpr int SEL_CAP = 2048;              // candidate keys held in shared memory
constexpr int SEL_KPT = SEL_CAP / SEL_NT;  // keys per thread during a selection
constexpr int SEL_KMAX = 1024;             // largest K the small (shared-memory) path serves
constexpr int SEL_GMAX = 32;               // CTAs sharing one image (one slot per lane of a warp)
constexpr int SEL_QCAP = 2048;             // prefilter hits queued per streamed block

// a materialised detection tensor (the operand of F.contrib.box_nms, yolo3.py:526)
struct RowParams {
    const float *data;           // (B, R, W)
    long long R;
    int W, coord_start, score_index, id_index, background_id;
    float valid_thresh;
};

struct SelBuf {
    u64 keys[SEL_CAP];
    u32 hist[256];
    // streaming threads only record WHERE a prefilter hit happened; the exact evaluation is done
    // afterwards by all threads of the CTA side by side (one queued hit per thread), so a hit costs
    // a few thread-instructions instead of a divergent warp-long detour
    u32 queue[SEL_QCAP];
    float it_conf[4][SEL_NT];   // per streaming thread: objectness of its 4 boxes
    u32 it_off[SEL_NT];         // element offset of its (b, a, pos0) in the scale's head map
    u32 it_row0[SEL_NT];        // reference row of (c=0, pos0, a)
    u32 it_scale[SEL_NT];
    int qcount;                 // hits queued since the last reset (may exceed SEL_QCAP: overflow)
    u64 thr;            // inclusive lower bound: a candidate is kept iff key >= thr
    u64 slot_pub;       // what this CTA last published in its slot
    int count;          // number of pushes since the last reset (may exceed SEL_CAP: overflow)
    int sel_digit, sel_above, sel_in;
    int flag;           // CTA-uniform scratch
    int snap;           // count snapshot written by thread 0 between barriers (pushes never touch it)
};

// per-job view of the global sharing state
struct SelJob {
    u64 *g_thr_b;       // best valid bound known for the image
    u64 *slots_b;       // [G] per-CTA ceil(K/G)-th largest keys
    int G, g, K, Kq;    // Kq = ceil(K / G)
};

__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// push; keys beyond the capacity are dropped but still counted (the caller detects count > CAP)
__device__ __forceinline__ void sel_push(SelBuf &S, u64 key) {
    const int slot = atomicAdd(&S.count, 1);
    if (slot < SEL_CAP) S.keys[slot] = key;
}

// Lower bound of the `rank`-th largest key of S.keys[0..n): the returned prefix p satisfies
// rank <= #{keys >= p} <= rank + slack (slack == 0: p IS the rank-th largest key).
// Preconditions: 1 <= rank <= n <= SEL_CAP, n CTA-uniform, no push in flight; called by every
// thread of a CTA of >= SEL_NT threads.  Contains barriers; S.keys is not modified.
static __device__ __noinline__ u64 sel_rank_bound(SelBuf &S, int n, int rank, int slack) {
    const int tid = threadIdx.x;
    u64 my[SEL_KPT];
#pragma unroll
    for (int i = 0; i < SEL_KPT; ++i) {
        const int idx = tid + i * SEL_NT;
        my[i] = (tid < SEL_NT && idx < n) ? S.keys[idx] : 0ull;
    }
    u64 prefix = 0;
    int kk = rank;                            // looking for the kk-th largest inside the prefix bucket
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (tid < 256) S.hist[tid] = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < SEL_KPT; ++i) {
            const u64 k = my[i];
            const bool in = (k != 0ull) && (shift == 56 || ((k ^ prefix) >> (shift + 8)) == 0ull);
            if (in) atomicAdd(&S.hist[(u32)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            // lane l owns bins 255-8l .. 248-8l (descending); find the bin where the running
            // count from the top crosses kk
            u32 c[8], s = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) { c[t] = S.hist[255 - 8 * tid - t]; s += c[t]; }
            u32 inc = s;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const u32 v = __shfl_up_sync(0xffffffffu, inc, off);
                if (tid >= off) inc += v;
            }
            const u32 exc = inc - s;
            if (exc < (u32)kk && (u32)kk <= inc) {
                u32 run = exc;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (run + c[t] >= (u32)kk) {
                        S.sel_digit = 255 - 8 * tid - t; S.sel_above = (int)run; S.sel_in = (int)c[t];
                        break;
                    }
                    run += c[t];
                }
            }
        }
        __syncthreads();
        const int d = S.sel_digit, above = S.sel_above, inb = S.sel_in;
        prefix |= (u64)d << shift;
        kk -= above;                          // 1 <= kk <= inb
        if (inb - kk <= slack) break;         // #{keys >= prefix} = rank + (inb - kk)
    }
    return prefix;
}

// Drop every key below thr, compacting the survivors to the front.  Returns the new count
// (CTA-uniform).  Same preconditions as sel_rank_bound; ends with a barrier.
static __device__ __noinline__ int sel_drop_below(SelBuf &S, int n, u64 thr) {
    const int tid = threadIdx.x;
    u64 my[SEL_KPT];
#pragma unroll
    for (int i = 0; i < SEL_KPT; ++i) {
        const int idx = tid + i * SEL_NT;
        my[i] = (tid < SEL_NT && idx < n) ? S.keys[idx] : 0ull;
    }
    __syncthreads();
    if (tid == 0) S.count = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SEL_KPT; ++i)
        if (my[i] >= thr && my[i] != 0ull) { const int slot = atomicAdd(&S.count, 1); S.keys[slot] = my[i]; }
    __syncthreads();
    const int m = S.count;
    __syncthreads();
    return m;
}

// Keep (about) the K largest keys: exact=false leaves at most K + K/4, exact=true exactly K.
// Raises S.thr.  Returns the new count.
static __device__ int sel_compact(SelBuf &S, int n, int K, bool exact) {
    if (n <= K) return n;
    const u64 p = sel_rank_bound(S, n, K, exact ? 0 : (K >> 2));
    __syncthreads();
    if (threadIdx.x == 0) {                            // (volatile: the read must not be hoisted above the predicate)
        const u64 cur = *(volatile u64 *)&S.thr;
        if (p > cur) S.thr = p;
    }
    return sel_drop_below(S, n, p);
}

// Selection event: refresh this CTA's slot, combine every bound known for the image, drop what
// fell below.  `must_free`: the buffer overflowed, make room (exact local K-th).  Returns count.
static __device__ __noinline__ int sel_update(SelBuf &S, int n, const SelJob &jb, bool must_free) {
    const int tid = threadIdx.x;
    u64 vq = 0, vk = 0;
    if (n >= jb.Kq) vq = sel_rank_bound(S, n, jb.Kq, jb.Kq >> 2);
    if (n > jb.K + (jb.K >> 2) || (must_free && n > jb.K))
        vk = sel_rank_bound(S, n, jb.K, must_free ? 0 : (jb.K >> 2));
    __syncthreads();
    if (tid < 32) {
        const u64 old = S.slot_pub;
        const u64 mine = vq > old ? vq : old;
        __syncwarp();
        if (tid == 0 && mine > old) { S.slot_pub = mine; st_relaxed_u64(jb.slots_b + jb.g, mine); }
        u64 v = ~0ull;
        if (tid < jb.G) v = (tid == jb.g) ? mine : ld_relaxed_u64(jb.slots_b + tid);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const u64 o = __shfl_xor_sync(0xffffffffu, v, off);
            v = o < v ? o : v;
        }
        if (tid == 0) {
            const u64 gthr = ld_relaxed_u64(jb.g_thr_b);
            u64 t = S.thr;
            if (vk > t) t = vk;
            if (v > t) t = v;                 // v == 0 while some CTA has not published yet
            if (t > gthr) atomicMax(jb.g_thr_b, t);
            if (gthr > t) t = gthr;
            S.flag = t > S.thr;
            S.thr = t;
        }
    }
    __syncthreads();
    if (S.flag) n = sel_drop_below(S, n, S.thr);
    return n;
}

// In-place descending bitonic sort of keys[0..npow2) in shared memory (npow2 a power of two; the
// caller pads with zeros).  Every thread of the CTA calls it.
static __device__ __noinline__ void sel_sort_desc(u64 *keys, int npow2) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = tid; p < (npow2 >> 1); p += nt) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const int l = i | j;
                const u64 a = keys[i], b = keys[l];
                const bool desc = (i & k) == 0;
                if ((a < b) == desc) { keys[i] = b; keys[l] = a; }
            }
            __syncthreads();
        }
    }
}

// Same result as sel_sort_desc for npow2 <= 2 * blockDim.x, with ~4x fewer CTA barriers: thread t owns the
// adjacent elements 2t and 2t+1 in registers; compare-exchange distances 1..32 stay inside the thread /
// the warp (shuffles), only distances >= 64 go through shared memory.
__device__ __forceinline__ u64 sel_shfl_xor_u64(u64 v, int lane_mask) {
    const u32 lo = __shfl_xor_sync(0xffffffffu, (u32)v, lane_mask);
    const u32 hi = __shfl_xor_sync(0xffffffffu, (u32)(v >> 32), lane_mask);
    return ((u64)hi << 32) | lo;
}
static __device__ __noinline__ void sel_sort_desc_fast(u64 *keys, int npow2) {
    const int tid = threadIdx.x;
    if (npow2 > 2 * (int)blockDim.x || npow2 < 64) { sel_sort_desc(keys, npow2); return; }
    const bool act = 2 * tid < npow2;                 // whole warps: npow2 >= 64
    const int p0 = 2 * tid;
    for (int k = 2; k <= npow2; k <<= 1) {
        int j = k >> 1;
        for (; j >= 64; j >>= 1) {                    // shared-memory stages
            for (int p = tid; p < (npow2 >> 1); p += blockDim.x) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const int l = i | j;
                const u64 a = keys[i], b = keys[l];
                const bool desc = (i & k) == 0;
                if ((a < b) == desc) { keys[i] = b; keys[l] = a; }
            }
            __syncthreads();
        }
        if (act) {                                    // distances 32 .. 1 in registers
            u64 a = keys[p0], b = keys[p0 + 1];
            for (; j >= 2; j >>= 1) {
                const u64 ya = sel_shfl_xor_u64(a, j >> 1), yb = sel_shfl_xor_u64(b, j >> 1);
                const bool lower = (p0 & j) == 0;     // both of my elements sit on the same side
                const bool desc = ((p0 & ~j) & k) == 0;
                const bool keep_max = lower == desc;
                a = keep_max ? (a > ya ? a : ya) : (a < ya ? a : ya);
                b = keep_max ? (b > yb ? b : yb) : (b < yb ? b : yb);
            }
            {                                         // distance 1: my own pair
                const bool desc = (p0 & k) == 0;
                if ((a < b) == desc) { const u64 t = a; a = b; b = t; }
            }
            keys[p0] = a; keys[p0 + 1] = b;
        }
        __syncthreads();
    }
}

// Descending sort of keys[0..npow2), 32 <= npow2 <= blockDim.x, as a chain of short phases instead of a
// bitonic network's log^2 dependent stages: every warp sorts its run of 32 keys in registers (15 shuffle
// stages), then rounds of 4-way (last round possibly 2-way) RANK merges: a key's place in the merged run
// is its place in its own run plus, for every sibling run, the number of keys there that precede it --
// log2(L)+1 probes of a binary search each, the searches of the sibling runs independent of one another.
// Equal keys are ordered by the run they came from (sibling runs to the left count their equals too), so
// the places are a permutation whatever the input.  2 CTA barriers per round; every thread of the CTA calls it.
#ifdef VY_FIN_TIMING
__device__ long long vy_fin_clk[16];
#endif
// While the rounds run, element p of the buffer lives at p ^ ((p >> 4) & 15): the probes of one binary-search
// step sit a multiple of 16 keys apart (one shared-memory bank pair), the swizzle spreads them over the banks.
__device__ __forceinline__ int sel_swz(int p) { return p ^ ((p >> 4) & 15); }
static __device__ __noinline__ void sel_sort_desc_merge(u64 *keys, int npow2) {
    const int tid = threadIdx.x, lane = tid & 31;
    const bool act = tid < npow2;                      // whole warps
    // the register phase runs in every warp (idle ones sort zeros): shuffles under a branch the compiler cannot
    // prove warp-uniform cost a collective sequence each
    u64 x = act ? keys[tid] : 0ull;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const u64 y = sel_shfl_xor_u64(x, j);
            const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
            x = keep_max ? (x > y ? x : y) : (x < y ? x : y);
        }
    }
    if (npow2 == 32) {
        if (act) keys[tid] = x;
        __syncthreads();
        return;
    }
    __syncwarp();                                      // the swizzle permutes inside a warp: every lane has read its key
    if (act) keys[sel_swz(tid)] = x;
    __syncthreads();
#ifdef VY_FIN_TIMING
    int tslot = 12;
    if (blockIdx.x == 0 && tid == 0) vy_fin_clk[tslot] = clock64();
#endif
    for (int L = 32; L < npow2;) {
        const int F = (L * 4 <= npow2) ? 4 : 2;
        const int G = L * F;
        int pos = 0;
        if (act) {
            const int g0 = tid & ~(G - 1);
            const int myrun = (tid - g0) / L;          // warp-uniform (L >= 32)
            int cnt[3] = {0, 0, 0};
            int r[3];
            u64 xq[3];
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                int q = myrun + 1 + o;
                if (q >= F) q -= F;
                r[o] = g0 + q * L;
                xq[o] = x - (u64)(q < myrun ? 1 : 0);   // runs to the left: their equals precede me (x >= 1)
            }
            const int no = F - 1;
            for (int s = L >> 1; s > 0; s >>= 1) {
#pragma unroll
                for (int o = 0; o < 3; ++o)
                    if (o < no && keys[sel_swz(r[o] + cnt[o] + s - 1)] > xq[o]) cnt[o] += s;
            }
#pragma unroll
            for (int o = 0; o < 3; ++o)
                if (o < no && keys[sel_swz(r[o] + cnt[o])] > xq[o]) ++cnt[o];
            pos = g0 + (tid & (L - 1));
#pragma unroll
            for (int o = 0; o < 3; ++o) if (o < no) pos += cnt[o];
        }
        const bool last = G >= npow2;                  // the last round leaves the keys in natural order
        __syncthreads();
        if (act) keys[last ? pos : sel_swz(pos)] = x;
        __syncthreads();
        if (act && !last) x = keys[sel_swz(tid)];
        L = G;
#ifdef VY_FIN_TIMING
        ++tslot;
        if (blockIdx.x == 0 && tid == 0 && tslot < 16) vy_fin_clk[tslot] = clock64();
#endif
    }
}
