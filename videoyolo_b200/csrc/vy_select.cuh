// vy_select.cuh -- CTA-level streaming top-K selection over 64-bit keys.
//
// A CTA streams candidates past a monotonically rising threshold key.  Candidates that beat the
// threshold are pushed into a shared-memory buffer; when the buffer fills, an MSB-first radix
// select finds a (bucket-granular, or exact) K-th largest key, everything below it is dropped
// and the threshold rises.  Any threshold produced this way is a valid lower bound of the
// image-wide K-th largest key, so thresholds can be shared between CTAs of one image through a
// global atomicMax without losing a single top-K candidate.
#pragma once
#include "vy_common.cuh"

constexpr int SEL_NT  = 256;               // threads per selecting CTA
constexpr int SEL_CAP = 2048;              // candidate keys held in shared memory
constexpr int SEL_KPT = SEL_CAP / SEL_NT;  // keys per thread during a compaction
constexpr int SEL_KMAX = 1024;             // largest K the small (shared-memory) path serves

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
    u64 thr;            // inclusive lower bound: a candidate is kept iff key >= thr
    int count;          // number of pushes since the last reset (may exceed SEL_CAP: overflow)
    int sel_digit, sel_above, sel_in;
    int flag;           // general CTA-uniform scratch
    int snap;           // count snapshot written by thread 0 between barriers (pushes never touch it)
};

__device__ __forceinline__ void sel_reset(SelBuf &S) {
    if (threadIdx.x == 0) { S.count = 0; S.thr = 0; }
}

// push; keys beyond the capacity are dropped but still counted (the caller detects count > CAP)
__device__ __forceinline__ void sel_push(SelBuf &S, u64 key) {
    const int slot = atomicAdd(&S.count, 1);
    if (slot < SEL_CAP) S.keys[slot] = key;
}

// Keep (about) the K largest keys of S.keys[0..count).  exact=false stops as soon as at most
// K + K/4 keys remain (bucket granularity); exact=true keeps exactly K.  Raises S.thr.
// Preconditions: n == S.count <= SEL_CAP is CTA-uniform, all pushes visible and no push in flight
// (__syncthreads before the call), called by every thread of a CTA of >= SEL_NT threads.
// Returns the new count (CTA-uniform).  Ends with a __syncthreads.
static __device__ int sel_compact(SelBuf &S, int n, int K, bool exact) {
    const int tid = threadIdx.x;
    if (n <= K) return n;                     // CTA-uniform
    u64 my[SEL_KPT];
#pragma unroll
    for (int i = 0; i < SEL_KPT; ++i) {
        const int idx = tid + i * SEL_NT;
        my[i] = (tid < SEL_NT && idx < n) ? S.keys[idx] : 0ull;
    }
    u64 prefix = 0;
    int kk = K;                               // looking for the kk-th largest inside the prefix bucket
    int kept = n;
    const int slack = exact ? 0 : (K >> 2);
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
        kept = (K - kk) + inb;                // keys >= prefix (low bits zero)
        if (kept <= K + slack) break;
    }
    __syncthreads();
    if (tid == 0) { S.count = 0; if (prefix > S.thr) S.thr = prefix; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SEL_KPT; ++i)
        if (my[i] >= prefix && my[i] != 0ull) { const int slot = atomicAdd(&S.count, 1); S.keys[slot] = my[i]; }
    __syncthreads();
    return kept;
}

// In-place descending bitonic sort of S.keys[0..npow2) (npow2 a power of two <= SEL_CAP; the
// caller pads with zeros).  Every thread of the CTA calls it.
static __device__ void sel_sort_desc(SelBuf &S, int npow2) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = tid; p < (npow2 >> 1); p += nt) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const int l = i | j;
                const u64 a = S.keys[i], b = S.keys[l];
                const bool desc = (i & k) == 0;
                if ((a < b) == desc) { S.keys[i] = b; S.keys[l] = a; }
            }
            __syncthreads();
        }
    }
}
