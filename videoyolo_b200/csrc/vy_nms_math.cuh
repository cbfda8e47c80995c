// vy_nms_math.cuh -- the IoU arithmetic of MXNet _contrib_box_nms (BoxArea / Intersect in
// src/operator/contrib/bounding_box-inl.h, restated in SURVEY.md App. B and oracle/vy_oracle.c),
// un-contracted IEEE fp32 in exactly the oracle's association order, plus the warp-level greedy scan
// over a suppression bitmask that both NMS kernels (vy_nms.cu finalize, vy_nms_large.cu tiles) use.
#pragma once
#include "vy_common.cuh"

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

// iou(ref, pos) > thr, exactly: inter / (area_ref + area_pos - inter) in IEEE fp32 (strict compare,
// 0/0 = NaN does not suppress).
__device__ __forceinline__ bool nms_suppresses(float4 r, float ar, float4 p, float ap, float thr, int fmt) {
    float inter = nms_isect(r.x, r.z, p.x, p.z, fmt);
    inter = __fmul_rn(inter, nms_isect(r.y, r.w, p.y, p.w, fmt));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ar, ap), inter)) > thr;
}

// Same predicate, but the IEEE division only runs when a 2-ulp reciprocal estimate lands within
// 2^-20 relative of the threshold (or the denominator leaves the range in which the estimate is
// specified): the estimate q satisfies |q - x| <= 2^-21 |x| for the true quotient x, so
//   q > thr (1 + 2^-20)  ==>  x > thr (1 + 2^-22)  ==>  RN(x) >= thr + 1 ulp > thr
//   q < thr (1 - 2^-20)  ==>  x < thr (1 - 2^-22)  ==>  RN(x) <= thr - 1 ulp < thr      (thr > 0)
// and the result is bit-identical to nms_suppresses.  thr_hi/thr_lo are those two products
// (caller passes thr_lo = +inf, thr_hi = -inf to force the exact path, e.g. for thr <= 0).
__device__ __forceinline__ bool nms_suppresses_fast(float4 r, float ar, float4 p, float ap, float thr,
                                                    float thr_lo, float thr_hi, int fmt) {
    float inter = nms_isect(r.x, r.z, p.x, p.z, fmt);
    inter = __fmul_rn(inter, nms_isect(r.y, r.w, p.y, p.w, fmt));
    const float u = __fsub_rn(__fadd_rn(ar, ap), inter);
    const float q = __fdividef(inter, u);
    const bool in_range = u > 1e-30f && u < 1e30f && inter < 1e30f;
    if (in_range && q < thr_lo) return false;
    if (in_range && q > thr_hi) return true;
    return __fdiv_rn(inter, u) > thr;
}

// Greedy scan by ONE warp over n <= 1024 entries in order.  mask[i * stride + w] holds, for reference
// i, the bits of the later entries (word w >= i/32) it would suppress; rowany[w] marks the entries of
// word w whose mask row is non-empty.  Words [i/32, (n-1)/32] of every marked row must be valid.
// Result: keepw[w] = survivors of word w.  Lane w carries the removed-bits of word w.
__device__ __forceinline__ void nms_greedy_scan_warp(const u32 *mask, int stride, const u32 *rowany, int n,
                                                     u32 *keepw, int lane) {
    const int nw = (n + 31) >> 5;
    u32 removed = 0;
    for (int blk = 0; blk < nw; ++blk) {
        const int r = (blk << 5) + lane;
        const u32 validm = (n - (blk << 5) >= 32) ? 0xffffffffu : ((1u << (n - (blk << 5))) - 1u);
        const u32 ra = rowany[blk] & validm;
        const u32 diag = (r < n && ((ra >> lane) & 1u)) ? mask[(size_t)r * stride + blk] : 0u;
        u32 rem = __shfl_sync(0xffffffffu, removed, blk);
        u32 pend = ra;
        while (pend) {
            const int i = __ffs(pend) - 1;
            pend &= pend - 1;
            const u32 di = __shfl_sync(0xffffffffu, diag, i);
            if (!((rem >> i) & 1u)) rem |= di;
        }
        const u32 keep = ~rem & validm;
        if (lane == 0) keepw[blk] = keep;
        u32 act = keep & ra;                            // surviving references reach into later words
        if (lane > blk && lane < nw) {
            u32 acc = 0;
            while (act) {
                const int i = __ffs(act) - 1;
                act &= act - 1;
                acc |= mask[(size_t)((blk << 5) + i) * stride + lane];
            }
            removed |= acc;
        }
    }
}
