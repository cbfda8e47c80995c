// vy_decode.cu -- anchor decode into the reference's materialised detection tensor.
// Replaces YOLOOutputV3.hybrid_forward (yolo3.py:151-199) for all scales + the concat (yolo3.py:523).
//
// One thread per box (pos, a), anchor fastest, so that consecutive lanes write consecutive 24-byte
// rows of the class block: the kernel is bound by the C-times-replicated output (6*C floats written
// per 5+C floats read) and its stores stream out as contiguous 768-byte warp segments.  The box and
// the objectness stay in registers across the class loop.
#include "vy_common.cuh"

constexpr int DEC_NT = 256;

__global__ void __launch_bounds__(DEC_NT)
vy_decode_kernel(VyHeads hd, float *__restrict__ dets, int s, int blocks_per_image) {
    const VyScale &sc = hd.sc[s];
    const int b = blockIdx.x / blocks_per_image;
    const long long e = (long long)(blockIdx.x % blocks_per_image) * DEC_NT + threadIdx.x;   // pos*A + a
    if (e >= sc.n_s) return;
    const int pos = (int)(e / hd.A), a = (int)(e % hd.A);
    const int y = pos / sc.W, x = pos % sc.W;
    const size_t HW = (size_t)sc.HW;
    const float *p = sc.head + ((size_t)(b * hd.A + a) * hd.P) * HW + pos;
    const float4 bx = vy_box(__ldg(p), __ldg(p + HW), __ldg(p + 2 * HW), __ldg(p + 3 * HW), x, y,
                             sc.stride, sc.aw[a], sc.ah[a]);
    const float conf = vy_sigmoid(__ldg(p + 4 * HW));
    float *out = dets + ((size_t)b * (size_t)hd.R + (size_t)sc.row_off + (size_t)e) * 6;
    if (hd.agnostic) {                                    // yolo3.py:184-188
        float2 *o = (float2 *)out;
        o[0] = make_float2(0.0f, conf); o[1] = make_float2(bx.x, bx.y); o[2] = make_float2(bx.z, bx.w);
        return;
    }
    const size_t cls_stride = (size_t)sc.n_s * 6;
    const float *pc = p + 5 * HW;
#pragma unroll 4
    for (int c = 0; c < hd.C; ++c) {                      // yolo3.py:175, 191-197
        const float sc_c = vy_score(__ldg(pc + (size_t)c * HW), conf);
        float2 *o = (float2 *)(out + (size_t)c * cls_stride);
        __stcs(o + 0, make_float2((float)c, sc_c));
        __stcs(o + 1, make_float2(bx.x, bx.y));
        __stcs(o + 2, make_float2(bx.z, bx.w));
    }
}

extern "C" int vy_decode_f32(const float *const *head, const int *H, const int *W, const float *stride,
                             const float *anchors, int n_scales, int B, int A, int C, int agnostic,
                             float *dets, vy_stream_t stream) {
    VyHeads hd;
    const int rc = vy_fill_heads(&hd, head, H, W, stride, anchors, n_scales, B, A, C, agnostic);
    if (rc != VY_OK) return rc;
    if (!dets) VY_FAIL(VY_EINVAL, "vy_decode_f32: dets is null");
    if (((uintptr_t)dets & 7) != 0) VY_FAIL(VY_EALIGN, "vy_decode_f32: dets must be 8-byte aligned");
    for (int s = 0; s < n_scales; ++s) {
        if (!head[s]) VY_FAIL(VY_EINVAL, "vy_decode_f32: head[%d] is null", s);
        const int bpi = (int)((hd.sc[s].n_s + DEC_NT - 1) / DEC_NT);
        const long long grid = (long long)bpi * B;
        if (grid > 0x7fffffffLL) VY_FAIL(VY_EINVAL, "vy_decode_f32: grid too large");
        VY_KERNEL(VY_K_DECODE, (cudaStream_t)stream,
                  vy_decode_kernel<<<(unsigned)grid, DEC_NT, 0, (cudaStream_t)stream>>>(hd, dets, s, bpi));
        VY_LAUNCH_CHECK("vy_decode_kernel");
    }
    return VY_OK;
}
