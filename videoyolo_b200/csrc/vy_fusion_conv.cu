// vy_fusion_conv.cu -- temporal fusion convolution (placeholder entry points until the tcgen05 kernel lands).
#include "vy_common.cuh"

extern "C" size_t vy_fusion_conv_workspace_bytes(int B, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw) {
    (void)B; (void)T; (void)H; (void)W; (void)Cin; (void)Cout; (void)kt; (void)kh; (void)kw;
    return 256;
}

extern "C" int vy_fusion_conv_bf16(const void *x, const void *w, const float *scale, const float *shift,
                                   float leaky_slope, int B, int T, int H, int W, int Cin, int Cout,
                                   int kt, int kh, int kw, void *y, int y_is_f32, void *workspace,
                                   size_t workspace_bytes, vy_stream_t stream) {
    (void)x; (void)w; (void)scale; (void)shift; (void)leaky_slope; (void)B; (void)T; (void)H; (void)W;
    (void)Cin; (void)Cout; (void)kt; (void)kh; (void)kw; (void)y; (void)y_is_f32; (void)workspace;
    (void)workspace_bytes; (void)stream;
    VY_FAIL(VY_EUNSUPPORTED, "vy_fusion_conv_bf16: kernel not built yet");
}

extern "C" int vy_temporal_pool_bf16(const void *x, int B, int T, long inner, int mode, void *y, vy_stream_t stream) {
    (void)x; (void)B; (void)T; (void)inner; (void)mode; (void)y; (void)stream;
    VY_FAIL(VY_EUNSUPPORTED, "vy_temporal_pool_bf16: kernel not built yet");
}
