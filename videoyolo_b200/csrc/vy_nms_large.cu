// vy_nms_large.cu -- box_nms when more than SEL_KMAX candidates take part (topk < 0 or large).
#include "vy_select.cuh"

size_t vy_box_nms_large_workspace_bytes(int B, long long R, int W_elem) {
    (void)B; (void)R; (void)W_elem;
    return 256;
}

int vy_box_nms_large(const RowParams &rp, int B, long long K, float overlap_thresh, int force_suppress,
                     int in_format, int out_format, long long out_rows, float *out, int32_t *kept_rows,
                     void *workspace, size_t workspace_bytes, cudaStream_t st) {
    (void)rp; (void)B; (void)overlap_thresh; (void)force_suppress; (void)in_format; (void)out_format;
    (void)out_rows; (void)out; (void)kept_rows; (void)workspace; (void)workspace_bytes; (void)st;
    VY_FAIL(VY_EUNSUPPORTED, "box_nms with %lld > %d participating candidates is not built yet", K, SEL_KMAX);
}
