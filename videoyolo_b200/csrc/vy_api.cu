// vy_api.cu -- error plumbing, device queries and head-map description shared by all entry points.
#include "vy_common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void vy_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int vy_version(void) { return VY_ABI_VERSION; }
extern "C" const char *vy_last_error(void) { return g_err; }

int vy_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int vy_fill_heads(VyHeads *h, const float *const *head, const int *H, const int *W, const float *stride,
                  const float *anchors, int n_scales, int B, int A, int C, int agnostic) {
    if (!h || !head || !H || !W || !stride || !anchors) VY_FAIL(VY_EINVAL, "null host array");
    if (n_scales < 1 || n_scales > VY_MAX_SCALES) VY_FAIL(VY_EINVAL, "n_scales=%d outside [1,%d]", n_scales, VY_MAX_SCALES);
    if (A < 1 || A > VY_MAX_ANCHORS) VY_FAIL(VY_EINVAL, "A=%d outside [1,%d]", A, VY_MAX_ANCHORS);
    if (B < 1 || C < 1) VY_FAIL(VY_EINVAL, "B=%d, C=%d must be >= 1", B, C);
    memset(h, 0, sizeof(*h));
    h->n_scales = n_scales; h->B = B; h->A = A; h->C = C; h->P = 5 + C;
    h->agnostic = agnostic ? 1 : 0;
    h->Ceff = agnostic ? 1 : C;
    long long boxes = 0;
    for (int s = 0; s < n_scales; ++s) {
        VyScale &sc = h->sc[s];
        // alloc_size=(128,128) caps the reference's offset map (yolo3.py:44,67-74)
        if (H[s] < 1 || W[s] < 1 || H[s] > 4096 || W[s] > 4096) VY_FAIL(VY_EINVAL, "bad H/W at scale %d", s);
        sc.head = head[s];
        sc.H = H[s]; sc.W = W[s]; sc.HW = H[s] * W[s];
        sc.stride = stride[s];
        for (int a = 0; a < A; ++a) {
            sc.aw[a] = anchors[(s * A + a) * 2 + 0];
            sc.ah[a] = anchors[(s * A + a) * 2 + 1];
        }
        sc.vec = (sc.HW % 4 == 0 && (((uintptr_t)head[s]) & 15) == 0) ? 4 : 1;
        sc.n_s = (long long)sc.HW * A;
        sc.row_off = (long long)h->Ceff * boxes;
        boxes += sc.n_s;
        if (head[s] && (((uintptr_t)head[s]) & 3) != 0) VY_FAIL(VY_EALIGN, "head[%d] not 4-byte aligned", s);
    }
    h->R = (long long)h->Ceff * boxes;
    if (h->R > 0xfffffffeLL) VY_FAIL(VY_EINVAL, "R=%lld rows per image exceed the 32-bit row index", h->R);
    if ((long long)B * A * h->P > 0x7fffffffLL) VY_FAIL(VY_EINVAL, "B*A*P overflows int");
    return VY_OK;
}
