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

// ------------------------------------------------------------------ launch accounting
#include <atomic>
#include <mutex>
#include <vector>
static const char *const g_kernel_names[VY_K_COUNT] = {
    "vy_decode_kernel", "vy_decode_select_kernel", "vy_rows_select_kernel", "vy_nms_finalize_kernel",
    "vy_fill_kernel", "vy_bbox_iou_kernel", "vy_fusion_conv_kernel", "vy_temporal_pool_kernel",
    "vy_nms_large_kernels", "vy_layout_kernels", "vy_decode_sample_kernel", "vy_decode_stream_kernel",
    "vy_decode_table_kernel"};
static std::atomic<long long> g_launches[VY_K_COUNT];
static std::atomic<int> g_prof_on{0};
struct ProfRec { int id; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static thread_local cudaEvent_t g_prof_open = nullptr;

bool vy_pdl_enabled() {
    static const bool on = getenv("VY_NO_PDL") == nullptr;
    return on;
}
void vy_prof_pre(int id, cudaStream_t st) {
    if (id < 0 || id >= VY_K_COUNT) return;
    g_launches[id].fetch_add(1, std::memory_order_relaxed);
    g_prof_open = nullptr;
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    cudaEvent_t a = nullptr;
    if (cudaEventCreate(&a) != cudaSuccess) return;
    cudaEventRecord(a, st);
    g_prof_open = a;
}
void vy_prof_post(int id, cudaStream_t st) {
    if (!g_prof_open) return;
    cudaEvent_t b = nullptr;
    if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(g_prof_open); g_prof_open = nullptr; return; }
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(ProfRec{id, g_prof_open, b});
    g_prof_open = nullptr;
}
extern "C" const char *vy_kernel_name(int id) { return (id >= 0 && id < VY_K_COUNT) ? g_kernel_names[id] : ""; }
extern "C" int vy_launch_counts(long long *counts, int n) {
    if (!counts || n < 0) VY_FAIL(VY_EINVAL, "vy_launch_counts: bad arguments");
    for (int i = 0; i < n; ++i) counts[i] = i < VY_K_COUNT ? g_launches[i].load() : 0;
    return VY_K_COUNT;
}
extern "C" int vy_prof_enable(int on) { g_prof_on.store(on ? 1 : 0); return VY_OK; }
extern "C" int vy_prof_read(double *ms, long long *launches, int n) {
    if (!ms || !launches || n < 0) VY_FAIL(VY_EINVAL, "vy_prof_read: bad arguments");
    for (int i = 0; i < n; ++i) { ms[i] = 0.0; launches[i] = 0; }
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int rc = VY_OK;
    for (const ProfRec &r : g_prof_recs) {
        float t = 0.0f;
        if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) rc = VY_ECUDA;
        else if (r.id < n) { ms[r.id] += (double)t; launches[r.id] += 1; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof_recs.clear();
    if (rc != VY_OK) vy_set_error("vy_prof_read: an event could not be read");
    return rc;
}

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

#include <map>
cudaError_t vy_ensure_dyn_smem(const void *func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> set;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    size_t &cur = set[std::make_pair(func, dev)];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
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
    for (int s = 0; s < n_scales; ++s)        // SelBuf.it_off and the streaming pass's element indices are 32-bit per scale
        if ((long long)B * A * h->P * h->sc[s].HW > 0xffffffffLL)
            VY_FAIL(VY_EUNSUPPORTED, "scale %d holds %lld elements (>= 2^32): split the batch", s,
                    (long long)B * A * h->P * h->sc[s].HW);
    return VY_OK;
}
