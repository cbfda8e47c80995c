// What can a kernel READ from HBM on this GPU?  Ceiling probes for the streaming pass (tools only, not product code):
//   A  ld.global.nc.v4 per lane, 8 loads in flight per thread, 4 CTAs x 256 threads per SM
//   B  cp.async.cg 16 B per lane into a per-warp shared-memory ring (the streaming kernel's mechanism), 12 in flight
//   C  1-D bulk copies (cp.async.bulk.shared::cluster.global, mbarrier complete_tx) of CHUNK bytes into a ring of
//      STAGES buffers per CTA, one elected thread issuing, everybody waiting on the barrier and touching one word
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/read_bw tools/read_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 4) probe_a(const float4 *__restrict__ x, size_t n4, float *out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(x + i + u * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x > 1e30f ? 1.f : 0.f;
    }
    if (acc != 0.f) out[0] = acc;
}

__global__ void __launch_bounds__(256, 4) probe_b(const float4 *__restrict__ x, size_t n4, float *out) {
    extern __shared__ float4 ring[];            // [8 warps][12][32]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4 *mine = ring + (size_t)wid * 12 * 32 + lane;
    const size_t nwarps = (size_t)gridDim.x * 8;
    const size_t per = n4 / 32 / nwarps;        // float4-rows of 32 lanes per warp
    const float4 *p = x + ((size_t)blockIdx.x * 8 + wid) * per * 32 + lane;
    float acc = 0.f;
    for (int u = 0; u < 12 && u < (int)per; ++u) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(mine + u * 32)), "l"(p + (size_t)u * 32) : "memory");
        if ((u & 3) == 3) asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int slot = 0;
    for (size_t r = 0; r + 4 <= per; r += 4) {
        asm volatile("cp.async.wait_group 2;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += mine[(slot + u) * 32].x > 1e30f ? 1.f : 0.f;
        if (r + 12 + 4 <= per) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(mine + (slot + u) * 32)), "l"(p + (r + 12 + u) * 32) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        slot = slot == 8 ? 0 : slot + 4;
    }
    if (acc != 0.f) out[0] = acc;
}

template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(128, 1) probe_c(const char *__restrict__ x, size_t bytes, float *out) {
    extern __shared__ __align__(128) char buf[];    // STAGES x CHUNK
    __shared__ uint64_t bar[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const size_t nchunks = bytes / CHUNK;
    const size_t per = nchunks / gridDim.x;
    const char *p = x + (size_t)blockIdx.x * per * CHUNK;
    auto issue = [&](size_t c, int s) {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((uint32_t)__cvta_generic_to_shared(buf + (size_t)s * CHUNK)), "l"(p + c * CHUNK), "r"(CHUNK), "r"(b) : "memory");
    };
    if (tid == 0) for (int s = 0; s < STAGES && s < (int)per; ++s) issue(s, s);
    float acc = 0.f;
    uint32_t phase = 0;
    int s = 0;
    for (size_t c = 0; c < per; ++c) {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[s]);
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(phase) : "memory");
        acc += ((const float *)(buf + (size_t)s * CHUNK))[tid] > 1e30f ? 1.f : 0.f;
        __syncthreads();                                   // everybody is done with the stage
        if (tid == 0 && c + STAGES < per) issue(c + STAGES, s);
        if (++s == STAGES) { s = 0; phase ^= 1u; }
    }
    if (acc != 0.f) out[0] = acc;
}

// D: the streaming kernel's geometry without its arithmetic: planes of HW floats, (b, a) blocks of P planes of which the
// last C are read; a warp owns ROWS x 128 consecutive positions of PU consecutive planes and keeps 12 float4 per lane in
// flight (12 / ROWS planes); consecutive warps own consecutive position chunks of the same plane group.
template <int ROWS>
__global__ void __launch_bounds__(256, 4) probe_d(const float *__restrict__ x, int HW, int P, int C, int PU, int n_ba, float *out) {
    extern __shared__ float4 ring[];            // [8 warps][12][32]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4 *mine = ring + (size_t)wid * 12 * 32 + lane;
    const int chunks = (HW + 128 * ROWS - 1) / (128 * ROWS), groups = (C + PU - 1) / PU;
    const long long n_units = (long long)n_ba * groups * chunks, n_warps = (long long)gridDim.x * 8;
    constexpr int PF = 12 / ROWS;               // planes in flight
    float acc = 0.f;
    for (long long unit = (long long)blockIdx.x * 8 + wid; unit < n_units; unit += n_warps) {
        const int chunk = (int)(unit % chunks);
        const long long q = unit / chunks;
        const int grp = (int)(q % groups);
        const long long ba = q / groups;
        const int c0 = grp * PU, np = min(PU, C - c0);
        int pos = chunk * 128 * ROWS + lane * 4;
        const float *base = x + ((size_t)ba * P + (P - C) + c0) * HW;
        auto issue = [&](int pl_i, int slot) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const int pp = min(pos + r * 128, HW - 4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(mine + (slot * ROWS + r) * 32)), "l"(base + (size_t)pl_i * HW + pp) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int k = 0; k < PF; ++k) { if (k < np) issue(k, k); else asm volatile("cp.async.commit_group;" ::: "memory"); }
        int slot = 0;
        for (int k = 0; k < np; ++k) {
            asm volatile("cp.async.wait_group %0;" :: "n"(PF - 1) : "memory");
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc += mine[(slot * ROWS + r) * 32].x > 1e30f ? 1.f : 0.f;
            if (k + PF < np) issue(k + PF, slot); else asm volatile("cp.async.commit_group;" ::: "memory");
            slot = slot + 1 == PF ? 0 : slot + 1;
        }
    }
    if (acc != 0.f) out[0] = acc;
}

// E: the contiguous alternative WITH the streaming pass's arithmetic: every warp streams its own contiguous region
// through a cp.async ring of DEPTH float4 per lane and tests each element against a per-position bound looked up in a
// shared-memory table of HW floats (position = element index mod HW), 4 compares per float4.
template <int DEPTH, int CTAS>
__global__ void __launch_bounds__(256, CTAS) probe_e(const float4 *__restrict__ x, size_t n4, int HW, float *out) {
    extern __shared__ float4 smem[];            // [8 warps][DEPTH][32] ring, then HW/4 float4 of bounds
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4 *mine = smem + (size_t)wid * DEPTH * 32 + lane;
    float4 *tab = smem + 8 * DEPTH * 32;
    const int HW4 = HW / 4;
    for (int i = threadIdx.x; i < HW4; i += 256) tab[i] = make_float4(1e30f, 1e30f, 1e30f, 1e30f);
    __syncthreads();
    const size_t nwarps = (size_t)gridDim.x * 8;
    const size_t per = n4 / 32 / nwarps;        // rows of 32 float4 per warp
    const size_t row0 = ((size_t)blockIdx.x * 8 + wid) * per;
    const float4 *p = x + row0 * 32 + lane;
    int tpos = (int)((row0 * 32 + lane) % HW4); // this lane's float4 index inside the plane
    unsigned hits = 0;
    constexpr int G = 4;                        // float4 rows per commit group
    for (int u = 0; u < DEPTH && u < (int)per; ++u) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(mine + u * 32)), "l"(p + (size_t)u * 32) : "memory");
        if ((u % G) == G - 1) asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int slot = 0;
    for (size_t r = 0; r + G <= per; r += G) {
        asm volatile("cp.async.wait_group %0;" :: "n"(DEPTH / G - 1) : "memory");
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const float4 v = mine[(slot + u) * 32];
            const float4 t = tab[tpos];
            hits |= (v.x >= t.x ? 1u : 0u) | (v.y >= t.y ? 2u : 0u) | (v.z >= t.z ? 4u : 0u) | (v.w >= t.w ? 8u : 0u);
            tpos += 32; if (tpos >= HW4) tpos -= HW4;
        }
        if (r + DEPTH + G <= per) {
#pragma unroll
            for (int u = 0; u < G; ++u)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(mine + (slot + u) * 32)), "l"(p + (r + DEPTH + u) * 32) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        slot = slot + G == DEPTH ? 0 : slot + G;
    }
    if (hits) out[0] = (float)hits;
}

template <class F> static float timeit(F f, int n = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    float best = 1e30f;
    for (int i = 0; i < n; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    const size_t bytes = (size_t)495 << 20;
    char *x; float *out;
    cudaMalloc(&x, bytes); cudaMalloc(&out, 4); cudaMemset(x, 0, bytes);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float ms = timeit([&] { probe_a<<<sms * 4, 256>>>((const float4 *)x, bytes / 16, out); });
    printf("A ld.global.nc.v4 x8              : %.1f us -> %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    cudaFuncSetAttribute(probe_b, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 12 * 32 * 16);
    ms = timeit([&] { probe_b<<<sms * 4, 256, 8 * 12 * 32 * 16>>>((const float4 *)x, bytes / 16, out); });
    printf("B cp.async.cg ring 3x4 per warp   : %.1f us -> %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    {
        constexpr int CH = 16384, ST = 8;
        cudaFuncSetAttribute(probe_c<CH, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST);
        ms = timeit([&] { probe_c<CH, ST><<<sms, 128, CH * ST>>>(x, bytes, out); });
        printf("C bulk copy 16 KB x 8 stages, 1/SM : %.1f us -> %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    }
    {
        constexpr int CH = 8192, ST = 24;
        cudaFuncSetAttribute(probe_c<CH, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST);
        ms = timeit([&] { probe_c<CH, ST><<<sms, 128, CH * ST>>>(x, bytes, out); });
        printf("C bulk copy 8 KB x 24 stages, 1/SM : %.1f us -> %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    }
    {
        constexpr int CH = 4096, ST = 12;
        cudaFuncSetAttribute(probe_c<CH, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST);
        ms = timeit([&] { probe_c<CH, ST><<<sms * 4, 128, CH * ST>>>(x, bytes, out); });
        printf("C bulk copy 4 KB x 12 stages, 4/SM : %.1f us -> %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    }
    {
        // COCO 608 scale 2: 76 x 76 planes, 85 planes per (b, a) of which 80 are read, 64 x 3 blocks = 377 MB (354 MB read)
        const int HW = 5776, P = 85, C = 80, n_ba = 192;
        const double rd = (double)n_ba * C * HW * 4;
        cudaFuncSetAttribute(probe_d<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 12 * 32 * 16);
        cudaFuncSetAttribute(probe_d<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 12 * 32 * 16);
        cudaFuncSetAttribute(probe_d<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 12 * 32 * 16);
        for (int PU : {20, 40, 80}) {
            ms = timeit([&] { probe_d<1><<<sms * 4, 256, 8 * 12 * 32 * 16>>>((const float *)x, HW, P, C, PU, n_ba, out); });
            printf("D planes 76x76, warp = 128 pos x %2d planes, 12 planes in flight : %.1f us -> %.0f GB/s\n", PU, ms * 1e3, rd / ms / 1e6);
            ms = timeit([&] { probe_d<2><<<sms * 4, 256, 8 * 12 * 32 * 16>>>((const float *)x, HW, P, C, PU, n_ba, out); });
            printf("D planes 76x76, warp = 256 pos x %2d planes,  6 planes in flight : %.1f us -> %.0f GB/s\n", PU, ms * 1e3, rd / ms / 1e6);
            ms = timeit([&] { probe_d<4><<<sms * 4, 256, 8 * 12 * 32 * 16>>>((const float *)x, HW, P, C, PU, n_ba, out); });
            printf("D planes 76x76, warp = 512 pos x %2d planes,  3 planes in flight : %.1f us -> %.0f GB/s\n", PU, ms * 1e3, rd / ms / 1e6);
        }
    }
    {
        const int HW = 5776;
        const size_t sm12 = (size_t)8 * 12 * 32 * 16 + HW * 4, sm8 = (size_t)8 * 8 * 32 * 16 + HW * 4;
        cudaFuncSetAttribute(probe_e<12, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm12);
        ms = timeit([&] { probe_e<12, 3><<<sms * 3, 256, sm12>>>((const float4 *)x, bytes / 16, HW, out); });
        printf("E contiguous + table lookup + compares, ring 12, 3 CTAs/SM : %.1f us -> %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        cudaFuncSetAttribute(probe_e<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm8);
        ms = timeit([&] { probe_e<8, 4><<<sms * 4, 256, sm8>>>((const float4 *)x, bytes / 16, HW, out); });
        printf("E contiguous + table lookup + compares, ring  8, 4 CTAs/SM : %.1f us -> %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
