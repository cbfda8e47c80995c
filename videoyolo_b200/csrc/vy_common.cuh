// vy_common.cuh -- shared device helpers: error plumbing, decode arithmetic, orderable keys.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/vyolo.h"

typedef unsigned long long u64;
typedef unsigned int u32;

// ------------------------------------------------------------------ host-side error plumbing
void vy_set_error(const char *fmt, ...);
#define VY_FAIL(code, ...) do { vy_set_error(__VA_ARGS__); return (code); } while (0)
#define VY_CUDA_CHECK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
    vy_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    return VY_ECUDA; } } while (0)
#define VY_LAUNCH_CHECK(name) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    vy_set_error("launch of %s failed: %s", name, cudaGetErrorString(e_)); return VY_ECUDA; } } while (0)

int vy_sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel): only raises the limit, remembers what it set
cudaError_t vy_ensure_dyn_smem(const void *func, size_t bytes);

// ------------------------------------------------------------------ launch accounting (vyolo.h: vy_prof_*)
// Every kernel launch of the library goes through VY_KERNEL: it is counted always and, while
// vy_prof_enable(1) is in force, bracketed by cudaEventRecord on the launching stream.
void vy_prof_pre(int kernel_id, cudaStream_t st);
void vy_prof_post(int kernel_id, cudaStream_t st);
#define VY_KERNEL(id, st, ...) do { vy_prof_pre((id), (st)); __VA_ARGS__; vy_prof_post((id), (st)); } while (0)

// ---- programmatic dependent launch (PDL): a kernel launched with vy_launch(..., pdl = true) may become
// resident while its predecessor in the stream is still running; it must call vy_grid_dep_wait() before it
// touches anything the predecessor wrote (a no-op without the attribute), and a predecessor lets its
// dependents in early with vy_grid_dep_trigger().  Hides the launch latency and ramp between the short
// kernels of one call.
__device__ __forceinline__ void vy_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void vy_grid_dep_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool vy_pdl_enabled();           // false with VY_NO_PDL=1 in the environment (A/B timing)
template <typename... KArgs, typename... Args>
static inline cudaError_t vy_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    bool pdl, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && vy_pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ------------------------------------------------------------------ head-map description
// One YOLO output scale (yolo3.py:43-74): NCHW head map + the constants the decode needs.
struct VyScale {
    const float *head;               // (B, A*P, H, W)
    int H, W, HW;
    int vec;                         // 4 when HW % 4 == 0 and head is 16-byte aligned, else 1
    float stride;
    float aw[VY_MAX_ANCHORS], ah[VY_MAX_ANCHORS];
    long long row_off;               // Ceff * sum_{s'<s} n_s'   (scale concat, yolo3.py:523)
    long long n_s;                   // H*W*A
};

struct VyHeads {
    VyScale sc[VY_MAX_SCALES];
    int n_scales, B, A, C, P;        // P = 5 + C (yolo3.py:48)
    int agnostic;                    // yolo3.py:184-188
    int Ceff;                        // agnostic ? 1 : C
    long long R;                     // rows per image of the reference detection tensor
};

int vy_fill_heads(VyHeads *h, const float *const *head, const int *H, const int *W,
                  const float *stride, const float *anchors, int n_scales, int B, int A, int C,
                  int agnostic);

// ------------------------------------------------------------------ decode arithmetic
// mshadow_op::sigmoid, 1/(1+exp(-x)) in fp32 (F.sigmoid at yolo3.py:172,174,175).  MUFU-based:
// ex2.approx(-x*log2e) and rcp.approx; relative error <= ~3e-7 + |x|*6e-8 (the rounding of
// x*log2e), i.e. < 2e-6 for |x| <= 24 -- inside the 1e-5 parity budget.  Every score the library
// ever produces goes through this one function, so the fused and the materialising paths order
// candidates identically.
__device__ __forceinline__ float vy_sigmoid(float x) {
    return __fdividef(1.0f, __fadd_rn(1.0f, __expf(-x)));
}

// score = sigmoid(class_pred) * confidence   (yolo3.py:175)
__device__ __forceinline__ float vy_score(float t_cls, float conf) {
    return __fmul_rn(vy_sigmoid(t_cls), conf);
}

// box corners from the 4 raw box logits (yolo3.py:172-173,176-177)
__device__ __forceinline__ float4 vy_box(float tx, float ty, float tw, float th,
                                         int x, int y, float stride, float aw, float ah) {
    const float cx = __fmul_rn(__fadd_rn(vy_sigmoid(tx), (float)x), stride);
    const float cy = __fmul_rn(__fadd_rn(vy_sigmoid(ty), (float)y), stride);
    const float hw = __fdiv_rn(__fmul_rn(expf(tw), aw), 2.0f);
    const float hh = __fdiv_rn(__fmul_rn(expf(th), ah), 2.0f);
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

// ------------------------------------------------------------------ orderable 64-bit keys
// key = (orderable(score) << 32) | ~row : larger key == earlier in MXNet's stable descending sort
// (score desc, ties -> lower source row first).  0 is never a valid key (it would need a NaN score).
__device__ __forceinline__ u32 vy_f2ord(float f) {
    u32 u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float vy_ord2f(u32 o) {
    u32 u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ u64 vy_make_key(float score, u32 row) {
    return ((u64)vy_f2ord(score) << 32) | (u64)(0xffffffffu - row);
}
__device__ __forceinline__ u32 vy_key_row(u64 key) { return 0xffffffffu - (u32)(key & 0xffffffffu); }
__device__ __forceinline__ float vy_key_score(u64 key) { return vy_ord2f((u32)(key >> 32)); }

__device__ __forceinline__ float4 vy_ldg128(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float vy_ldg32(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// L1-allocating variants: the selection kernel re-reads the rare planes that hold a hit from L1
__device__ __forceinline__ float4 vy_ldg128_ca(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float vy_ldg32_ca(const float *p) {
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
