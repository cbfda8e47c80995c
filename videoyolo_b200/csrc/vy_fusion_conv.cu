// vy_fusion_conv.cu -- temporal fusion convolution: LeakyReLU(BN(ConvND(x))) as a tcgen05/TMEM implicit
// GEMM fed by TMA.  Replaces the Conv / _conv2d / _conv3d / _conv21d cells of
// models/definitions/layers.py:63-89,135-158 as used by YOLODetectionBlockV3 (yolo3.py:229-253).
//
// Activation layout ("P layout", the library's own; vy_pack_* / vy_unpack_* convert from and to the
// reference's NCDHW fp32):   [T][B][Hp = H+2][Wp = W+2][C]  bf16, the one-pixel spatial border is zero.
//   * time is the OUTERMOST axis, so all frames of the batch form one matrix X_t[rows = B*Hp*Wp][C];
//   * with the border materialised, the input of filter tap (dt, dh, dw) for 128 consecutive output
//     rows r0.. is simply the 128 consecutive rows r0 + dh*Wp + dw .. of X_{t+dt}: a plain TMA tile.
//     Frames t+dt outside [0, T) and rows outside the matrix are zero-filled by TMA itself
//     (the (kt/2) temporal 'same' padding of layers.py:76,86 costs nothing: those taps are skipped).
//   => the convolution is   Y_t[r0:r0+128, n0:n0+BN] = sum_{taps, Cin blocks} A_tap(128 x 64) * W_tap(BN x 64)^T
//      with every operand tile fetched by one cp.async.bulk.tensor, and no im2col buffer anywhere.
//   Outputs are produced for border rows too (garbage) and overwritten with zeros in the epilogue, so
//   the result is again a valid P-layout tensor and conv cells chain without repacking.
//
// Kernel: persistent, one CTA per SM, 192 threads = warp 0 TMA producer (one lane), warp 1 MMA issuer
// (one lane; owns TMEM), warps 2-5 epilogue.  4-stage smem ring (A 16 KB + B BN*128 B per stage,
// 128-byte swizzle), accumulators double-buffered in TMEM (2 x BN columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1.  Epilogue: tcgen05.ld -> scale/shift (folded BN) -> LeakyReLU ->
// border zeroing -> bf16 (or fp32) -> global.
#include "vy_common.cuh"
#include <stdlib.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

namespace {

constexpr int CV_BM = 128;            // output rows per tile (UMMA M)
constexpr int CV_BK = 64;             // channels per k-block: 64 bf16 = one 128-byte swizzle row
constexpr int CV_STAGES = 4;
constexpr int CV_NT = 192;
constexpr int CV_MAX_TAPS = 27;

struct ConvParams {
    int T, rows, Hp, Wp, Cin, Cout;
    int kt, kh, kw;
    int m_tiles, n_tiles, cin_blocks;
    long long n_tiles_total;
    float slope;
    const float *scale, *shift;
    void *y;
    int y_is_f32;
    int nchw_C, nchw_H, nchw_W;  // nchw_C > 0: y is the reference's fp32 (B, nchw_C, H, W) tensor (T == 1): interior pixels, first nchw_C channels
    int pool_max;                // y is ONE P-layout frame [rows][Cout] bf16 pre-filled with -inf: max over the T frames (TemporalPooling 'max')
    // a_cpb > 0: 1 x 1 conv over the 'cat'-joined window WITHOUT materialising the join (yolo3.py:1135-1136): k-block cb
    // reads channels (cb % a_cpb) * 64 of frame (cb / a_cpb) % a_frames of x; the k range beyond a_frames * a_cpb blocks
    // wraps around (the activation repeated for split weights)
    int a_cpb, a_frames;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, u64 *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, u64 *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(u64 *bar) {      // arrives on bar when all MMAs issued so far are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, one 128 x N x 16 step
__device__ __forceinline__ void tc_mma(u32 d_tmem, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane = output row)
__device__ __forceinline__ void tc_ld32(u32 taddr, u32 (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor of a K-major tile stored as rows of 128 bytes with the 128-byte
// swizzle (what TMA SWIZZLE_128B writes): 8-row groups 1024 B apart (SBO), descriptor version 1,
// layout type 2 (SWIZZLE_128B).  Stepping K by 16 bf16 (32 B) inside the row adds 2 to the address field.
__device__ __forceinline__ u64 umma_desc(u32 smem_addr) {
    return (u64)((smem_addr >> 4) & 0x3fffu) | ((u64)1 << 16) | ((u64)(1024 >> 4) << 32) | ((u64)1 << 46) | ((u64)2 << 61);
}
// instruction descriptor: D fp32 (bit 4), A/B bf16 (bits 7, 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr u32 umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(N >> 3) << 17) | ((u32)(M >> 4) << 24);
}

// taps (dt, dh, dw) of a tile at frame t, temporal taps that leave [0, T) dropped
struct TapIter {
    int kt, kh, kw, T, t;
    __device__ __forceinline__ int count() const {
        int n = 0;
        for (int a = 0; a < kt; ++a) { const int tt = t + a - kt / 2; n += (tt >= 0 && tt < T); }
        return n * kh * kw;
    }
};

// epilogue of one accumulator tile: 32 rows of the tile per warp (thread = TMEM lane = output row), BN columns in
// chunks of 32: folded BN scale/shift, LeakyReLU, border rows zeroed, bf16 / fp32 P layout or fp32 NCHW store
template <int BN>
__device__ __forceinline__ void conv_epilogue_tile(const ConvParams &cp, u32 tmem_acc, int quarter, int r, int n0, int t) {
    const int plane = cp.Hp * cp.Wp;
    const bool in_range = r < cp.rows;
    const int rr = r % plane, hp = rr / cp.Wp, wp = rr % cp.Wp;
    const bool interior = hp > 0 && hp < cp.Hp - 1 && wp > 0 && wp < cp.Wp - 1;
    const size_t out_off = ((size_t)(cp.pool_max ? 0 : t) * cp.rows + (size_t)r) * cp.Cout + n0;
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
        u32 v[32];
        tc_ld32(tmem_acc + ((u32)(quarter * 32) << 16) + (u32)(ch * 32), v);
        if (in_range) {
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int n = n0 + ch * 32 + i;
                float x = fmaf(__uint_as_float(v[i]), __ldg(cp.scale + n), __ldg(cp.shift + n));   // layers.py:68,77
                x = x > 0.0f ? x : x * cp.slope;                                                   // layers.py:69,78
                o[i] = interior ? x : 0.0f;
            }
            if (cp.nchw_C > 0) {
                // the head maps of YOLOOutputV3 leave the GEMM in the reference's own layout (yolo3.py:157-158):
                // consecutive lanes hold consecutive pixels, so every channel is one 128-byte run per warp
                if (interior) {
                    const int bimg = r / plane;
                    float *dst = (float *)cp.y + ((size_t)bimg * cp.nchw_C * cp.nchw_H + (size_t)(hp - 1)) * cp.nchw_W + (wp - 1);
                    const size_t cstride = (size_t)cp.nchw_H * cp.nchw_W;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int n = n0 + ch * 32 + i;
                        if (n < cp.nchw_C) dst[(size_t)n * cstride] = o[i];
                    }
                }
            } else if (cp.pool_max) {
                // the late 'max' join over the window (TemporalPooling, layers.py:201-205; yolo3.py:1134-1138) in the
                // epilogue: the K frames' tiles meet in the one pooled frame through 16-byte bf16x2 max reductions at
                // L2 (order-independent, so still deterministic); the un-pooled tip is never written
                unsigned *dst = (unsigned *)((__nv_bfloat16 *)cp.y + out_off + ch * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(o[8 * i], o[8 * i + 1]);
                    __nv_bfloat162 p1 = __floats2bfloat162_rn(o[8 * i + 2], o[8 * i + 3]);
                    __nv_bfloat162 p2 = __floats2bfloat162_rn(o[8 * i + 4], o[8 * i + 5]);
                    __nv_bfloat162 p3 = __floats2bfloat162_rn(o[8 * i + 6], o[8 * i + 7]);
                    asm volatile("red.global.max.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};"
                                 :: "l"(dst + 4 * i), "r"(*(u32 *)&p0), "r"(*(u32 *)&p1), "r"(*(u32 *)&p2), "r"(*(u32 *)&p3) : "memory");
                }
            } else if (cp.y_is_f32) {
                float4 *dst = (float4 *)((float *)cp.y + out_off + ch * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            } else {
                uint4 *dst = (uint4 *)((__nv_bfloat16 *)cp.y + out_off + ch * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(o[8 * i], o[8 * i + 1]);
                    __nv_bfloat162 p1 = __floats2bfloat162_rn(o[8 * i + 2], o[8 * i + 3]);
                    __nv_bfloat162 p2 = __floats2bfloat162_rn(o[8 * i + 4], o[8 * i + 5]);
                    __nv_bfloat162 p3 = __floats2bfloat162_rn(o[8 * i + 6], o[8 * i + 7]);
                    dst[i] = make_uint4(*(u32 *)&p0, *(u32 *)&p1, *(u32 *)&p2, *(u32 *)&p3);
                }
            }
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(CV_NT, 1)
vy_fusion_conv_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ ConvParams cp) {
    constexpr u32 A_BYTES = CV_BM * CV_BK * 2, B_BYTES = BN * CV_BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr u32 TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) u64 bar_full[CV_STAGES], bar_empty[CV_STAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ u32 tmem_base_sh;
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < CV_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base_sh)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tmem_base = tmem_base_sh;
    // programmatic dependent launch: everything above ran beside the tail of the kernel before this one in the stream;
    // nothing below touches global memory before that kernel has completed, and the next conv of the chain may set up now
    vy_grid_dep_wait();
    vy_grid_dep_trigger();

    const int khw = cp.kh * cp.kw;
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; u32 phase = 0;
            for (long long tile = blockIdx.x; tile < cp.n_tiles_total; tile += gridDim.x) {
                const int n_t = (int)(tile % cp.n_tiles);
                const long long rest = tile / cp.n_tiles;
                const int m_t = (int)(rest % cp.m_tiles), t = (int)(rest / cp.m_tiles);
                const int r0 = m_t * CV_BM, n0 = n_t * BN;
                for (int a = 0; a < cp.kt; ++a) {
                    const int tt = t + a - cp.kt / 2;
                    if (tt < 0 || tt >= cp.T) continue;
                    for (int hw = 0; hw < khw; ++hw) {
                        const int dh = hw / cp.kw - cp.kh / 2, dw = hw % cp.kw - cp.kw / 2;
                        const int row = r0 + dh * cp.Wp + dw;
                        const int kbase = (a * khw + hw) * cp.Cin;
                        for (int cb = 0; cb < cp.cin_blocks; ++cb) {
                            mbar_wait(&bar_empty[stage], phase ^ 1u);
                            unsigned char *sa = tiles + (size_t)stage * STAGE_BYTES, *sb = sa + A_BYTES;
                            mbar_expect_tx(&bar_full[stage], STAGE_BYTES);
                            tma_load_3d(sa, &map_x, &bar_full[stage], cp.a_cpb ? (cb % cp.a_cpb) * CV_BK : cb * CV_BK, row,
                                        cp.a_cpb ? (cb / cp.a_cpb) % cp.a_frames : tt);
                            tma_load_2d(sb, &map_w, &bar_full[stage], kbase + cb * CV_BK, n0);
                            if (++stage == CV_STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr u32 idesc = umma_idesc(CV_BM, BN);
            int stage = 0; u32 phase = 0;
            int acc = 0; u32 acc_phase = 0;
            for (long long tile = blockIdx.x; tile < cp.n_tiles_total; tile += gridDim.x) {
                const int t = (int)((tile / cp.n_tiles) / cp.m_tiles);
                TapIter ti{cp.kt, cp.kh, cp.kw, cp.T, t};
                const int n_kb = ti.count() * cp.cin_blocks;
                mbar_wait(&bar_acc_empty[acc], acc_phase ^ 1u);       // epilogue has drained this accumulator
                tc_fence_after();
                const u32 d_tmem = tmem_base + (u32)(acc * BN);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&bar_full[stage], phase);
                    tc_fence_after();
                    const u32 sa = smem_u32(tiles + (size_t)stage * STAGE_BYTES), sb = sa + A_BYTES;
                    const u64 da = umma_desc(sa), db = umma_desc(sb);
#pragma unroll
                    for (int k = 0; k < CV_BK / 16; ++k)
                        tc_mma(d_tmem, da + (u64)(2 * k), db + (u64)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&bar_empty[stage]);                     // smem slot free once these MMAs retire
                    if (++stage == CV_STAGES) { stage = 0; phase ^= 1u; }
                }
                tc_commit(&bar_acc_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                                  // TMEM lanes 32*quarter .. +31
        int acc = 0; u32 acc_phase = 0;
        for (long long tile = blockIdx.x; tile < cp.n_tiles_total; tile += gridDim.x) {
            const int n_t = (int)(tile % cp.n_tiles);
            const long long rest = tile / cp.n_tiles;
            const int m_t = (int)(rest % cp.m_tiles), t = (int)(rest / cp.m_tiles);
            const int r = m_t * CV_BM + quarter * 32 + lane, n0 = n_t * BN;
            mbar_wait(&bar_acc_full[acc], acc_phase);
            tc_fence_after();
            conv_epilogue_tile<BN>(cp, tmem_base + (u32)(acc * BN), quarter, r, n0, t);
            tc_fence_before();
            mbar_arrive(&bar_acc_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ CTA-pair variant (tcgen05 cta_group::2)
// Two CTAs of one cluster (the two SMs of a TPC) compute ONE 256 x BN tile: CTA r fetches its own 128 rows of A and
// HALF of the weight tile (rows n0 + r*BN/2 ..), and a single tcgen05.mma.cta_group::2 issued by the leader reads A
// from both CTAs and the two halves of B from both shared memories: per k-block a CTA receives 16 KB + BN*64 B
// instead of 16 KB + BN*128 B -- the operand traffic from L2, which bounds the one-CTA kernel in full waves, drops by
// a third at BN = 256, and the freed shared memory holds 6 stages instead of 4.  Accumulators: CTA r's TMEM lanes
// hold rows 128 r .. of the tile, double-buffered (2 x BN columns), each CTA runs its own epilogue.
//   barriers   full[s]      leader's; the leader's producer expects the bytes of BOTH CTAs, both CTAs' TMA loads
//                           complete on it (cp.async.bulk.tensor ... cta_group::2, barrier address with the peer bit cleared)
//              empty[s]     one per CTA; tcgen05.commit.cta_group::2 multicast from the leader frees the slot in both
//              acc_full[a]  one per CTA, same multicast commit after the tile's last k-block
//              acc_empty[a] leader's; 2 x 128 epilogue threads arrive (the peer's through mapa)
__device__ __forceinline__ u32 cluster_ctarank() { u32 r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr u32 PEER_BIT_MASK = 0xFEFFFFFFu;     // shared::cluster address of the same offset in the pair's even CTA
__device__ __forceinline__ void tma2_load_3d(void *dst, const CUtensorMap *map, u64 *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void *dst, const CUtensorMap *map, u64 *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc2_commit(u64 *bar) {     // arrives on `bar` of BOTH CTAs when all MMAs issued so far are done
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void tc2_mma(u32 d_tmem, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(u64 *bar, u32 cta) {      // arrive on the barrier at this offset in CTA `cta`
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        :: "r"(smem_u32(bar)), "r"(cta) : "memory");
}

constexpr int CV2_BM = 2 * CV_BM;
template <int BN> struct Conv2Cfg {
    static constexpr u32 A_BYTES = CV_BM * CV_BK * 2, B_BYTES = (BN / 2) * CV_BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES > 8 ? 8 : (192 * 1024) / STAGE_BYTES;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CV_NT, 1)
vy_fusion_conv2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                       const __grid_constant__ ConvParams cp) {
    using Cfg = Conv2Cfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr u32 A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr u32 TMEM_COLS = 2 * BN;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) u64 bar_full[STAGES], bar_empty[STAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ u32 tmem_base_sh;
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 rank = cluster_ctarank();
    const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {          // the same warp of both CTAs, same shared-memory offset for the result
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base_sh)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();       // barriers of the peer initialised, TMEM of both CTAs allocated
    tc_fence_after();
    const u32 tmem_base = tmem_base_sh;
    vy_grid_dep_wait();       // (programmatic dependent launch, as in the one-CTA kernel)
    vy_grid_dep_trigger();

    const int khw = cp.kh * cp.kw;
    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0; u32 phase = 0;
            for (long long tile = pair; tile < cp.n_tiles_total; tile += n_pairs) {
                const int n_t = (int)(tile % cp.n_tiles);
                const long long rest = tile / cp.n_tiles;
                const int m_t = (int)(rest % cp.m_tiles), t = (int)(rest / cp.m_tiles);
                const int r0 = m_t * CV2_BM + (int)rank * CV_BM, n0 = n_t * BN + (int)rank * (BN / 2);
                for (int a = 0; a < cp.kt; ++a) {
                    const int tt = t + a - cp.kt / 2;
                    if (tt < 0 || tt >= cp.T) continue;
                    for (int hw = 0; hw < khw; ++hw) {
                        const int dh = hw / cp.kw - cp.kh / 2, dw = hw % cp.kw - cp.kw / 2;
                        const int row = r0 + dh * cp.Wp + dw;
                        const int kbase = (a * khw + hw) * cp.Cin;
                        for (int cb = 0; cb < cp.cin_blocks; ++cb) {
                            mbar_wait(&bar_empty[stage], phase ^ 1u);
                            unsigned char *sa = tiles + (size_t)stage * STAGE_BYTES, *sb = sa + A_BYTES;
                            if (rank == 0) mbar_expect_tx(&bar_full[stage], 2 * STAGE_BYTES);
                            tma2_load_3d(sa, &map_x, &bar_full[stage], cp.a_cpb ? (cb % cp.a_cpb) * CV_BK : cb * CV_BK, row,
                                         cp.a_cpb ? (cb / cp.a_cpb) % cp.a_frames : tt);
                            tma2_load_2d(sb, &map_w, &bar_full[stage], kbase + cb * CV_BK, n0);
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0 && lane == 0) {
            constexpr u32 idesc = umma_idesc(CV2_BM, BN);
            int stage = 0; u32 phase = 0;
            int acc = 0; u32 acc_phase = 0;
            for (long long tile = pair; tile < cp.n_tiles_total; tile += n_pairs) {
                const int t = (int)((tile / cp.n_tiles) / cp.m_tiles);
                TapIter ti{cp.kt, cp.kh, cp.kw, cp.T, t};
                const int n_kb = ti.count() * cp.cin_blocks;
                mbar_wait(&bar_acc_empty[acc], acc_phase ^ 1u);       // both epilogues have drained this accumulator
                tc_fence_after();
                const u32 d_tmem = tmem_base + (u32)(acc * BN);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&bar_full[stage], phase);
                    tc_fence_after();
                    const u32 sa = smem_u32(tiles + (size_t)stage * STAGE_BYTES), sb = sa + A_BYTES;
                    const u64 da = umma_desc(sa), db = umma_desc(sb);
#pragma unroll
                    for (int k = 0; k < CV_BK / 16; ++k)
                        tc2_mma(d_tmem, da + (u64)(2 * k), db + (u64)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc2_commit(&bar_empty[stage]);                    // the slot is free in both CTAs once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                tc2_commit(&bar_acc_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5 of both CTAs) =====================
        const int quarter = warp & 3;
        int acc = 0; u32 acc_phase = 0;
        for (long long tile = pair; tile < cp.n_tiles_total; tile += n_pairs) {
            const int n_t = (int)(tile % cp.n_tiles);
            const long long rest = tile / cp.n_tiles;
            const int m_t = (int)(rest % cp.m_tiles), t = (int)(rest / cp.m_tiles);
            const int r = m_t * CV2_BM + (int)rank * CV_BM + quarter * 32 + lane, n0 = n_t * BN;
            mbar_wait(&bar_acc_full[acc], acc_phase);
            tc_fence_after();
            conv_epilogue_tile<BN>(cp, tmem_base + (u32)(acc * BN), quarter, r, n0, t);
            tc_fence_before();
            mbar_arrive_cta(&bar_acc_empty[acc], 0u);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tc_fence_before();
    cluster_sync_all();       // nobody leaves (shared memory, barriers, TMEM) while the pair still works
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ layout kernels
// (B, C, T, H, W)-strided fp32 (the reference's NCDHW after swapaxes, or (B, K, C, H, W) before it: the
// strides say which) <-> P layout.  One CTA moves a tile of LY_C channels x LY_P consecutive positions of one
// (t, b) frame through shared memory: the fp32 side is touched in runs of LY_P consecutive floats of one
// channel plane, the P side in runs of LY_C consecutive channels of one pixel.
constexpr int LY_C = 64, LY_P = 128, LY_NT = 256;

__global__ void __launch_bounds__(LY_NT)
vy_pack_kernel(const float *__restrict__ x, long long sb, long long sc, long long st, int B, int C,
               int T, int H, int W, __nv_bfloat16 *__restrict__ y, int Ctot, int split, int vec, int border) {
    // split > 0 (vy_pack_f32_split_to_p_bf16): every value v goes out as hi = bf16(v) at channel c and split + c and as
    // lo = bf16(v - hi) at 2*split + c of a pixel of Ctot = 3*split channels
    __shared__ float tile[LY_C][LY_P + 1];
    const int HW = H * W, Wp = W + 2;
    const int p0 = blockIdx.x * LY_P, c0 = blockIdx.y * LY_C;
    const int b = blockIdx.z % B, t = blockIdx.z / B;
    const float *src = x + (size_t)b * sb + (size_t)t * st;
    if (vec) {
        // even grids (H*W, the strides and the base are multiples of 4 floats): the whole tile in one round of eight
        // 16-byte loads per thread.  A warp takes 4 channels x 32 positions at a time: four full 128-byte runs on the
        // global side, and on the shared side (row stride 129 words) lane l writes bank (c + 4 (l % 8) + u + l / 8) % 32
        // -- no conflicts.
        const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float4 v[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int task = it * 8 + wq, c = 4 * (task >> 2) + (lane >> 3), p = 32 * (task & 3) + 4 * (lane & 7);
            v[it] = (c0 + c < C && p0 + p < HW) ? __ldg((const float4 *)(src + (size_t)(c0 + c) * sc + p0 + p))
                                                 : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int task = it * 8 + wq, c = 4 * (task >> 2) + (lane >> 3), p = 32 * (task & 3) + 4 * (lane & 7);
            tile[c][p] = v[it].x; tile[c][p + 1] = v[it].y; tile[c][p + 2] = v[it].z; tile[c][p + 3] = v[it].w;
        }
    } else
    // 8 independent 4-byte loads per thread and round: the fp32 side has no alignment to vectorise on
    // (H*W is odd at 13x13), so memory-level parallelism comes from unrolling
    for (int i0 = threadIdx.x; i0 < LY_C * LY_P; i0 += 8 * LY_NT) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * LY_NT, c = i / LY_P, p = i % LY_P;
            v[u] = (c0 + c < C && p0 + p < HW) ? __ldg(src + (size_t)(c0 + c) * sc + p0 + p) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * LY_NT;
            tile[i / LY_P][i % LY_P] = v[u];
        }
    }
    __syncthreads();
    __nv_bfloat16 *dst = y + ((size_t)t * B + b) * (size_t)(H + 2) * Wp * Ctot;
    if (split > 0) {
        for (int i = threadIdx.x; i < LY_P * LY_C; i += LY_NT) {
            const int p = i / LY_C, c = i % LY_C;
            const int pos = p0 + p;
            if (pos >= HW || c0 + c >= C) continue;
            const int h = pos / W, w = pos % W;
            const float v = tile[c][p];
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
            __nv_bfloat16 *o = dst + ((size_t)(h + 1) * Wp + (w + 1)) * Ctot + c0 + c;
            o[0] = hi; o[split] = hi; o[2 * split] = lo;
        }
        return;
    }
    if ((C & 7) == 0) {                                    // 8 channels = one 16-byte store
        for (int i = threadIdx.x; i < LY_P * (LY_C / 8); i += LY_NT) {
            const int p = i / (LY_C / 8), c = (i % (LY_C / 8)) * 8;
            const int pos = p0 + p;
            if (pos >= HW || c0 + c >= C) continue;
            const int h = pos / W, w = pos % W;
            uint4 o;
            __nv_bfloat162 *oh = (__nv_bfloat162 *)&o;
#pragma unroll
            for (int k = 0; k < 4; ++k) oh[k] = __floats2bfloat162_rn(tile[c + 2 * k][p], tile[c + 2 * k + 1][p]);
            *(uint4 *)(dst + ((size_t)(h + 1) * Wp + (w + 1)) * C + c0 + c) = o;
        }
        if (border && blockIdx.x == 0) {
            // the frame's one-pixel zero border, this CTA's channel block of it (no separate launch)
            const int Hp = H + 2, nb = 2 * Wp + 2 * H;
            for (int i = threadIdx.x; i < nb * (LY_C / 8); i += LY_NT) {
                const int q = i / (LY_C / 8), c = (i % (LY_C / 8)) * 8;
                if (c0 + c >= C) continue;
                int hp, wp;
                if (q < Wp) { hp = 0; wp = q; }
                else if (q < 2 * Wp) { hp = Hp - 1; wp = q - Wp; }
                else { const int r = q - 2 * Wp; hp = 1 + (r >> 1); wp = (r & 1) ? Wp - 1 : 0; }
                *(uint4 *)(dst + ((size_t)hp * Wp + wp) * C + c0 + c) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        return;
    }
    for (int i = threadIdx.x; i < LY_P * LY_C; i += LY_NT) {
        const int p = i / LY_C, c = i % LY_C;
        const int pos = p0 + p;
        if (pos >= HW || c0 + c >= C) continue;
        const int h = pos / W, w = pos % W;
        dst[((size_t)(h + 1) * Wp + (w + 1)) * C + c0 + c] = __float2bfloat16_rn(tile[c][p]);
    }
}

// the one-pixel zero border of every (t, b) frame of a P-layout tensor (2*Wp + 2*H pixels x C channels)
// Small planes whose channels lie back to back in memory (sc == H*W: the reference's contiguous (.., C, H, W) tensors; 13 x 13
// = 169 floats, no alignment per plane to vectorise on): 64 channels x H*W positions are ONE contiguous run of 64*H*W floats,
// read with 16-byte loads and scattered into the shared tile by flat index.  Writes the P-layout pixels and the zero border.
constexpr int LYF_MAX_HW = 176;
__global__ void __launch_bounds__(LY_NT)
vy_pack_flat_kernel(const float *__restrict__ x, long long sb, long long st, int B, int C, int H, int W, __nv_bfloat16 *__restrict__ y) {
    __shared__ float tile[LY_C * (LYF_MAX_HW + 1)];
    const int HW = H * W, Wp = W + 2, Hp = H + 2;
    const int S = HW | 1;                                  // odd row stride: the transposed read below is 2-way conflicted at worst
    const int c0 = blockIdx.x * LY_C;
    const int b = blockIdx.y % B, t = blockIdx.y / B;
    const float4 *src = (const float4 *)(x + (size_t)b * sb + (size_t)t * st + (size_t)c0 * HW);
    const int n4 = LY_C * HW / 4;
    for (int i0 = threadIdx.x; i0 < n4; i0 += 8 * LY_NT) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * LY_NT;
            v[u] = i < n4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * LY_NT;
            if (i < n4) {
                int c = (4 * i) / HW, p = 4 * i - c * HW;
                const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tile[c * S + p] = e[k];
                    if (++p == HW) { p = 0; ++c; }
                }
            }
        }
    }
    __syncthreads();
    __nv_bfloat16 *dst = y + ((size_t)t * B + b) * (size_t)Hp * Wp * C;
    for (int i = threadIdx.x; i < HW * (LY_C / 8); i += LY_NT) {
        const int p = i / (LY_C / 8), c = (i % (LY_C / 8)) * 8;
        const int h = p / W, w = p % W;
        uint4 o;
        __nv_bfloat162 *oh = (__nv_bfloat162 *)&o;
#pragma unroll
        for (int k = 0; k < 4; ++k) oh[k] = __floats2bfloat162_rn(tile[(c + 2 * k) * S + p], tile[(c + 2 * k + 1) * S + p]);
        *(uint4 *)(dst + ((size_t)(h + 1) * Wp + (w + 1)) * C + c0 + c) = o;
    }
    const int nb = 2 * Wp + 2 * H;
    for (int i = threadIdx.x; i < nb * (LY_C / 8); i += LY_NT) {
        const int q = i / (LY_C / 8), c = (i % (LY_C / 8)) * 8;
        int hp, wp;
        if (q < Wp) { hp = 0; wp = q; }
        else if (q < 2 * Wp) { hp = Hp - 1; wp = q - Wp; }
        else { const int r = q - 2 * Wp; hp = 1 + (r >> 1); wp = (r & 1) ? Wp - 1 : 0; }
        *(uint4 *)(dst + ((size_t)hp * Wp + wp) * C + c0 + c) = make_uint4(0u, 0u, 0u, 0u);
    }
}

__global__ void vy_zero_border_kernel(__nv_bfloat16 *__restrict__ y, int H, int W, int C) {
    const int Hp = H + 2, Wp = W + 2;
    const int nb = 2 * Wp + 2 * H;
    __nv_bfloat16 *frame = y + (size_t)blockIdx.y * Hp * Wp * C;
    for (int q = blockIdx.x; q < nb; q += gridDim.x) {
        int hp, wp;
        if (q < Wp) { hp = 0; wp = q; }
        else if (q < 2 * Wp) { hp = Hp - 1; wp = q - Wp; }
        else { const int r = q - 2 * Wp; hp = 1 + (r >> 1); wp = (r & 1) ? Wp - 1 : 0; }
        __nv_bfloat16 *o = frame + ((size_t)hp * Wp + wp) * C;
        for (int c = threadIdx.x; c < C; c += blockDim.x) o[c] = __float2bfloat16_rn(0.0f);
    }
}

// P layout (bf16 or fp32, Cp channels per pixel) -> (B, C, T, H, W)-strided fp32 of its first C <= Cp channels
template <typename TIn>
__global__ void __launch_bounds__(LY_NT)
vy_unpack_kernel(const TIn *__restrict__ y, int B, int Cp, int C, int T, int H, int W, float *__restrict__ x,
                 long long sb, long long sc, long long st) {
    __shared__ float tile[LY_C][LY_P + 1];
    const int HW = H * W, Wp = W + 2;
    const int p0 = blockIdx.x * LY_P, c0 = blockIdx.y * LY_C;
    const int b = blockIdx.z % B, t = blockIdx.z / B;
    const TIn *src = y + ((size_t)t * B + b) * (size_t)(H + 2) * Wp * Cp;
    constexpr int V = 16 / (int)sizeof(TIn);               // channels per 16-byte load
    if ((Cp % V) == 0) {                                   // vector loads, several in flight per thread
        for (int i = threadIdx.x; i < LY_P * (LY_C / V); i += LY_NT) {
            const int p = i / (LY_C / V), c = (i % (LY_C / V)) * V;
            const int pos = p0 + p;
            float f[V];
#pragma unroll
            for (int k = 0; k < V; ++k) f[k] = 0.0f;
            if (pos < HW && c0 + c < Cp) {                 // a vector never straddles the end of the pixel (Cp % V == 0)
                const int h = pos / W, w = pos % W;
                const uint4 q = *(const uint4 *)(src + ((size_t)(h + 1) * Wp + (w + 1)) * Cp + c0 + c);
                if (sizeof(TIn) == 2) {
                    const __nv_bfloat162 *qh = (const __nv_bfloat162 *)&q;
#pragma unroll
                    for (int k = 0; k < V / 2; ++k) { const float2 v = __bfloat1622float2(qh[k]); f[2 * k] = v.x; f[2 * k + 1] = v.y; }
                } else {
                    const float *qf = (const float *)&q;
#pragma unroll
                    for (int k = 0; k < V; ++k) f[k] = qf[k];
                }
            }
#pragma unroll
            for (int k = 0; k < V; ++k) tile[c + k][p] = f[k];
        }
    } else {
        for (int i = threadIdx.x; i < LY_P * LY_C; i += LY_NT) {
            const int p = i / LY_C, c = i % LY_C;
            const int pos = p0 + p;
            float v = 0.0f;
            if (pos < HW && c0 + c < C) {
                const int h = pos / W, w = pos % W;
                v = (float)src[((size_t)(h + 1) * Wp + (w + 1)) * Cp + c0 + c];
            }
            tile[c][p] = v;
        }
    }
    __syncthreads();
    float *dst = x + (size_t)b * sb + (size_t)t * st;
    for (int i = threadIdx.x; i < LY_C * LY_P; i += LY_NT) {
        const int c = i / LY_P, p = i % LY_P;
        if (c0 + c < C && p0 + p < HW) dst[(size_t)(c0 + c) * sc + p0 + p] = tile[c][p];
    }
}

// TemporalPooling 'direct' (layers.py:201-205) on P layout: y[i] = max / mean over t of x[t][i]
__device__ __forceinline__ void pool_accumulate(const uint4 &q, bool first, int mode, float (&acc)[8]) {
    const __nv_bfloat162 *h = (const __nv_bfloat162 *)&q;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        if (first) { acc[2 * k] = f.x; acc[2 * k + 1] = f.y; }
        else if (mode == 0) { acc[2 * k] = fmaxf(acc[2 * k], f.x); acc[2 * k + 1] = fmaxf(acc[2 * k + 1], f.y); }
        else { acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
    }
}
__global__ void vy_temporal_pool_kernel(const uint4 *__restrict__ x, int T, long long inner8, int mode, uint4 *__restrict__ y) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < inner8; i += stride) {
        float acc[8];
        // the first four frames are fetched together (one 16-byte load in flight per thread reaches 2/3 of the HBM rate)
        uint4 q0, q1, q2, q3;
        q0 = __ldg(x + i);
        if (T > 1) q1 = __ldg(x + (size_t)inner8 + i);
        if (T > 2) q2 = __ldg(x + (size_t)2 * inner8 + i);
        if (T > 3) q3 = __ldg(x + (size_t)3 * inner8 + i);
        pool_accumulate(q0, true, mode, acc);
        if (T > 1) pool_accumulate(q1, false, mode, acc);
        if (T > 2) pool_accumulate(q2, false, mode, acc);
        if (T > 3) pool_accumulate(q3, false, mode, acc);
        for (int t = 4; t < T; ++t) pool_accumulate(__ldg(x + (size_t)t * inner8 + i), false, mode, acc);
        uint4 o;
        __nv_bfloat162 *oh = (__nv_bfloat162 *)&o;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = mode == 0 ? acc[2 * k] : acc[2 * k] / (float)T;
            const float b = mode == 0 ? acc[2 * k + 1] : acc[2 * k + 1] / (float)T;
            oh[k] = __floats2bfloat162_rn(a, b);
        }
        y[i] = o;
    }
}


// Depthwise temporal merge (_conv1d, layers.py:50-60, as used by HDarknet h_darknet.py:97-119): a window of
// exactly T = kernel frames, Conv3D(kernel (T,1,1), groups = C, padding 0) + BN + LeakyReLU -> one frame:
//   y[i] = LReLU(scale[c] * sum_t w[c][t] * x[t][i] + shift[c]),  c = i % C,  border pixels stay zero.
// Bandwidth-bound: (T + 1) * inner * 2 bytes; 8 channels (one 16-byte vector) per thread and frame.
__global__ void vy_temporal_dwconv_kernel(const uint4 *__restrict__ x, int T, long long inner8, int C8, int Hp, int Wp,
                                          const float *__restrict__ w, const float *__restrict__ scale,
                                          const float *__restrict__ shift, float slope, uint4 *__restrict__ y) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < inner8; i += stride) {
        const int c0 = (int)(i % C8) * 8;
        const long long pix = i / C8;
        const int wp = (int)(pix % Wp), hp = (int)((pix / Wp) % Hp);
        const bool interior = hp > 0 && hp < Hp - 1 && wp > 0 && wp < Wp - 1;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
        for (int t = 0; t < T; ++t) {
            const uint4 q = x[(size_t)t * inner8 + i];
            const __nv_bfloat162 *h = (const __nv_bfloat162 *)&q;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(h[k]);
                acc[2 * k] = fmaf(f.x, __ldg(w + (size_t)(c0 + 2 * k) * T + t), acc[2 * k]);
                acc[2 * k + 1] = fmaf(f.y, __ldg(w + (size_t)(c0 + 2 * k + 1) * T + t), acc[2 * k + 1]);
            }
        }
        uint4 o;
        __nv_bfloat162 *oh = (__nv_bfloat162 *)&o;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a = fmaf(acc[2 * k], __ldg(scale + c0 + 2 * k), __ldg(shift + c0 + 2 * k));
            float b = fmaf(acc[2 * k + 1], __ldg(scale + c0 + 2 * k + 1), __ldg(shift + c0 + 2 * k + 1));
            a = a > 0.0f ? a : a * slope;
            b = b > 0.0f ? b : b * slope;
            oh[k] = __floats2bfloat162_rn(interior ? a : 0.0f, interior ? b : 0.0f);
        }
        y[i] = o;
    }
}

__global__ void vy_fill_u32x4_kernel(uint4 *__restrict__ y, long long n16, unsigned v) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) y[i] = make_uint4(v, v, v, v);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#ifndef CV_REL128
#define CV_REL128 1.43     // time per output column of a 128-wide tile relative to a 256-wide one (measured at equal wave
#define CV_REL64 2.36      // counts, tools/conv_bench.py with VY_CONV_BN: 0.112 / 0.185 ms against 0.078 ms at 26^2 256->512)
#define CV_REL_PAIR256 0.88   // CTA pairs, per column and per 128 rows of a CTA, relative to the one-CTA 256-wide tile (equal wave
#define CV_REL_PAIR128 1.24   // counts at 52^2 128->256, VY_CONV_CTA2: 0.074 / 0.104 ms against 0.084 ms)
#endif

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int BN>
int launch_conv(const CUtensorMap &mx, const CUtensorMap &mw, const ConvParams &cp, cudaStream_t st) {
    const size_t smem = (size_t)CV_STAGES * (CV_BM * CV_BK * 2 + BN * CV_BK * 2) + 1024;
    VY_CUDA_CHECK(vy_ensure_dyn_smem((const void *)vy_fusion_conv_kernel<BN>, smem));
    long long grid = cp.n_tiles_total < vy_sm_count() ? cp.n_tiles_total : vy_sm_count();
    cudaError_t le = cudaSuccess;
    VY_KERNEL(VY_K_FUSION_CONV, st, (le = vy_launch(vy_fusion_conv_kernel<BN>, dim3((unsigned)grid), dim3(CV_NT), smem, st, true, mx, mw, cp)));
    if (le != cudaSuccess) VY_FAIL(VY_ECUDA, "launch of vy_fusion_conv_kernel failed: %s", cudaGetErrorString(le));
    return VY_OK;
}

template <int BN>
int launch_conv2(const CUtensorMap &mx, const CUtensorMap &mw, const ConvParams &cp, cudaStream_t st) {
    VY_CUDA_CHECK(vy_ensure_dyn_smem((const void *)vy_fusion_conv2_kernel<BN>, Conv2Cfg<BN>::SMEM));
    const long long pairs_max = vy_sm_count() / 2;
    const long long pairs = cp.n_tiles_total < pairs_max ? cp.n_tiles_total : pairs_max;
    cudaError_t le = cudaSuccess;
    VY_KERNEL(VY_K_FUSION_CONV, st, (le = vy_launch(vy_fusion_conv2_kernel<BN>, dim3((unsigned)(2 * pairs)), dim3(CV_NT),
                                                    Conv2Cfg<BN>::SMEM, st, true, mx, mw, cp)));
    if (le != cudaSuccess) VY_FAIL(VY_ECUDA, "launch of vy_fusion_conv2_kernel failed: %s", cudaGetErrorString(le));
    return VY_OK;
}

}  // namespace

extern "C" size_t vy_p_layout_elems(int B, int T, int H, int W, int C) {
    return (size_t)T * B * (H + 2) * (W + 2) * C;
}

extern "C" int vy_pack_f32_to_p_bf16(const float *x, long long stride_b, long long stride_c, long long stride_t,
                                     int B, int C, int T, int H, int W, void *y_p, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !y_p || B < 1 || C < 1 || T < 1 || H < 1 || W < 1) VY_FAIL(VY_EINVAL, "vy_pack_f32_to_p_bf16: bad arguments");
    if ((long long)T * B > 65535) VY_FAIL(VY_EUNSUPPORTED, "vy_pack_f32_to_p_bf16: T*B must be <= 65535");
    if (((uintptr_t)y_p & 15) != 0) VY_FAIL(VY_EALIGN, "vy_pack_f32_to_p_bf16: y_p must be 16-byte aligned");
    if (C % LY_C == 0 && stride_c == (long long)H * W && H * W <= LYF_MAX_HW && (H * W) % 4 != 0 && stride_b % 4 == 0 &&
        stride_t % 4 == 0 && ((uintptr_t)x & 15) == 0) {
        // (even planes take the 16-byte path of the general kernel)
        VY_KERNEL(VY_K_LAYOUT, st, (vy_pack_flat_kernel<<<dim3(C / LY_C, T * B), LY_NT, 0, st>>>(x, stride_b, stride_t, B, C, H, W,
                                                                                              (__nv_bfloat16 *)y_p)));
        VY_LAUNCH_CHECK("vy_pack_flat_kernel");
        return VY_OK;
    }
    const int fused_border = (C & 7) == 0;          // the pack kernel's 16-byte path also writes the zero border
    if (!fused_border) {
        VY_KERNEL(VY_K_LAYOUT, st, (vy_zero_border_kernel<<<dim3(64, T * B), 128, 0, st>>>((__nv_bfloat16 *)y_p, H, W, C)));
        VY_LAUNCH_CHECK("vy_zero_border_kernel");
    }
    const int vec = ((H * W) % 4 == 0) && (stride_b % 4 == 0) && (stride_c % 4 == 0) && (stride_t % 4 == 0) && (((uintptr_t)x & 15) == 0);
    const dim3 grid((H * W + LY_P - 1) / LY_P, (C + LY_C - 1) / LY_C, T * B);
    VY_KERNEL(VY_K_LAYOUT, st, (vy_pack_kernel<<<grid, LY_NT, 0, st>>>(x, stride_b, stride_c, stride_t, B, C, T, H, W,
                                                                     (__nv_bfloat16 *)y_p, C, 0, vec, fused_border)));
    VY_LAUNCH_CHECK("vy_pack_kernel");
    return VY_OK;
}

extern "C" int vy_pack_f32_split_to_p_bf16(const float *x, long long stride_b, long long stride_c, long long stride_t,
                                           int B, int C, int Cpad, int T, int H, int W, void *y_p, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !y_p || B < 1 || C < 1 || Cpad < C || T < 1 || H < 1 || W < 1) VY_FAIL(VY_EINVAL, "vy_pack_f32_split_to_p_bf16: bad arguments");
    if ((long long)T * B > 65535) VY_FAIL(VY_EUNSUPPORTED, "vy_pack_f32_split_to_p_bf16: T*B must be <= 65535");
    // channels C .. Cpad of every third and the border stay zero: the whole tensor is cleared first
    VY_CUDA_CHECK(cudaMemsetAsync(y_p, 0, (size_t)T * B * (H + 2) * (W + 2) * 3 * Cpad * sizeof(__nv_bfloat16), st));
    const dim3 grid((H * W + LY_P - 1) / LY_P, (C + LY_C - 1) / LY_C, T * B);
    VY_KERNEL(VY_K_LAYOUT, st, (vy_pack_kernel<<<grid, LY_NT, 0, st>>>(x, stride_b, stride_c, stride_t, B, C, T, H, W,
                                                                     (__nv_bfloat16 *)y_p, 3 * Cpad, Cpad, 0, 0)));
    VY_LAUNCH_CHECK("vy_pack_kernel");
    return VY_OK;
}

// y[b][hp][wp][r*(T*C) + t*C + c] = x[t][b][hp][wp][c]: the 'cat' join (reshape (0,-3,-2), yolo3.py:1136) on P-layout
// data, repeated `rep` times along the channels (rep = 2 feeds the split-weight prediction conv).  One thread per
// 16-byte vector of the output; the zero border is copied with the rest.
__global__ void vy_cat_repeat_kernel(const uint4 *__restrict__ x, long long frame8, int T, int C8, int rep, long long npix,
                                     uint4 *__restrict__ y) {
    const int Co8 = rep * T * C8;
    const long long total = npix * Co8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % Co8);
        const long long pix = i / Co8;
        const int tc = v % (T * C8), t = tc / C8, c = tc - t * C8;
        y[i] = x[(long long)t * frame8 + pix * C8 + c];
    }
}

extern "C" int vy_cat_repeat_bf16(const void *x, int B, int T, int H, int W, int C, int rep, void *y, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !y || B < 1 || T < 1 || H < 1 || W < 1 || C < 1 || rep < 1) VY_FAIL(VY_EINVAL, "vy_cat_repeat_bf16: bad arguments");
    if (C % 8 != 0 || (((uintptr_t)x | (uintptr_t)y) & 15) != 0)
        VY_FAIL(VY_EALIGN, "vy_cat_repeat_bf16: C must be a multiple of 8 and x, y 16-byte aligned");
    const long long npix = (long long)B * (H + 2) * (W + 2);
    const long long total = npix * rep * T * (C / 8);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)vy_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    VY_KERNEL(VY_K_LAYOUT, st, (vy_cat_repeat_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint4 *)x, npix * (C / 8), T, C / 8, rep,
                                                                                        npix, (uint4 *)y)));
    VY_LAUNCH_CHECK("vy_cat_repeat_kernel");
    return VY_OK;
}

extern "C" int vy_unpack_p_channels_to_f32(const void *y_p, int p_is_f32, int B, int Cp, int C, int T, int H, int W, float *x,
                                           long long stride_b, long long stride_c, long long stride_t, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !y_p || B < 1 || C < 1 || Cp < C || T < 1 || H < 1 || W < 1) VY_FAIL(VY_EINVAL, "vy_unpack_p_to_f32: bad arguments");
    if ((long long)T * B > 65535) VY_FAIL(VY_EUNSUPPORTED, "vy_unpack_p_to_f32: T*B must be <= 65535");
    if (((uintptr_t)y_p & 15) != 0) VY_FAIL(VY_EALIGN, "vy_unpack_p_to_f32: y_p must be 16-byte aligned");
    const dim3 grid((H * W + LY_P - 1) / LY_P, (C + LY_C - 1) / LY_C, T * B);
    if (p_is_f32) {
        VY_KERNEL(VY_K_LAYOUT, st, (vy_unpack_kernel<float><<<grid, LY_NT, 0, st>>>((const float *)y_p, B, Cp, C, T, H, W, x,
                                                                                   stride_b, stride_c, stride_t)));
    } else {
        VY_KERNEL(VY_K_LAYOUT, st, (vy_unpack_kernel<__nv_bfloat16><<<grid, LY_NT, 0, st>>>(
            (const __nv_bfloat16 *)y_p, B, Cp, C, T, H, W, x, stride_b, stride_c, stride_t)));
    }
    VY_LAUNCH_CHECK("vy_unpack_kernel");
    return VY_OK;
}

extern "C" int vy_unpack_p_to_f32(const void *y_p, int p_is_f32, int B, int C, int T, int H, int W, float *x,
                                  long long stride_b, long long stride_c, long long stride_t, vy_stream_t stream) {
    return vy_unpack_p_channels_to_f32(y_p, p_is_f32, B, C, C, T, H, W, x, stride_b, stride_c, stride_t, stream);
}

extern "C" size_t vy_fusion_conv_workspace_bytes(int B, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw) {
    (void)B; (void)T; (void)H; (void)W; (void)Cin; (void)Cout; (void)kt; (void)kh; (void)kw;
    return 0;      // operands are read in place through TMA; nothing is staged in global memory
}

static int conv_launch(const void *x, const void *w, const float *scale, const float *shift,
                       float leaky_slope, int B, int T, int H, int W, int Cin, int Cout,
                       int kt, int kh, int kw, void *y, int y_is_f32, int nchw_C, int pool_max, vy_stream_t stream,
                       int x_C = 0, int x_T = 0) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !w || !scale || !shift || !y) VY_FAIL(VY_EINVAL, "vy_fusion_conv_bf16: null pointer");
    if (B < 1 || T < 1 || H < 1 || W < 1) VY_FAIL(VY_EINVAL, "vy_fusion_conv_bf16: bad shape");
    if (Cin % 64 != 0 || Cout % 64 != 0 || Cin < 64 || Cout < 64)
        VY_FAIL(VY_EUNSUPPORTED, "vy_fusion_conv_bf16: Cin and Cout must be multiples of 64 (got %d, %d)", Cin, Cout);
    if ((kt != 1 && kt != 3) || (kh != 1 && kh != 3) || (kw != 1 && kw != 3))
        VY_FAIL(VY_EUNSUPPORTED, "vy_fusion_conv_bf16: kernel extents must be 1 or 3");
    if (kt > 1 && T < 2) VY_FAIL(VY_EINVAL, "vy_fusion_conv_bf16: temporal kernel needs T > 1 (yolo3.py:979-985)");
    if ((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) != 0) VY_FAIL(VY_EALIGN, "vy_fusion_conv_bf16: x, w, y must be 16-byte aligned");
    EncodeTiledFn enc = encode_fn();
    if (!enc) VY_FAIL(VY_ECUDA, "vy_fusion_conv_bf16: cuTensorMapEncodeTiled is not available from this driver");

    const int Hp = H + 2, Wp = W + 2;
    const long long rows = (long long)B * Hp * Wp;
    if (rows > 0x7fffff00LL) VY_FAIL(VY_EINVAL, "vy_fusion_conv_bf16: B*Hp*Wp too large");
    // Tile shape.  The widest N tile wastes the least operand traffic, and a CTA pair (256 rows, the weight tile split
    // over the two CTAs) less still -- but the persistent grid runs ceil(tiles / units) waves of whole tiles, units = SMs
    // or SM pairs: 180 tiles of 128 x 256 at 13^2 x 8 windows are two waves, the second 22 % full; 150 pair tiles at
    // 26^2 x 8 windows are three waves on 74 pairs where 294 one-CTA tiles are two on 148 SMs.  Pick the shape with the
    // lowest estimated time = waves x width x measured relative cost per column (one CTA: 128 wide re-reads A twice as
    // often, 1.43x; pairs: 0.88x at 256, 1.24x at 128 -- DESIGN.md section 4.5).
    // VY_CONV_BN = one-CTA width, VY_CONV_CTA2 = 0 (never pairs) | 128 | 256 (that pair width): A/B runs.
    int BN = 64, pairBN = 0;
    {
        static const char *force = getenv("VY_CONV_BN");
        static const char *pe = getenv("VY_CONV_CTA2");
        const int want_pair = pe ? atoi(pe) : -1;
        const int sms = vy_sm_count();
        double best = 1e300;
        const int cand[5] = {256, 128, 64, 256, 128};
        const double rel[5] = {1.0, CV_REL128, CV_REL64, CV_REL_PAIR256, CV_REL_PAIR128};
        for (int i = 0; i < 5; ++i) {
            const bool is_pair = i >= 3;
            if (Cout % cand[i] != 0) continue;
            if (is_pair) {
                if (want_pair == 0 || sms < 2) continue;
                if (want_pair > 0 && want_pair != cand[i] && Cout % want_pair == 0) continue;
            } else {
                if (want_pair > 0 && Cout % want_pair == 0) continue;
                if (force && atoi(force) != cand[i] && Cout % atoi(force) == 0) continue;
            }
            const int bm = is_pair ? CV2_BM : CV_BM;
            const long long units = is_pair ? sms / 2 : sms;
            const long long tiles = (long long)T * ((rows + bm - 1) / bm) * (Cout / cand[i]);
            const long long waves = (tiles + units - 1) / units;
            const double cost = (double)waves * cand[i] * rel[i];
            if (cost < best) { best = cost; BN = cand[i]; pairBN = is_pair ? cand[i] : 0; }
        }
    }
    const int BM = pairBN ? CV2_BM : CV_BM;
    const long long Ktot = (long long)kt * kh * kw * Cin;

    CUtensorMap mx, mw;
    {   // X: (C, rows, T) bf16, box (64, 128, 1)
        // (x_C > 0: the joined 1 x 1 conv -- x holds x_T frames of x_C channels, Cin = the joined, possibly repeated, K)
        const int xc = x_C > 0 ? x_C : Cin, xt = x_C > 0 ? x_T : T;
        cuuint64_t dim[3] = {(cuuint64_t)xc, (cuuint64_t)rows, (cuuint64_t)xt};
        cuuint64_t str[2] = {(cuuint64_t)xc * 2, (cuuint64_t)rows * xc * 2};
        cuuint32_t box[3] = {CV_BK, CV_BM, 1}, es[3] = {1, 1, 1};
        const CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(x), dim, str, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) VY_FAIL(VY_ECUDA, "cuTensorMapEncodeTiled(x) failed: %d", (int)r);
    }
    {   // W: (Ktot, Cout) bf16, box (64, BN)
        cuuint64_t dim[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
        cuuint64_t str[1] = {(cuuint64_t)Ktot * 2};
        cuuint32_t box[2] = {CV_BK, (cuuint32_t)(pairBN ? BN / 2 : BN)}, es[2] = {1, 1};
        const CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w), dim, str, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) VY_FAIL(VY_ECUDA, "cuTensorMapEncodeTiled(w) failed: %d", (int)r);
    }
    ConvParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.T = T; cp.rows = (int)rows; cp.Hp = Hp; cp.Wp = Wp; cp.Cin = Cin; cp.Cout = Cout;
    cp.kt = kt; cp.kh = kh; cp.kw = kw;
    cp.m_tiles = (int)((rows + BM - 1) / BM);
    cp.n_tiles = Cout / BN;
    cp.cin_blocks = Cin / CV_BK;
    cp.n_tiles_total = (long long)T * cp.m_tiles * cp.n_tiles;
    cp.slope = leaky_slope; cp.scale = scale; cp.shift = shift; cp.y = y; cp.y_is_f32 = y_is_f32;
    cp.nchw_C = nchw_C; cp.nchw_H = H; cp.nchw_W = W;
    cp.pool_max = pool_max;
    if (x_C > 0) { cp.a_cpb = x_C / CV_BK; cp.a_frames = x_T; }
    if (pairBN == 256) return launch_conv2<256>(mx, mw, cp, st);
    if (pairBN == 128) return launch_conv2<128>(mx, mw, cp, st);
    if (BN == 256) return launch_conv<256>(mx, mw, cp, st);
    if (BN == 128) return launch_conv<128>(mx, mw, cp, st);
    return launch_conv<64>(mx, mw, cp, st);
}

extern "C" int vy_fusion_conv_bf16(const void *x, const void *w, const float *scale, const float *shift,
                                   float leaky_slope, int B, int T, int H, int W, int Cin, int Cout,
                                   int kt, int kh, int kw, void *y, int y_is_f32, void *workspace,
                                   size_t workspace_bytes, vy_stream_t stream) {
    (void)workspace; (void)workspace_bytes;
    return conv_launch(x, w, scale, shift, leaky_slope, B, T, H, W, Cin, Cout, kt, kh, kw, y, y_is_f32, 0, 0, stream);
}

extern "C" int vy_fusion_conv_bf16_maxpool(const void *x, const void *w, const float *scale, const float *shift,
                                           float leaky_slope, int B, int T, int H, int W, int Cin, int Cout,
                                           int kt, int kh, int kw, void *y, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!y || B < 1 || T < 1 || H < 1 || W < 1 || Cout < 64 || Cout % 64 != 0) VY_FAIL(VY_EINVAL, "vy_fusion_conv_bf16_maxpool: bad arguments");
    if (((uintptr_t)y & 15) != 0) VY_FAIL(VY_EALIGN, "vy_fusion_conv_bf16_maxpool: y must be 16-byte aligned");
    // the pooled frame starts at -inf (bf16 0xFF80); the conv's epilogue raises it
    const long long n16 = (long long)B * (H + 2) * (W + 2) * Cout * 2 / 16;
    long long blocks = (n16 + 255) / 256;
    const long long cap = (long long)vy_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    VY_KERNEL(VY_K_LAYOUT, st, (vy_fill_u32x4_kernel<<<(unsigned)blocks, 256, 0, st>>>((uint4 *)y, n16, 0xFF80FF80u)));
    VY_LAUNCH_CHECK("vy_fill_u32x4_kernel");
    return conv_launch(x, w, scale, shift, leaky_slope, B, T, H, W, Cin, Cout, kt, kh, kw, y, 0, 0, 1, stream);
}

extern "C" int vy_fusion_conv_bf16_nchw(const void *x, const void *w, const float *scale, const float *shift,
                                        float leaky_slope, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                                        float *y, int out_channels, vy_stream_t stream) {
    if (out_channels < 1 || out_channels > Cout) VY_FAIL(VY_EINVAL, "vy_fusion_conv_bf16_nchw: out_channels must be in [1, Cout]");
    if (((uintptr_t)y & 3) != 0) VY_FAIL(VY_EALIGN, "vy_fusion_conv_bf16_nchw: y must be 4-byte aligned");
    return conv_launch(x, w, scale, shift, leaky_slope, B, 1, H, W, Cin, Cout, 1, kh, kw, y, 1, out_channels, 0, stream);
}

extern "C" int vy_fusion_conv_bf16_nchw_joined(const void *x, const void *w, const float *scale, const float *shift,
                                               float leaky_slope, int B, int T, int H, int W, int C, int rep, int Cout,
                                               float *y, int out_channels, vy_stream_t stream) {
    if (T < 1 || rep < 1 || C < 64 || C % 64 != 0) VY_FAIL(VY_EINVAL, "vy_fusion_conv_bf16_nchw_joined: C must be a multiple of 64, T, rep >= 1");
    if (out_channels < 1 || out_channels > Cout) VY_FAIL(VY_EINVAL, "vy_fusion_conv_bf16_nchw_joined: out_channels must be in [1, Cout]");
    if (((uintptr_t)y & 3) != 0) VY_FAIL(VY_EALIGN, "vy_fusion_conv_bf16_nchw_joined: y must be 4-byte aligned");
    return conv_launch(x, w, scale, shift, leaky_slope, B, 1, H, W, rep * T * C, Cout, 1, 1, 1, y, 1, out_channels, 0, stream, C, T);
}

extern "C" int vy_temporal_pool_bf16(const void *x, int T, long inner, int mode, void *y, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !y || T < 1 || inner < 1 || (mode != 0 && mode != 1)) VY_FAIL(VY_EINVAL, "vy_temporal_pool_bf16: bad arguments");
    if (inner % 8 != 0 || (((uintptr_t)x | (uintptr_t)y) & 15) != 0)
        VY_FAIL(VY_EALIGN, "vy_temporal_pool_bf16: inner must be a multiple of 8 elements and pointers 16-byte aligned");
    const long long n8 = inner / 8;
    long long blocks = (n8 + 255) / 256;
    const long long cap = (long long)vy_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    VY_KERNEL(VY_K_TEMPORAL_POOL, st, (vy_temporal_pool_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint4 *)x, T, n8, mode, (uint4 *)y)));
    VY_LAUNCH_CHECK("vy_temporal_pool_kernel");
    return VY_OK;
}

// _upsample(x, 2) + slice_like + channel concat of the YOLO neck (layers.py:11-20, yolo3.py:1170-1177) on P-layout
// data: out[f][y][x] = [ up[f][y/2][x/2] (Cu channels) | route[f][y][x] (Cr channels) ], zero border.  One thread per
// 16-byte vector (8 channels) of the output.
__global__ void vy_upsample_concat_kernel(const uint4 *__restrict__ up, const uint4 *__restrict__ route, int F, int H, int W,
                                          int Hu, int Wu, int Cu8, int Cr8, uint4 *__restrict__ out) {
    const int Hp = H + 2, Wp = W + 2, Co8 = Cu8 + Cr8;
    const long long total = (long long)F * Hp * Wp * Co8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % Co8);
        long long pix = i / Co8;
        const int wp = (int)(pix % Wp);
        pix /= Wp;
        const int hp = (int)(pix % Hp);
        const long long f = pix / Hp;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (hp >= 1 && hp <= H && wp >= 1 && wp <= W) {
            if (v < Cu8) {
                const int hu = (hp - 1) / 2 + 1, wu = (wp - 1) / 2 + 1;       // nearest (pixel repeat), in padded coordinates
                o = up[((f * (Hu + 2) + hu) * (Wu + 2) + wu) * Cu8 + v];
            } else {
                o = route[((f * Hp + hp) * Wp + wp) * Cr8 + (v - Cu8)];
            }
        }
        out[i] = o;
    }
}

extern "C" int vy_upsample_concat_bf16(const void *up, const void *route, int B, int T, int H, int W, int Hu, int Wu,
                                       int Cu, int Cr, void *out, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!up || !route || !out || B < 1 || T < 1 || H < 1 || W < 1 || Hu < 1 || Wu < 1 || Cu < 1 || Cr < 1)
        VY_FAIL(VY_EINVAL, "vy_upsample_concat_bf16: bad arguments");
    if (2 * Hu < H || 2 * Wu < W) VY_FAIL(VY_EINVAL, "vy_upsample_concat_bf16: the upsampled map must cover the route (slice_like only crops)");
    if (Cu % 8 != 0 || Cr % 8 != 0 || (((uintptr_t)up | (uintptr_t)route | (uintptr_t)out) & 15) != 0)
        VY_FAIL(VY_EALIGN, "vy_upsample_concat_bf16: channels must be multiples of 8 and pointers 16-byte aligned");
    const long long total = (long long)B * T * (H + 2) * (W + 2) * ((Cu + Cr) / 8);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)vy_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    VY_KERNEL(VY_K_LAYOUT, st, (vy_upsample_concat_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        (const uint4 *)up, (const uint4 *)route, B * T, H, W, Hu, Wu, Cu / 8, Cr / 8, (uint4 *)out)));
    VY_LAUNCH_CHECK("vy_upsample_concat_kernel");
    return VY_OK;
}

extern "C" int vy_temporal_dwconv_bf16(const void *x, const float *w, const float *scale, const float *shift,
                                       float leaky_slope, int B, int T, int H, int W, int C, void *y, vy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !w || !scale || !shift || !y || B < 1 || T < 1 || H < 1 || W < 1 || C < 1)
        VY_FAIL(VY_EINVAL, "vy_temporal_dwconv_bf16: bad arguments");
    if (C % 8 != 0 || (((uintptr_t)x | (uintptr_t)y) & 15) != 0)
        VY_FAIL(VY_EALIGN, "vy_temporal_dwconv_bf16: C must be a multiple of 8 and x, y 16-byte aligned");
    const long long n8 = (long long)B * (H + 2) * (W + 2) * (C / 8);
    long long blocks = (n8 + 255) / 256;
    const long long cap = (long long)vy_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    VY_KERNEL(VY_K_TEMPORAL_POOL, st, (vy_temporal_dwconv_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        (const uint4 *)x, T, n8, C / 8, H + 2, W + 2, w, scale, shift, leaky_slope, (uint4 *)y)));
    VY_LAUNCH_CHECK("vy_temporal_dwconv_kernel");
    return VY_OK;
}
