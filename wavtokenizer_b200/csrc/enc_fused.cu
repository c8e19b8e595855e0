// Encoder level 0 -> 1 in ONE tcgen05 kernel (reference encoder/modules/seanet.py:123-129 strided conv, :45-63 ResBlock;
// encoder/modules/conv.py:195-211 padding rule):
//
//   x1 = W_d * a + b_d            strided conv 32 -> 64, k = 4, stride 2; `a` = window of the ELU(y0) planes (K = 128)
//   h1 = W_k3 * ELU(x1) + b_1     k3 conv 64 -> 32 with reflect padding 1 / 1
//   y1 = W_1x1 * ELU(h1) + (W_sc W_d) * a + b_2      1x1 conv 32 -> 64 + shortcut composed with the strided conv
//   out = ELU(y1) as split-fp16 planes in the reflect-padded layout of the next strided conv
//
// The unfused path runs three GEMM launches for this and moves x1 / ELU(x1) / ELU(h1) / the y0 windows a second time
// through HBM (3.4 GB per 64 clips); here a persistent CTA keeps all weights resident in shared memory (96 KB), reads a
// tile of `a` once (TMA) and writes ELU(y1) once (1.1 GB per 64 clips). Intermediates live in TMEM and shared memory:
//
//   GEMM1  [a_hi, a_lo] x [Wc_lo; Wc_hi; Wd_hi; Wd_lo]  -> TMEM  [sc_hl | sc_hh+lh | x1_hh+lh | x1_hl]   (N = 256 and 128)
//   ep 1   x1 + b_d -> ELU -> split -> shared-memory tile E (128-byte swizzle, K-major: the A operand of GEMM2)
//   GEMM2  E x [Wk3_hi; Wk3_lo] with the three taps as COLUMN blocks  -> P[r, tap*32 + c]               (N = 192 and 96)
//   ep 2   h1[r] = P0[r-1] + P1[r] + P2[r+1] (+ reflect at the clip ends) -> ELU -> split -> tile A2 (64-byte swizzle)
//   GEMM3  A2 x [W1x1_lo; W1x1_hi] accumulated onto the sc columns                                        (N = 128 and 64)
//   ep 3   y1 + b_2 -> ELU -> split planes -> HBM (row re-map + mirrored halo rows, as the generic GEMM epilogue)
//
// The row shift of the k3 conv happens between ACCUMULATOR rows (warp shuffles), not between operand rows: a tile is four
// independent groups of 32 rows (one per TMEM lane quarter = one epilogue warp row range), each loaded with its own
// one-row halo, so 30 of every 32 rows produce output and no shift ever crosses a warp. All three products use the
// 3-pass split-fp16 scheme of gemm_tc.cu (hi*hi + hi*lo + lo*hi, fp32 accumulate).
// TMEM: two regions of 256 columns; tile i keeps GEMM1 / GEMM3 in region i & 1 and P in the other one, so the last
// epilogue of tile i overlaps GEMM1 of tile i + 1.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace wt {

namespace {

constexpr int F_QROWS = 30;                 // output rows per lane quarter (32 rows loaded)
constexpr int F_TILE = 4 * F_QROWS;         // output rows per tile
constexpr int F_THREADS = 64 + 16 * 32;     // warp 0 TMA, warp 1 MMA, 16 epilogue warps (4 per lane quarter)

constexpr uint32_t F_W1 = 0;                         // 2 k-blocks x [256 rows x 128 B]
constexpr uint32_t F_W1_KB = 256 * 128;
constexpr uint32_t F_W2 = F_W1 + 2 * F_W1_KB;        // [192 rows x 128 B]
constexpr uint32_t F_W3 = F_W2 + 192 * 128;          // [128 rows x 64 B]
constexpr uint32_t F_A0 = F_W3 + 128 * 64;           // 2 planes x 2 k-blocks x [128 rows x 128 B]
constexpr uint32_t F_A0_KB = 128 * 128;
constexpr uint32_t F_A0_PLANE = 2 * F_A0_KB;
constexpr uint32_t F_E = F_A0 + 2 * F_A0_PLANE;      // 2 planes x [128 rows x 128 B]
constexpr uint32_t F_E_PLANE = 128 * 128;
constexpr uint32_t F_A2 = F_E + 2 * F_E_PLANE;       // 2 planes x [128 rows x 64 B]
constexpr uint32_t F_A2_PLANE = 128 * 64;
constexpr uint32_t F_BAR = F_A2 + 2 * F_A2_PLANE;    // 9 mbarriers + TMEM slot
constexpr uint32_t F_BIAS = F_BAR + 128;             // b_d[64] | b_1[32] | b_2[64]
constexpr uint32_t F_SMEM = F_BIAS + 160 * 4 + 1024; // + alignment slack
constexpr uint32_t F_W_BYTES = 2 * F_W1_KB + 192 * 128 + 128 * 64;
static_assert(F_W2 % 1024 == 0 && F_W3 % 1024 == 0 && F_A0 % 1024 == 0 && F_E % 1024 == 0 && F_A2 % 1024 == 0,
              "swizzled tiles start on 1024-byte boundaries");
static_assert(F_SMEM <= 227 * 1024, "shared memory budget");

struct FArgs {
    const float* bias;
    int Mtot, n_tiles, T1;
    RowMap map;
    __half* ye_hi;
    __half* ye_lo;
    float* y_f32;
    long long* dbg;
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

__global__ void __launch_bounds__(F_THREADS, 1)
enc_l1_fused_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                    const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                    const __grid_constant__ CUtensorMap mapW3, const FArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
    auto bar = [&](int i) { return sbase + F_BAR + 8u * i; };
    enum { B_W = 0, B_A0_FULL, B_A0_EMPTY, B_G1, B_G2, B_G3, B_E1, B_E2, B_E3 };
    const uint32_t tmem_slot = sbase + F_BAR + 80;
    const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(sptr + F_BAR + 80);
    float* sbias = reinterpret_cast<float*>(sptr + F_BIAS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#if WT_TIMELINE
    long long* dbg = (a.dbg && blockIdx.x == 0) ? a.dbg : nullptr;
    const long long t_begin = dbg ? clock64() : 0;
    auto stamp = [&](int it, int slot) { if (dbg && it >= 2 && it < 6) dbg[(it - 2) * 16 + slot] = clock64() - t_begin; };
#else
    auto stamp = [](int, int) {};
#endif

    for (int i = threadIdx.x; i < 160; i += F_THREADS) sbias[i] = a.bias[i];
    if (warp == 0 && lane == 0) {
        mbar_init(bar(B_W), 1);
        mbar_init(bar(B_A0_FULL), 1);
        mbar_init(bar(B_A0_EMPTY), 1);
        mbar_init(bar(B_G1), 1);
        mbar_init(bar(B_G2), 1);
        mbar_init(bar(B_G3), 1);
        mbar_init(bar(B_E1), 16);
        mbar_init(bar(B_E2), 16);
        mbar_init(bar(B_E3), 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_expect_tx(bar(B_W), F_W_BYTES);
            tma_load_2d(sbase + F_W1, &mapW1, 0, 0, bar(B_W));
            tma_load_2d(sbase + F_W1 + F_W1_KB, &mapW1, 64, 0, bar(B_W));
            tma_load_2d(sbase + F_W2, &mapW2, 0, 0, bar(B_W));
            tma_load_2d(sbase + F_W3, &mapW3, 0, 0, bar(B_W));
            int it = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
                mbar_wait(bar(B_A0_EMPTY), (uint32_t)(it & 1) ^ 1u);
                mbar_expect_tx(bar(B_A0_FULL), 2 * F_A0_PLANE);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int row = tile * F_TILE + q * F_QROWS - 1;  // one halo row in front of the quarter's 30 outputs
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint32_t dst = sbase + F_A0 + kb * F_A0_KB + q * 4096;
                        tma_load_2d(dst, &mapA_hi, kb * 64, row, bar(B_A0_FULL));
                        tma_load_2d(dst + F_A0_PLANE, &mapA_lo, kb * 64, row, bar(B_A0_FULL));
                    }
                }
                stamp(it, 0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint64_t d128 = umma_desc_hi(64), d64 = umma_desc_hi(32);
            constexpr uint32_t i256 = umma_idesc_f16(256), i192 = umma_idesc_f16(192), i128 = umma_idesc_f16(128),
                               i96 = umma_idesc_f16(96), i64 = umma_idesc_f16(64);
            mbar_wait(bar(B_W), 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
                const uint32_t ph = (uint32_t)(it & 1);
                const uint32_t Rp = tmem_base + ph * 256u, Rq = tmem_base + (ph ^ 1u) * 256u;
                // ---- GEMM1: strided conv + composed shortcut ----
                mbar_wait(bar(B_A0_FULL), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                stamp(it, 1);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t koff = k * 32;
                        const uint32_t sa = sbase + F_A0 + kb * F_A0_KB + koff;
                        const uint32_t sw = sbase + F_W1 + kb * F_W1_KB + koff;
                        umma_f16(Rp, umma_desc_at(d128, sa), umma_desc_at(d128, sw), i256, (kb | k) != 0);
                        umma_f16(Rp + 64, umma_desc_at(d128, sa + F_A0_PLANE), umma_desc_at(d128, sw + 64 * 128), i128, 1);
                    }
                }
                umma_commit(bar(B_A0_EMPTY));
                umma_commit(bar(B_G1));
                // ---- GEMM2: k3 conv, taps as column blocks ----
                mbar_wait(bar(B_E1), ph);
                if (it > 0) mbar_wait(bar(B_E3), ph ^ 1u);  // y of the previous tile (region Rq) has been drained
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                stamp(it, 2);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t koff = k * 32;
                    const uint64_t w = umma_desc_at(d128, sbase + F_W2 + koff);
                    umma_f16(Rq, umma_desc_at(d128, sbase + F_E + koff), w, i192, k != 0);
                    umma_f16(Rq, umma_desc_at(d128, sbase + F_E + F_E_PLANE + koff), w, i96, 1);
                }
                umma_commit(bar(B_G2));
                // ---- GEMM3: 1x1 conv onto the shortcut columns ----
                mbar_wait(bar(B_E2), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                stamp(it, 3);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint32_t koff = k * 32;
                    const uint32_t sw = sbase + F_W3 + koff;
                    umma_f16(Rp, umma_desc_at(d64, sbase + F_A2 + koff), umma_desc_at(d64, sw), i128, 1);
                    umma_f16(Rp + 64, umma_desc_at(d64, sbase + F_A2 + F_A2_PLANE + koff), umma_desc_at(d64, sw + 64 * 64), i64, 1);
                }
                umma_commit(bar(B_G3));
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3;            // TMEM lane quarter (hardware: warp id % 4)
        const int cg = (warp - 2) >> 2;    // column share 0..3
        const int row = q * 32 + lane;     // tile row = TMEM lane
        const int Pin = a.map.Pin, T1 = a.T1;
        int it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const uint32_t ph = (uint32_t)(it & 1);
            const uint32_t lane_off = (uint32_t)(q * 32) << 16;
            const uint32_t Rp = tmem_base + lane_off + ph * 256u, Rq = tmem_base + lane_off + (ph ^ 1u) * 256u;
            const int m = tile * F_TILE + q * F_QROWS - 1 + lane;
            const int bq = m >= 0 ? m / Pin : 0;
            const int t = m - bq * Pin;
            const bool out_ok = lane >= 1 && lane <= F_QROWS && m >= 0 && m < a.Mtot && t < T1;
            // ---- epilogue 1: ELU(x1) -> tile E ----
            mbar_wait(bar(B_G1), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (threadIdx.x == 64) stamp(it, 4);
            {
                uint32_t r[16], r2[16];
                tmem_ld_pair(Rp + 128 + cg * 16, r, Rp + 192 + cg * 16, r2);
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float v0 = elu1(__uint_as_float(r[2 * i]) + __uint_as_float(r2[2 * i]) + sbias[cg * 16 + 2 * i]);
                    const float v1 = elu1(__uint_as_float(r[2 * i + 1]) + __uint_as_float(r2[2 * i + 1]) + sbias[cg * 16 + 2 * i + 1]);
                    split2(v0, v1, hi[i], lo[i]);
                }
                const uint32_t rb = sbase + F_E + row * 128;
                const uint32_t c0 = (uint32_t)((2 * cg) ^ (row & 7)) * 16, c1 = (uint32_t)((2 * cg + 1) ^ (row & 7)) * 16;
                sts128(rb + c0, hi[0], hi[1], hi[2], hi[3]);
                sts128(rb + c1, hi[4], hi[5], hi[6], hi[7]);
                sts128(rb + F_E_PLANE + c0, lo[0], lo[1], lo[2], lo[3]);
                sts128(rb + F_E_PLANE + c1, lo[4], lo[5], lo[6], lo[7]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core (async proxy) reads
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_E1));
            // ---- epilogue 2: h1 = P0[r-1] + P1[r] + P2[r+1] -> ELU -> tile A2 ----
            mbar_wait(bar(B_G2), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (threadIdx.x == 64) stamp(it, 5);
            {
                uint32_t a0[8], a1[8], b0[8], b1[8], c0[8], c1[8];
                tmem_ld8_nowait(Rq + 0 + cg * 8, a0);
                tmem_ld8_nowait(Rq + 96 + cg * 8, a1);
                tmem_ld8_nowait(Rq + 32 + cg * 8, b0);
                tmem_ld8_nowait(Rq + 128 + cg * 8, b1);
                tmem_ld8_nowait(Rq + 64 + cg * 8, c0);
                tmem_ld8_nowait(Rq + 160 + cg * 8, c1);
                asm volatile("tcgen05.wait::ld.sync.aligned;"
                             : "+r"(a0[0]), "+r"(a0[1]), "+r"(a0[2]), "+r"(a0[3]), "+r"(a0[4]), "+r"(a0[5]), "+r"(a0[6]), "+r"(a0[7]),
                               "+r"(a1[0]), "+r"(a1[1]), "+r"(a1[2]), "+r"(a1[3]), "+r"(a1[4]), "+r"(a1[5]), "+r"(a1[6]), "+r"(a1[7]),
                               "+r"(b0[0]), "+r"(b0[1]), "+r"(b0[2]), "+r"(b0[3]), "+r"(b0[4]), "+r"(b0[5]), "+r"(b0[6]), "+r"(b0[7]),
                               "+r"(b1[0]), "+r"(b1[1]), "+r"(b1[2]), "+r"(b1[3]), "+r"(b1[4]), "+r"(b1[5]), "+r"(b1[6]), "+r"(b1[7])
                             :
                             : "memory");
                asm volatile("" : "+r"(c0[0]), "+r"(c0[1]), "+r"(c0[2]), "+r"(c0[3]), "+r"(c0[4]), "+r"(c0[5]), "+r"(c0[6]), "+r"(c0[7]),
                                  "+r"(c1[0]), "+r"(c1[1]), "+r"(c1[2]), "+r"(c1[3]), "+r"(c1[4]), "+r"(c1[5]), "+r"(c1[6]), "+r"(c1[7])
                             :
                             : "memory");
                // reflect padding of the k3 conv at the clip ends (conv.py:200-210): position -1 reads 1, T1 reads T1 - 2
                const bool first = t == 0, last = t == T1 - 1;
                float h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float p0 = __uint_as_float(a0[i]) + __uint_as_float(a1[i]);
                    const float p1 = __uint_as_float(b0[i]) + __uint_as_float(b1[i]);
                    const float p2 = __uint_as_float(c0[i]) + __uint_as_float(c1[i]);
                    const float p0u = __shfl_up_sync(0xffffffffu, p0, 1), p0d = __shfl_down_sync(0xffffffffu, p0, 1);
                    const float p2u = __shfl_up_sync(0xffffffffu, p2, 1), p2d = __shfl_down_sync(0xffffffffu, p2, 1);
                    h[i] = elu1((first ? p0d : p0u) + p1 + (last ? p2u : p2d) + sbias[64 + cg * 8 + i]);
                }
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split2(h[2 * i], h[2 * i + 1], hi[i], lo[i]);
                const uint32_t rb = sbase + F_A2 + row * 64 + (uint32_t)(cg ^ ((row >> 1) & 3)) * 16;
                sts128(rb, hi[0], hi[1], hi[2], hi[3]);
                sts128(rb + F_A2_PLANE, lo[0], lo[1], lo[2], lo[3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_E2));
            // ---- epilogue 3: ELU(y1) planes -> HBM ----
            mbar_wait(bar(B_G3), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (threadIdx.x == 64) stamp(it, 6);
            {
                uint32_t r[16], r2[16];
                tmem_ld_pair(Rp + 64 + cg * 16, r, Rp + 0 + cg * 16, r2);
                if (out_ok) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        v[i] = __uint_as_float(r[i]) + __uint_as_float(r2[i]) + sbias[96 + cg * 16 + i];
                    const long long sb = a.map.sb ? a.map.sb : a.map.Pout, st = a.map.st ? a.map.st : 1;
                    const long long base = (long long)bq * sb + (long long)a.map.off * st;
                    const long long dst = base + t * st;
                    long long mir_l = -1, mir_r = -1;
                    if (t >= 1 && t <= a.map.hl) mir_l = base - t * st;
                    if (t <= T1 - 2 && t >= T1 - 1 - a.map.hr) mir_r = base + (2 * (T1 - 1) - t) * st;
                    if (a.y_f32) {
                        float* o = a.y_f32 + dst * 64 + cg * 16;
#pragma unroll
                        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (mir_l >= 0) {
                            o = a.y_f32 + mir_l * 64 + cg * 16;
#pragma unroll
                            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        }
                        if (mir_r >= 0) {
                            o = a.y_f32 + mir_r * 64 + cg * 16;
#pragma unroll
                            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        }
                    }
                    store_planes<16, true>(a.ye_hi, a.ye_lo, dst * 64 + cg * 16, v);
                    if (mir_l >= 0) store_planes<16, true>(a.ye_hi, a.ye_lo, mir_l * 64 + cg * 16, v);
                    if (mir_r >= 0) store_planes<16, true>(a.ye_hi, a.ye_lo, mir_r * 64 + cg * 16, v);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_E3));
            if (threadIdx.x == 64) stamp(it, 7);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

bool enc_l1_fused_supported(int cin, int stride) {
    static const bool off = [] { const char* e = std::getenv("WT_ENC_L1_FUSED"); return e && std::atoi(e) == 0; }();
    return !off && cin == 32 && stride == 2;
}

void launch_enc_l1_fused(const EncL1Weights& w, const EncL1Args& a, cudaStream_t s) {
    if (a.Bc <= 0 || a.T1 < 2) return;
    static PerDevice<bool> attr_dev;
    bool& attr = attr_dev.get();
    if (!attr) {
        WT_CUDA(cudaFuncSetAttribute(enc_l1_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
        attr = true;
    }
    const int Pin = a.T1 + 2;
    if (a.map.Pin != Pin || a.map.Tvalid != a.T1) throw Error(4, "enc_l1_fused: row map must describe T1 + 2 slots per clip");
    const long long Mtot = (long long)a.Bc * Pin;
    if (Mtot > (1LL << 30)) throw Error(4, "enc_l1_fused: chunk too large");
    const long long rowsA = a.y0_elems >= 128 ? (a.y0_elems - 128) / 64 + 1 : 0;
    const CUtensorMap mA_hi = tc_make_map(a.y0_hi, rowsA, 128, 64, 32, 64);
    const CUtensorMap mA_lo = tc_make_map(a.y0_lo, rowsA, 128, 64, 32, 64);
    const CUtensorMap mW1 = tc_make_map(w.w1, 256, 128, 128, 256, 64);
    const CUtensorMap mW2 = tc_make_map(w.w2, 192, 64, 64, 192, 64);
    const CUtensorMap mW3 = tc_make_map(w.w3, 128, 32, 32, 128, 32);
    FArgs f;
    f.bias = w.bias; f.Mtot = (int)Mtot; f.n_tiles = (int)((Mtot + F_TILE - 1) / F_TILE); f.T1 = a.T1;
    f.map = a.map; f.ye_hi = a.ye_hi; f.ye_lo = a.ye_lo; f.y_f32 = a.y_f32;
    f.dbg = tc_debug_timeline() ? tc_debug_timeline() + 148 * 64 + 64 : nullptr;
    const int grid = std::min(f.n_tiles, tc_num_sms());
    enc_l1_fused_kernel<<<grid, F_THREADS, F_SMEM, s>>>(mA_hi, mA_lo, mW1, mW2, mW3, f);
    WT_CUDA(cudaGetLastError());
}

}  // namespace wt
