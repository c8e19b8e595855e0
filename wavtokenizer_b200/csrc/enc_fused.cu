// Encoder level 0 -> 1 in ONE tcgen05 kernel (reference encoder/modules/seanet.py:123-129 strided conv, :45-63 ResBlock;
// encoder/modules/conv.py:195-211 padding rule):
//
//   x1 = W_d * a + b_d            strided conv 32 -> 64, k = 2 * stride, stride 2 or 4; `a` = window of the ELU(y0) planes (K = 128 / 256)
//   h1 = W_k3 * ELU(x1) + b_1     k3 conv 64 -> 32 with reflect padding 1 / 1
//   y1 = W_1x1 * ELU(h1) + (W_sc W_d) * a + b_2      1x1 conv 32 -> 64 + shortcut composed with the strided conv
//   out = ELU(y1) as split-fp16 planes in the reflect-padded layout of the next strided conv
//
// The unfused path runs three GEMM launches for this and moves x1 / ELU(x1) / ELU(h1) / the y0 windows a second time
// through HBM (3.4 GB per 64 clips at stride 2); here a persistent CTA keeps all weights resident in shared memory (96 /
// 160 KB), reads a tile of `a` once (TMA) and writes ELU(y1) once (1.1 GB per 64 clips). Everything in between lives in
// TENSOR MEMORY: the accumulators, and the A operands of the second and third product, which the epilogue warps write
// back with tcgen05.st as packed fp16 pairs (A-from-TMEM form of tcgen05.mma: rows = lanes, a k-step = 8 columns):
//
//   GEMM1  [a_hi, a_lo] (smem) x [Wc; Wd]                     -> columns [sc | x1]            (N = 128, K = 128 / 256)
//   ep 1   x1 + b_d -> ELU -> split -> E = [hi | lo] per k-step, IN PLACE over the x1 columns
//   GEMM2  E (TMEM) x Wk3 with the three taps as COLUMN blocks -> P[r, tap*32 + c]             (N = 96, K = 64)
//   ep 2   h1[r] = P0[r-1] + P1[r] + P2[r+1] (+ reflect at the clip ends) -> ELU -> split -> A2 in place over P
//   GEMM3  A2 (TMEM) x W1x1 accumulated onto the sc columns                                    (N = 64, K = 32)
//   ep 3   y1 + b_2 -> ELU -> split planes -> HBM (row re-map + mirrored halo rows, as the generic GEMM epilogue)
//
// The row shift of the k3 conv happens between ACCUMULATOR rows (warp shuffles), not between operand rows: a tile is four
// independent groups of 32 rows (one per TMEM lane quarter = the rows one epilogue warp can read), each loaded with its
// own one-row halo, so 30 of every 32 rows produce output and no shift ever crosses a warp. All three products use the
// 3-pass split-fp16 scheme of gemm_tc.cu (hi*hi + hi*lo + lo*hi, fp32 accumulate).
// TWO tiles are in flight: tile slot s = (tile number of this CTA) & 1 owns 224 TMEM columns and one group of 8 epilogue
// warps; the single MMA-issuing thread polls the barriers of both slots and issues whichever product is ready, so the
// tensor pipe works on one tile while the epilogue warps of the other one run (v1 of this kernel ran the chain of one tile at
// a time: 7.2 k cycles per tile, of which the tensor pipe was busy 2.3 k).
// The window tile streams through a RING of 32 KB k-block stages (hi + lo plane of 64 window elements each) that GEMM1
// consumes in tile order, one k-block per commit: 4 stages beside 96 KB of weights for stride 2 (K = 128: small-320 /
// medium), 2 stages beside 160 KB for stride 4 (K = 256: small-600).
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
constexpr int F_THREADS = 64 + 16 * 32;     // warp 0 TMA, warp 1 MMA, 2 groups of 8 epilogue warps (2 per lane quarter)

// Shared-memory layout for a window of KB0 k-blocks of 64 elements (stride 2: K = 128, KB0 = 2; stride 4: K = 256, KB0 = 4):
// resident weights, then a ring of two stages of the window tile (hi + lo planes). Both layouts fill 224 KB: 96 KB of
// weights + two 64 KB stages (a whole window each), or 160 KB + two 32 KB stages (one k-block each).
template <int KB0>
struct FL {
    static constexpr uint32_t W1_KB = 256 * 128;              // one k-block of W1: rows [Wc_hi | Wd_hi | Wc_lo | Wd_lo] x 128 B
    static constexpr uint32_t W1 = 0;
    static constexpr uint32_t W2 = KB0 * W1_KB;               // [192 rows x 128 B]: rows [Wk3_hi | Wk3_lo]
    static constexpr uint32_t W3 = W2 + 192 * 128;            // [128 rows x 64 B]: rows [W1x1_hi | W1x1_lo]
    static constexpr uint32_t RING = W3 + 128 * 64;
    static constexpr uint32_t PLANE = 128 * 128;              // [128 rows x 128 B] of one plane
    static constexpr uint32_t KBLK = 2 * PLANE;               // one k-block of the window tile: hi + lo plane
    static constexpr int KPS = KB0 == 2 ? 2 : 1;              // k-blocks per ring stage (stride 2: a stage is a whole window:
                                                              // one barrier round per tile measured 5 % faster than two)
    static constexpr uint32_t STG = KPS * KBLK;
    static constexpr int NST = 2;
    static constexpr uint32_t BAR = RING + NST * STG;         // W, full[NST], empty[NST], per slot {G1, G2, G3, E1, E2, E3}
    static constexpr int NBAR = 1 + 2 * NST + 12;
    static constexpr uint32_t TSLOT = BAR + 8 * NBAR;
    static constexpr uint32_t BIAS = TSLOT + 16;              // b_d[64] | b_1[32] | b_2[64]
    static constexpr uint32_t SMEM = BIAS + 160 * 4 + 1024;   // + alignment slack
    static constexpr uint32_t W_BYTES = KB0 * W1_KB + 192 * 128 + 128 * 64;
    static_assert(W2 % 1024 == 0 && W3 % 1024 == 0 && RING % 1024 == 0, "swizzled tiles start on 1024-byte boundaries");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// TMEM columns of a tile slot (two slots of 256 columns)
constexpr uint32_t C_SC = 0;     // [0, 64)   shortcut accumulators, then y1
constexpr uint32_t C_X1 = 64;    // [64, 128) x1 accumulators, then E: k-step k = [hi 8 cols | lo 8 cols] at 64 + 16 k
constexpr uint32_t C_P = 128;    // [128, 224) P = three tap blocks of 32, then A2: k-step k at 128 + 16 k

template <int V>
struct IC { static constexpr int value = V; };

struct FArgs {
    const float* bias;
    int Mtot, n_tiles, T1;
    RowMap map;
    __half* ye_hi;
    __half* ye_lo;
    float* y_f32;
    long long* dbg;
};

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
__device__ __forceinline__ void pin8(uint32_t (&r)[8]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) : : "memory");
}

template <int KB0>
__global__ void __launch_bounds__(F_THREADS, 1)
enc_l1_fused_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                    const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                    const __grid_constant__ CUtensorMap mapW3, const FArgs a) {
    using L = FL<KB0>;
    constexpr int NST = L::NST;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
    enum { G1 = 0, G2, G3, E1, E2, E3 };
    auto bar = [&](int s, int i) { return sbase + L::BAR + 8u * (1 + 2 * NST + 6 * s + i); };
    auto bar_full = [&](int i) { return sbase + L::BAR + 8u * (1 + i); };
    auto bar_empty = [&](int i) { return sbase + L::BAR + 8u * (1 + NST + i); };
    const uint32_t bar_w = sbase + L::BAR;
    const uint32_t tmem_slot = sbase + L::TSLOT;
    const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(sptr + L::TSLOT);
    float* sbias = reinterpret_cast<float*>(sptr + L::BIAS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_local = a.n_tiles > (int)blockIdx.x ? (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
#if WT_TIMELINE
    long long* dbg = (a.dbg && blockIdx.x == 0) ? a.dbg : nullptr;
    const long long t_begin = dbg ? clock64() : 0;
    auto stamp = [&](int it, int slot) { if (dbg && it >= 4 && it < 8) dbg[(it - 4) * 16 + slot] = clock64() - t_begin; };
#else
    auto stamp = [](int, int) {};
#endif

    for (int i = threadIdx.x; i < 160; i += F_THREADS) sbias[i] = a.bias[i];
    if (warp == 0 && lane == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < NST; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(s, G1), 1);
            mbar_init(bar(s, G2), 1);
            mbar_init(bar(s, G3), 1);
            mbar_init(bar(s, E1), 8);
            mbar_init(bar(s, E2), 8);
            mbar_init(bar(s, E3), 8);
        }
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
            mbar_expect_tx(bar_w, L::W_BYTES);
#pragma unroll
            for (int kb = 0; kb < KB0; ++kb) tma_load_2d(sbase + L::W1 + kb * L::W1_KB, &mapW1, kb * 64, 0, bar_w);
            tma_load_2d(sbase + L::W2, &mapW2, 0, 0, bar_w);
            tma_load_2d(sbase + L::W3, &mapW3, 0, 0, bar_w);
            uint32_t n = 0;  // ring stages issued so far: stage n % NST, phase (n / NST) & 1
            for (int it = 0; it < n_local; ++it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
#pragma unroll 1
                for (int kg = 0; kg < KB0 / L::KPS; ++kg, ++n) {
                    const uint32_t st = n % NST, ph = (n / NST) & 1u;
                    mbar_wait(bar_empty(st), ph ^ 1u);
                    mbar_expect_tx(bar_full(st), L::STG);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int row = tile * F_TILE + q * F_QROWS - 1;  // one halo row in front of the quarter's 30 outputs
#pragma unroll
                        for (int j = 0; j < L::KPS; ++j) {  // the k-blocks of a row group back to back: k-block j + 1 of window m
                            const int kb = kg * L::KPS + j;  // is k-block j of window m + 1, i.e. mostly the lines just requested
                            const uint32_t dst = sbase + L::RING + st * L::STG + j * L::KBLK + q * 4096;
                            tma_load_2d(dst, &mapA_hi, kb * 64, row, bar_full(st));
                            tma_load_2d(dst + L::PLANE, &mapA_lo, kb * 64, row, bar_full(st));
                        }
                    }
                }
                stamp(it, 0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread, both tile slots, whichever product is ready =====================
        if (elect_one()) {
            const uint64_t d128 = umma_desc_hi(64), d64 = umma_desc_hi(32);
            constexpr uint32_t i128 = umma_idesc_f16(128), i96 = umma_idesc_f16(96), i64 = umma_idesc_f16(64);
            mbar_wait(bar_w, 0);
            int it_s[2] = {0, 1};   // next tile (local index) of each slot
            int stage[2] = {0, 0};  // 0: GEMM1 of the tile not fully issued yet, 1: GEMM2 next, 2: GEMM3 next
            int it1 = 0, kb1 = 0;   // GEMM1: next tile in ring order and its next k-block
            uint32_t idle = 0;
            while (it_s[0] < n_local || it_s[1] < n_local) {
                bool progress = false;
                // ---- GEMM1: strided conv + composed shortcut, one k-block of the window per turn (A from the ring) ----
#pragma unroll
                for (int s = 0; s < 2; ++s) {  // fully unrolled: the per-slot state and every address stay compile-time
                    const int it = it_s[s];
                    if (it >= n_local) continue;
                    const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                    const uint32_t T = tmem_base + 256u * s;
                    if (stage[s] == 0) {
                        // ---- GEMM1: strided conv + composed shortcut (A from the ring, consumed in tile order) ----
                        constexpr int NG = KB0 / L::KPS;  // ring stages per tile: global stage index n = it * NG + kg
                        // the MMA-issuing thread is a serial resource: every address below is a compile-time constant
                        auto issue = [&](auto KBASE, const uint32_t stc) {
                            constexpr int kbase = decltype(KBASE)::value;
#pragma unroll
                            for (int j = 0; j < L::KPS; ++j) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint32_t koff = k * 32;
                                    const uint32_t sa = sbase + L::RING + stc * L::STG + j * L::KBLK + koff;
                                    const uint32_t sw = sbase + L::W1 + (kbase + j) * L::W1_KB + koff;
                                    const uint64_t a_hi = umma_desc_at(d128, sa), a_lo = umma_desc_at(d128, sa + L::PLANE);
                                    const uint64_t w_hi = umma_desc_at(d128, sw), w_lo = umma_desc_at(d128, sw + 128 * 128);
                                    umma_f16(T + C_SC, a_hi, w_hi, i128, (kbase | j | k) != 0);
                                    umma_f16(T + C_SC, a_hi, w_lo, i128, 1);
                                    umma_f16(T + C_SC, a_lo, w_hi, i128, 1);
                                }
                            }
                        };
                        if constexpr (NG == 1) {
                            // each slot owns a stage (stage = it & 1 = s, phase = ph): the slots' GEMM1s may overtake each other
                            if (!mbar_test(bar_full(s), ph)) continue;
                            if (it >= 2 && !mbar_test(bar(s, E3), ph ^ 1u)) continue;  // the slot's previous tile has been drained
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            stamp(it, 1);
                            issue(IC<0>{}, (uint32_t)s);  // s is a constant after the unrolling of the slot loop
                            umma_commit(bar_empty(s));
                            umma_commit(bar(s, G1));
                            stage[s] = 1;
                            progress = true;
                        } else {
                            // the two stages are shared by the slots: the ring is consumed in tile order, one k-block per stage
                            if (it != it1) continue;
#pragma unroll 1
                            for (int turn = 0; turn < NG; ++turn) {
                                const int kg = kb1;  // KPS == 1
                                const uint32_t st = (uint32_t)(kg & 1), php = (uint32_t)((kg >> 1) & 1);  // n = 4 it + kg, two stages
                                if (!mbar_test(bar_full(st), php)) break;
                                if (kg == 0 && it >= 2 && !mbar_test(bar(s, E3), ph ^ 1u)) break;
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                if (kg == 0) stamp(it, 1);
                                switch (kg) {
                                    case 0: issue(IC<0>{}, 0u); break;
                                    case 1: issue(IC<1>{}, 1u); break;
                                    case 2: issue(IC<2>{}, 0u); break;
                                    default: issue(IC<3>{}, 1u); break;
                                }
                                umma_commit(bar_empty(st));
                                progress = true;
                                if (++kb1 == KB0) {
                                    umma_commit(bar(s, G1));
                                    kb1 = 0;
                                    ++it1;
                                    stage[s] = 1;
                                    break;
                                }
                            }
                        }
                        if (stage[s] == 0) continue;
                    }
                    if (stage[s] == 1) {
                        // ---- GEMM2: k3 conv, taps as column blocks (A = E from tensor memory) ----
                        if (!mbar_test(bar(s, E1), ph)) continue;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        stamp(it, 2);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t sw = sbase + L::W2 + k * 32;
                            const uint64_t w_hi = umma_desc_at(d128, sw), w_lo = umma_desc_at(d128, sw + 96 * 128);
                            const uint32_t e_hi = T + C_X1 + 16 * k, e_lo = e_hi + 8;
                            umma_f16_ts(T + C_P, e_hi, w_hi, i96, k != 0);
                            umma_f16_ts(T + C_P, e_hi, w_lo, i96, 1);
                            umma_f16_ts(T + C_P, e_lo, w_hi, i96, 1);
                        }
                        umma_commit(bar(s, G2));
                        stage[s] = 2;
                        progress = true;
                    } else {
                        // ---- GEMM3: 1x1 conv onto the shortcut columns (A = A2 from tensor memory) ----
                        if (!mbar_test(bar(s, E2), ph)) continue;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        stamp(it, 3);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint32_t sw = sbase + L::W3 + k * 32;
                            const uint64_t w_hi = umma_desc_at(d64, sw), w_lo = umma_desc_at(d64, sw + 64 * 64);
                            const uint32_t h_hi = T + C_P + 16 * k, h_lo = h_hi + 8;
                            umma_f16_ts(T + C_SC, h_hi, w_hi, i64, 1);
                            umma_f16_ts(T + C_SC, h_hi, w_lo, i64, 1);
                            umma_f16_ts(T + C_SC, h_lo, w_hi, i64, 1);
                        }
                        umma_commit(bar(s, G3));
                        stage[s] = 0;
                        it_s[s] = it + 2;
                        progress = true;
                    }
                }
                if (progress) idle = 0;
                else if (++idle > SPIN_LIMIT) asm volatile("trap;");
            }
        }
    } else {
        // ===================== epilogue warps: group g (8 warps) serves tile slot g =====================
        const int q = warp & 3;                   // TMEM lane quarter (hardware: warp id % 4)
        const int g = (warp - 2) >> 3;            // tile slot
        const int hh = ((warp - 2) >> 2) & 1;     // column half of the quarter's rows
        const int Pin = a.map.Pin, T1 = a.T1;
        const uint32_t T = tmem_base + ((uint32_t)(q * 32) << 16) + 256u * g;
        for (int it = g; it < n_local; it += 2) {
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int m = tile * F_TILE + q * F_QROWS - 1 + lane;
            const int bq = m >= 0 ? m / Pin : 0;
            const int t = m - bq * Pin;
            const bool out_ok = lane >= 1 && lane <= F_QROWS && m >= 0 && m < a.Mtot && t < T1;
            // ---- epilogue 1: ELU(x1) -> E, in place over this warp's 32 x1 columns ----
            mbar_wait(bar(g, G1), ph);
            __syncwarp();  // tcgen05.ld / st are .sync.aligned: re-converge after the predicated stores of the last tile
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if ((threadIdx.x & 255) == 64) stamp(it, 4);
            {
                uint32_t r0[16], r1[16];
                tmem_ld16_nowait(T + C_X1 + 32 * hh, r0);
                tmem_ld16_nowait(T + C_X1 + 32 * hh + 16, r1);
                tmem_wait_ld(r0);
                pin16(r1);
                const float* bd = sbias + 32 * hh;
                {
                    uint32_t e[16];  // [hi 8 | lo 8] of k-step 2 hh
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        split2(elu1(__uint_as_float(r0[2 * i]) + bd[2 * i]), elu1(__uint_as_float(r0[2 * i + 1]) + bd[2 * i + 1]), e[i], e[8 + i]);
                    tmem_st16(T + C_X1 + 32 * hh, e);
                }
                {
                    uint32_t e[16];  // k-step 2 hh + 1
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        split2(elu1(__uint_as_float(r1[2 * i]) + bd[16 + 2 * i]), elu1(__uint_as_float(r1[2 * i + 1]) + bd[16 + 2 * i + 1]), e[i], e[8 + i]);
                    tmem_st16(T + C_X1 + 32 * hh + 16, e);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(g, E1));
            // ---- epilogue 2: h1 = P0[r-1] + P1[r] + P2[r+1] -> ELU -> A2, in place over this warp's tap-0 columns ----
            mbar_wait(bar(g, G2), ph);
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if ((threadIdx.x & 255) == 64) stamp(it, 5);
            {
                // reflect padding of the k3 conv at the clip ends (conv.py:200-210): position -1 reads 1, T1 reads T1 - 2
                const bool first = t == 0, last = t == T1 - 1;
                const bool any_first = __any_sync(0xffffffffu, first), any_last = __any_sync(0xffffffffu, last);
                const float* b1 = sbias + 64 + 16 * hh;
                uint32_t h2[16];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t p0[8], p1[8], p2[8];
                    tmem_ld8_nowait(T + C_P + 16 * hh + 8 * half, p0);
                    tmem_ld8_nowait(T + C_P + 32 + 16 * hh + 8 * half, p1);
                    tmem_ld8_nowait(T + C_P + 64 + 16 * hh + 8 * half, p2);
                    tmem_wait_ld8(p0);
                    pin8(p1);
                    pin8(p2);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float v[2];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float c0 = __uint_as_float(p0[2 * i + j]), c2 = __uint_as_float(p2[2 * i + j]);
                            float lo_tap = __shfl_up_sync(0xffffffffu, c0, 1);    // P0 of row r - 1
                            float hi_tap = __shfl_down_sync(0xffffffffu, c2, 1);  // P2 of row r + 1
                            if (any_first) { const float d = __shfl_down_sync(0xffffffffu, c0, 1); if (first) lo_tap = d; }
                            if (any_last) { const float u = __shfl_up_sync(0xffffffffu, c2, 1); if (last) hi_tap = u; }
                            v[j] = elu1(lo_tap + __uint_as_float(p1[2 * i + j]) + hi_tap + b1[8 * half + 2 * i + j]);
                        }
                        split2(v[0], v[1], h2[4 * half + i], h2[8 + 4 * half + i]);
                    }
                }
                tmem_st16(T + C_P + 16 * hh, h2);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(g, E2));
            // ---- epilogue 3: ELU(y1) planes -> HBM ----
            mbar_wait(bar(g, G3), ph);
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if ((threadIdx.x & 255) == 64) stamp(it, 6);
            {
                uint32_t r0[16], r1[16];
                tmem_ld16_nowait(T + C_SC + 32 * hh, r0);
                tmem_ld16_nowait(T + C_SC + 32 * hh + 16, r1);
                tmem_wait_ld(r0);
                pin16(r1);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(g, E3));  // the accumulators are in registers: the slot may be refilled
                if (out_ok) {
                    const long long sb = a.map.sb ? a.map.sb : a.map.Pout, st = a.map.st ? a.map.st : 1;
                    const long long base = (long long)bq * sb + (long long)a.map.off * st;
                    long long rows[3] = {base + t * st, -1, -1};
                    if (t >= 1 && t <= a.map.hl) rows[1] = base - t * st;
                    if (t <= T1 - 2 && t >= T1 - 1 - a.map.hr) rows[2] = base + (2 * (T1 - 1) - t) * st;
                    const float* b2 = sbias + 96 + 32 * hh;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(half ? r1[i] : r0[i]) + b2[16 * half + i];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if (rows[k] < 0) continue;
                            const long long off = rows[k] * 64 + 32 * hh + 16 * half;
                            if (a.y_f32) {
                                float* o = a.y_f32 + off;
#pragma unroll
                                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                            }
                            store_planes<16, true>(a.ye_hi, a.ye_lo, off, v);
                        }
                    }
                }
            }
            if ((threadIdx.x & 255) == 64) stamp(it, 7);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

bool enc_l1_fused_supported(int cin, int stride) {
    static const bool off = [] { const char* e = std::getenv("WT_ENC_L1_FUSED"); return e && std::atoi(e) == 0; }();
    return !off && cin == 32 && (stride == 2 || stride == 4);
}

namespace {
template <int KB0>
void launch_l1(const EncL1Weights& w, const EncL1Args& a, const FArgs& f, cudaStream_t s) {
    using L = FL<KB0>;
    constexpr int K0 = 64 * KB0;
    static PerDevice<bool> attr_dev;
    bool& attr = attr_dev.get();
    if (!attr) {
        WT_CUDA(cudaFuncSetAttribute(enc_l1_fused_kernel<KB0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
        attr = true;
    }
    // window m of the ELU(y0) planes = elements [m * K0 / 2, m * K0 / 2 + K0): k = 2 * stride positions of 32 channels
    const long long rowsA = a.y0_elems >= K0 ? (a.y0_elems - K0) / (K0 / 2) + 1 : 0;
    const CUtensorMap mA_hi = tc_make_map(a.y0_hi, rowsA, K0, K0 / 2, 32, 64);
    const CUtensorMap mA_lo = tc_make_map(a.y0_lo, rowsA, K0, K0 / 2, 32, 64);
    const CUtensorMap mW1 = tc_make_map(w.w1, 256, K0, K0, 256, 64);
    const CUtensorMap mW2 = tc_make_map(w.w2, 192, 64, 64, 192, 64);
    const CUtensorMap mW3 = tc_make_map(w.w3, 128, 32, 32, 128, 32);
    const int grid = std::min(f.n_tiles, tc_num_sms());
    enc_l1_fused_kernel<KB0><<<grid, F_THREADS, L::SMEM, s>>>(mA_hi, mA_lo, mW1, mW2, mW3, f);
    WT_CUDA(cudaGetLastError());
}
}  // namespace

void launch_enc_l1_fused(const EncL1Weights& w, const EncL1Args& a, cudaStream_t s) {
    if (a.Bc <= 0 || a.T1 < 2) return;
    const int Pin = a.T1 + 2;
    if (a.map.Pin != Pin || a.map.Tvalid != a.T1) throw Error(4, "enc_l1_fused: row map must describe T1 + 2 slots per clip");
    const long long Mtot = (long long)a.Bc * Pin;
    if (Mtot > (1LL << 30)) throw Error(4, "enc_l1_fused: chunk too large");
    FArgs f;
    f.bias = w.bias; f.Mtot = (int)Mtot; f.n_tiles = (int)((Mtot + F_TILE - 1) / F_TILE); f.T1 = a.T1;
    f.map = a.map; f.ye_hi = a.ye_hi; f.ye_lo = a.ye_lo; f.y_f32 = a.y_f32;
    f.dbg = tc_debug_timeline() ? tc_debug_timeline() + 148 * 64 + 64 : nullptr;
    if (w.k0 == 128) launch_l1<2>(w, a, f, s);
    else if (w.k0 == 256) launch_l1<4>(w, a, f, s);
    else throw Error(4, "enc_l1_fused: window of 128 or 256 elements expected");
}

}  // namespace wt
