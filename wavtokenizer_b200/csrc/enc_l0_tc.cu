// Encoder level 0 (conv0 + ResBlock 0) with the two channel-mixing products on the tensor cores (reference
// encoder/modules/seanet.py:107-110 first conv, :45-63 ResBlock; encoder/modules/conv.py:195-211 padding rule):
//
//   x0 = conv_k7(wav) + b0            1 -> 32, reflect padding 3 / 3        CUDA cores (K = 7), weights from the constant bank
//   h1 = conv_k3(ELU(x0)) + b1        32 -> 16, reflect padding 1 / 1       tcgen05: ELU(x0) [128 x 32] (TMEM) x taps as COLUMN
//                                                                           blocks -> P[r, tap*16 + c]; h1[r] = P0[r-1] + P1[r] + P2[r+1]
//   y0 = conv_1x1(ELU(h1)) + shortcut(x0)                                   tcgen05: ELU(h1) [128 x 16] (TMEM) x W (N = 32); the shortcut
//                                     = (W_sc W_0) * wav window + const     composed with conv0 (fp64 at load): a k7 conv of the raw audio
//   out = ELU(y0) as split-fp16 planes in the reflect-padded layout of the level-0 strided conv
//
// resblock0_fused_kernel (encoder_ops.cu) does all of this on the CUDA cores: 2.5 k FMA per sample, FMA-issue bound. Here
// 61 % + 20 % of those FMAs (k3 conv, 1x1 conv) run as 3-pass split-fp16 MMAs whose A operands the compute warps write
// straight into tensor memory (tcgen05.st of packed fp16 pairs, A-from-TMEM MMAs) - no shared-memory tile, no TMA: a thread
// owns one audio position end to end (7 samples in registers serve conv0 AND the composed shortcut).
// Four tiles of 128 rows are in flight per CTA (4 groups of 4 warps, one TMEM slot of 128 columns each); one thread polls
// the four slots and issues whichever product is ready. A tile is four independent groups of 32 rows (one per TMEM lane
// quarter), each with a one-row halo, so 30 of 32 rows produce output and the k3 row shift never leaves a warp.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace wt {

namespace {

constexpr int Z_QROWS = 30;
constexpr int Z_TILE = 4 * Z_QROWS;
constexpr int Z_GROUPS = 4;
constexpr int Z_THREADS = 64 + 16 * 32;

constexpr uint32_t Z_WK = 0;                 // [96 rows x 64 B] (64-byte swizzle): rows [Wk3_hi (tap*16 + n) | Wk3_lo], K = 32
constexpr uint32_t Z_W1 = 96 * 64;           // [64 rows x 32 B] (32-byte swizzle): rows [W1x1_hi | W1x1_lo], K = 16
constexpr uint32_t Z_BAR = Z_W1 + 64 * 32;   // 1 + 4 * 4 mbarriers + TMEM slot
constexpr uint32_t Z_SMEM_USED = Z_BAR + 8 * 18 + 1024;
constexpr uint32_t Z_SMEM = 120 * 1024;      // more than half of the SM: ONE CTA per SM (it allocates all 512 TMEM columns)
static_assert(Z_W1 % 1024 == 0 && Z_SMEM_USED <= Z_SMEM, "layout");

// TMEM columns of a tile slot (four slots of 128 columns)
constexpr uint32_t ZC_E = 0;    // [0, 32)   ELU(x0): k-step k = [hi 8 | lo 8] at 16 k
constexpr uint32_t ZC_P = 32;   // [32, 80)  P: tap blocks of 16
constexpr uint32_t ZC_A2 = 80;  // [80, 96)  ELU(h1): [hi 8 | lo 8]
constexpr uint32_t ZC_Y = 96;   // [96, 128) y0 accumulators

struct L0Const {  // lives in the kernel-parameter constant bank: the FMAs of conv0 / shortcut take it as an operand
    float w0[7 * 32];   // [tap][channel]
    float b0[32];
    float wsc[7 * 32];  // composed shortcut [tap][channel]
    float b1[16];
    float b2[32];
};

struct ZArgs {
    const float* wav;
    int Bc, T, Mtot, n_tiles;
    RowMap map;
    __half* ye_hi;
    __half* ye_lo;
    float* y_f32;
};


template <bool PK>  // PK: conv0 and the composed shortcut as packed fp32 pairs (FFMA2); false: one FFMA per channel (WT_ENC_L0_FFMA2=0)
__global__ void __launch_bounds__(Z_THREADS, 1)
enc_l0_tc_kernel(const __grid_constant__ CUtensorMap mapWk, const __grid_constant__ CUtensorMap mapW1,
                 const __grid_constant__ L0Const K, const ZArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
    // barriers: B_W, then per slot g: 1 + 4 g + {E_READY, P_READY, A2_READY, Y_READY}
    enum { E_READY = 0, P_READY, A2_READY, Y_READY };
    auto bar = [&](int g, int i) { return sbase + Z_BAR + 8u * (1 + 4 * g + i); };
    const uint32_t bar_w = sbase + Z_BAR;
    const uint32_t tmem_slot = sbase + Z_BAR + 8u * 17;
    const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(sptr + Z_BAR + 8u * 17);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_local = a.n_tiles > (int)blockIdx.x ? (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0 && lane == 0) {
        mbar_init(bar_w, 1);
        for (int g = 0; g < Z_GROUPS; ++g) {
            mbar_init(bar(g, E_READY), 4);
            mbar_init(bar(g, P_READY), 1);
            mbar_init(bar(g, A2_READY), 4);
            mbar_init(bar(g, Y_READY), 1);
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
        if (elect_one()) {  // the two weight tiles, once
            mbar_expect_tx(bar_w, 96 * 64 + 64 * 32);
            tma_load_2d(sbase + Z_WK, &mapWk, 0, 0, bar_w);
            tma_load_2d(sbase + Z_W1, &mapW1, 0, 0, bar_w);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread, four tile slots, whichever product is ready =====================
        if (elect_one()) {
            const uint64_t d64 = umma_desc_hi(32), d32 = umma_desc_hi(16);
            constexpr uint32_t i48 = umma_idesc_f16(48), i32 = umma_idesc_f16(32);
            mbar_wait(bar_w, 0);
            int it_g[Z_GROUPS];
            int stage[Z_GROUPS];
#pragma unroll
            for (int g = 0; g < Z_GROUPS; ++g) { it_g[g] = g; stage[g] = 0; }
            uint32_t idle = 0;
            for (;;) {
                bool progress = false, any_left = false;
#pragma unroll
                for (int g = 0; g < Z_GROUPS; ++g) {
                    const int it = it_g[g];
                    if (it >= n_local) continue;
                    any_left = true;
                    const uint32_t ph = (uint32_t)(it >> 2) & 1u;
                    const uint32_t T = tmem_base + 128u * g;
                    if (stage[g] == 0) {
                        // ---- k3 conv: ELU(x0) (TMEM) x [taps as column blocks], K = 32 ----
                        if (!mbar_test(bar(g, E_READY), ph)) continue;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint32_t sw = sbase + Z_WK + k * 32;
                            const uint64_t w_hi = umma_desc_at(d64, sw), w_lo = umma_desc_at(d64, sw + 48 * 64);
                            const uint32_t e_hi = T + ZC_E + 16 * k, e_lo = e_hi + 8;
                            umma_f16_ts(T + ZC_P, e_hi, w_hi, i48, k != 0);
                            umma_f16_ts(T + ZC_P, e_hi, w_lo, i48, 1);
                            umma_f16_ts(T + ZC_P, e_lo, w_hi, i48, 1);
                        }
                        umma_commit(bar(g, P_READY));
                        stage[g] = 1;
                        progress = true;
                    } else {
                        // ---- 1x1 conv: ELU(h1) (TMEM) x W, K = 16 ----
                        if (!mbar_test(bar(g, A2_READY), ph)) continue;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sw = sbase + Z_W1;
                        const uint64_t w_hi = umma_desc_at(d32, sw), w_lo = umma_desc_at(d32, sw + 32 * 32);
                        const uint32_t h_hi = T + ZC_A2, h_lo = h_hi + 8;
                        umma_f16_ts(T + ZC_Y, h_hi, w_hi, i32, 0);
                        umma_f16_ts(T + ZC_Y, h_hi, w_lo, i32, 1);
                        umma_f16_ts(T + ZC_Y, h_lo, w_hi, i32, 1);
                        umma_commit(bar(g, Y_READY));
                        stage[g] = 0;
                        it_g[g] = it + Z_GROUPS;
                        progress = true;
                    }
                }
                if (!any_left) break;
                if (progress) idle = 0;
                else if (++idle > SPIN_LIMIT) asm volatile("trap;");
            }
        }
    } else {
        // ===================== compute warps: group g (4 warps, one per lane quarter) owns tile slot g =====================
        const int q = warp & 3;
        const int g = (warp - 2) >> 2;
        const int Pin = a.map.Pin, T = a.T;
        const uint32_t TM = tmem_base + ((uint32_t)(q * 32) << 16) + 128u * g;
        // the 7 audio samples of a position, reflect-padded at the clip ends (conv.py:79-96); loaded one tile ahead so that the
        // first FMA of step A does not wait for them (5 % of the stall samples sat there, profiles/r02_ncu_stalls_final.txt)
        auto load_samples = [&](int it, float (&sv)[7], int& bq_o, int& t_o, bool& ok_o) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int m = tile * Z_TILE + q * Z_QROWS - 1 + lane;
            int bq = (m >= 0 && it < n_local) ? m / Pin : 0;
            const int t = m - bq * Pin;
            ok_o = lane >= 1 && lane <= Z_QROWS && m >= 0 && m < a.Mtot && t < T;
            if (bq >= a.Bc) bq = a.Bc - 1;  // rows past the last clip: any readable address (their results are dropped)
            bq_o = bq; t_o = t;
            const float* x = a.wav + (long long)bq * T;
            const int tc = (t >= 0 && t < T) ? t : T - 1;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                int idx = tc - 3 + j;
                if (idx < 0) idx = -idx;
                if (idx >= T) idx = 2 * (T - 1) - idx;
                sv[j] = __ldg(x + idx);
            }
        };
        float sn[7];
        int bq_n, t_n;
        bool ok_n;
        load_samples(g, sn, bq_n, t_n, ok_n);
        for (int it = g; it < n_local; it += Z_GROUPS) {
            const uint32_t ph = (uint32_t)(it >> 2) & 1u;
            float s[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) s[j] = sn[j];
            const int bq = bq_n, t = t_n;
            const bool out_ok = ok_n;
            __syncwarp();
            // ---- step A: ELU(conv0) -> packed planes in tensor memory ----
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                uint32_t e[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = 16 * k + 2 * i;
                    // two channels per instruction (FFMA2: the weight pair comes from a uniform-register pair of the constant
                    // bank, the sample is the broadcast scalar operand): 224 instead of 448 FMA issue slots per position
                    float2 v = make_float2(K.b0[c], K.b0[c + 1]);
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        if (PK) {
                            v = __ffma2_rn(make_float2(K.w0[j * 32 + c], K.w0[j * 32 + c + 1]), make_float2(s[j], s[j]), v);
                        } else {
                            v.x = fmaf(K.w0[j * 32 + c], s[j], v.x);
                            v.y = fmaf(K.w0[j * 32 + c + 1], s[j], v.y);
                        }
                    }
                    split2(elu1(v.x), elu1(v.y), e[i], e[8 + i]);
                }
                tmem_st16(TM + ZC_E + 16 * k, e);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(g, E_READY));
            load_samples(it + Z_GROUPS, sn, bq_n, t_n, ok_n);  // next tile's samples: in flight behind steps B and C
            // ---- step B: h1 = P0[r-1] + P1[r] + P2[r+1] -> ELU -> packed planes ----
            mbar_wait(bar(g, P_READY), ph);
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                uint32_t p0[16], p1[16], p2[16];
                tmem_ld16_nowait(TM + ZC_P, p0);
                tmem_ld16_nowait(TM + ZC_P + 16, p1);
                tmem_ld16_nowait(TM + ZC_P + 32, p2);
                tmem_wait_ld(p0);
                pin16(p1);
                pin16(p2);
                const bool first = t == 0, last = t == T - 1;
                const bool any_first = __any_sync(0xffffffffu, first), any_last = __any_sync(0xffffffffu, last);
                uint32_t h2[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float v[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float c0 = __uint_as_float(p0[2 * i + j]), c2 = __uint_as_float(p2[2 * i + j]);
                        float lo_tap = __shfl_up_sync(0xffffffffu, c0, 1);    // P0 of row r - 1
                        float hi_tap = __shfl_down_sync(0xffffffffu, c2, 1);  // P2 of row r + 1
                        if (any_first) { const float d = __shfl_down_sync(0xffffffffu, c0, 1); if (first) lo_tap = d; }
                        if (any_last) { const float u = __shfl_up_sync(0xffffffffu, c2, 1); if (last) hi_tap = u; }
                        v[j] = elu1(lo_tap + __uint_as_float(p1[2 * i + j]) + hi_tap + K.b1[2 * i + j]);
                    }
                    split2(v[0], v[1], h2[i], h2[8 + i]);
                }
                tmem_st16(TM + ZC_A2, h2);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(g, A2_READY));
            // ---- step C: y0 = 1x1 product + composed shortcut -> ELU -> planes in HBM ----
            mbar_wait(bar(g, Y_READY), ph);
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                uint32_t r0[16], r1[16];
                tmem_ld16_nowait(TM + ZC_Y, r0);
                tmem_ld16_nowait(TM + ZC_Y + 16, r1);
                tmem_wait_ld(r0);
                pin16(r1);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (out_ok) {
                    const long long sb = a.map.sb ? a.map.sb : a.map.Pout, st = a.map.st ? a.map.st : 1;
                    const long long base = (long long)bq * sb + (long long)a.map.off * st;
                    long long rows[3] = {base + t * st, -1, -1};
                    if (t >= 1 && t <= a.map.hl) rows[1] = base - t * st;
                    if (t <= T - 2 && t >= T - 1 - a.map.hr) rows[2] = base + (2 * (T - 1) - t) * st;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const int c = 16 * half + i;
                            float2 acc = make_float2(__uint_as_float(half ? r1[i] : r0[i]) + K.b2[c],
                                                     __uint_as_float(half ? r1[i + 1] : r0[i + 1]) + K.b2[c + 1]);
#pragma unroll
                            for (int j = 0; j < 7; ++j) {
                                if (PK) {
                                    acc = __ffma2_rn(make_float2(K.wsc[j * 32 + c], K.wsc[j * 32 + c + 1]), make_float2(s[j], s[j]), acc);
                                } else {
                                    acc.x = fmaf(K.wsc[j * 32 + c], s[j], acc.x);
                                    acc.y = fmaf(K.wsc[j * 32 + c + 1], s[j], acc.y);
                                }
                            }
                            v[i] = acc.x; v[i + 1] = acc.y;
                        }
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if (rows[k] < 0) continue;
                            const long long off = rows[k] * 32 + 16 * half;
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
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

bool enc_l0_tc_supported() {
    static const bool off = [] { const char* e = std::getenv("WT_ENC_L0_FUSED_TC"); return e && std::atoi(e) == 0; }();
    return !off;
}

void launch_enc_l0_tc(const EncL0Weights& w, const float* wav, __half* ye_hi, __half* ye_lo, float* y_f32, int B, int T,
                      const RowMap& map, cudaStream_t s) {
    if (B <= 0) return;
    if (T < 8) throw Error(4, "enc_l0_tc: clip too short");
    static PerDevice<bool> attr_dev;
    bool& attr = attr_dev.get();
    if (!attr) {
        WT_CUDA(cudaFuncSetAttribute(enc_l0_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Z_SMEM));
        WT_CUDA(cudaFuncSetAttribute(enc_l0_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Z_SMEM));
        attr = true;
    }
    static const bool packed = [] { const char* e = std::getenv("WT_ENC_L0_FFMA2"); return !e || std::atoi(e) != 0; }();
    if (map.Pin != T + 2 || map.Tvalid != T) throw Error(4, "enc_l0_tc: row map must describe T + 2 slots per clip");
    const long long Mtot = (long long)B * (T + 2);
    if (Mtot > (1LL << 30)) throw Error(4, "enc_l0_tc: chunk too large");
    static_assert(sizeof(L0Const) == 528 * sizeof(float), "constant block");
    L0Const K;
    std::memcpy(&K, w.consts, sizeof(K));
    const CUtensorMap mWk = tc_make_map(w.wk3, 96, 32, 32, 96, 32);
    const CUtensorMap mW1 = tc_make_map(w.w1x1, 64, 16, 16, 64, 16);
    ZArgs z;
    z.wav = wav; z.Bc = B; z.T = T; z.Mtot = (int)Mtot; z.n_tiles = (int)((Mtot + Z_TILE - 1) / Z_TILE);
    z.map = map; z.ye_hi = ye_hi; z.ye_lo = ye_lo; z.y_f32 = y_f32;
    const int grid = std::min(z.n_tiles, tc_num_sms());
    if (packed) enc_l0_tc_kernel<true><<<grid, Z_THREADS, Z_SMEM, s>>>(mWk, mW1, K, z);
    else enc_l0_tc_kernel<false><<<grid, Z_THREADS, Z_SMEM, s>>>(mWk, mW1, K, z);
    WT_CUDA(cudaGetLastError());
}

}  // namespace wt
