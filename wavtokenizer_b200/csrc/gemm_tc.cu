// tcgen05 GEMM for sm_100a: the tensor-core contraction behind every Conv1d / Linear of the path
// (reference encoder/modules/conv.py:195-211, encoder/modules/seanet.py:45-63,105-141,
// encoder/modules/lstm.py:20, decoder/models.py:58-127,177, decoder/modules.py:43-60, decoder/heads.py:53).
//
//   out[m, n] = epi( sum_seg sum_kb A_seg[row(m, kb), cols(kb)] . W[n, kb*64 : kb*64+64] )
//
// * Operands are fp16 "split" planes: x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi) (~22 mantissa
//   bits). PASSES = 3 issues hi*hi + hi*lo + lo*hi per k-step (fp32 accumulate in TMEM), which keeps the
//   path inside the parity bars of BASELINE.json; PASSES = 1 uses the hi planes only.
// * A operands are TMA tensor maps over channels-last activations. Two addressing modes (gemm_tc.cuh):
//   "tap" mode (a k-tap stride-1 conv is k row-shifted loads of the same 2-D tensor; rows outside the
//   tensor are zero-filled by TMA) and "window" mode (rows of the tensor map OVERLAP: row stride =
//   conv_stride*C, row length = k*C, i.e. the im2col matrix of a strided Conv1d without materialising it).
//   Up to two segments accumulate into one tile (ResBlock: conv1x1(ELU(h)) + shortcut1x1(x)).
// * One CTA per SM, persistent over output tiles (128 x BN). Warp 0 = TMA producer (one elected thread), warp 1 = MMA
//   issuer (one elected thread, tcgen05.mma kind::f16) + TMEM allocator, warps 2..17 = sixteen epilogue warps, four per
//   TMEM lane quarter (tcgen05.ld 32x32b -> bias / GELU / layer-scale / residual / ELU / LSTM cell / VQ argmin -> fp32
//   and/or split-fp16 stores, optionally re-mapped into the reflect-padded layout of the consumer). Wide accumulators
//   (256 columns): two TMEM stages, all 16 warps drain one tile while the mainloop of the next runs; narrow accumulators
//   (<= 128 columns): four TMEM stages and four independent groups of 4 warps (struct Cfg).
// * 256-column tiles run as CTA PAIRS (cluster of 2, two consecutive row tiles of one column tile): by default one
//   tcgen05.mma.cta_group::2 of M = 256 per step with each CTA holding half of the weight rows (CLM = 2); CLM = 1 is the
//   earlier variant (two cta_group::1 MMAs, weight halves multicast into both CTAs).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

#ifndef WT_TIMELINE
#define WT_TIMELINE 0
#endif

namespace wt {

namespace {

template <int CW>
__device__ __forceinline__ void store_row(const TcGemm& g, long long row, int nb, const float (&v)[CW], bool full) {
    if (full) {
        if (g.out_f32) {
            float* o = g.out_f32 + row * g.ldo + nb;
            if (CW % 8 == 0 && ((row * g.ldo + nb) & 7) == 0) {
#pragma unroll
                for (int i = 0; i < CW; i += 8) {
                    uint32_t w[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) w[k] = __float_as_uint(v[i + k]);
                    st256(o + i, w);
                }
            } else {
#pragma unroll
                for (int i = 0; i < CW; i += 4)
                    *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
        if (g.out_hi) {
            if (g.plane_shift) {
                float w[CW];
#pragma unroll
                for (int i = 0; i < CW; ++i) w[i] = v[i] - g.plane_shift[nb + i];
                store_planes<CW, false>(g.out_hi, g.out_lo, row * g.ldh + nb, w);
            } else {
                store_planes<CW, false>(g.out_hi, g.out_lo, row * g.ldh + nb, v);
            }
        }
        if (g.elu_hi) store_planes<CW, true>(g.elu_hi, g.elu_lo, row * g.ldh2 + nb, v);
    } else {
#pragma unroll 1
        for (int i = 0; i < CW; ++i) {
            const int n = nb + i;
            if (n >= g.N) break;
            float x = 0.f;
#pragma unroll
            for (int k = 0; k < CW; ++k) x = (k == i) ? v[k] : x;  // keeps v[] in registers
            if (g.out_f32) g.out_f32[row * g.ldo + n] = x;
            if (g.out_hi) {
                if (g.plane_shift) x -= g.plane_shift[n];
                __half h = __float2half_rn(x);
                g.out_hi[row * g.ldh + n] = h;
                if (g.out_lo) g.out_lo[row * g.ldh + n] = __float2half_rn(x - __half2float(h));
            }
            if (g.elu_hi) {
                float e = elu1(x);
                __half h = __float2half_rn(e);
                g.elu_hi[row * g.ldh2 + n] = h;
                if (g.elu_lo) g.elu_lo[row * g.ldh2 + n] = __float2half_rn(e - __half2float(h));
            }
        }
    }
}

template <int BN, int PASSES, bool CG2 = false>
struct Cfg {
    static constexpr int A_PLANE = BM * BK * 2;  // bytes
    static constexpr int B_PLANE = (CG2 ? BN / 2 : BN) * BK * 2;  // CTA-pair MMA: this CTA holds half of the W rows
    static constexpr int PLANES = PASSES == 3 ? 2 : 1;
    static constexpr int STAGE = PLANES * (A_PLANE + B_PLANE);
    static constexpr int STAGES = (200 * 1024) / STAGE > 8 ? 8 : (200 * 1024) / STAGE;
    static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int BIAS_SMEM = 12 * 1024;  // GELU-only epilogue: the whole bias vector (N <= 3072) staged once per CTA
    // An SS-mode MMA (M = 128, K = 16) costs >= ~128 cycles whatever N is (A-tile fetch from shared memory), so
    // for N <= 128 the hi and lo weight tiles (adjacent in shared memory) are multiplied by A_hi in ONE MMA of
    // N = 2*BN: hh lands in accumulator columns [0, BN), hl in [BN, 2*BN); A_lo x W_hi then adds lh to [0, BN).
    // Two MMAs per k-step instead of three; the epilogue adds the two column blocks.
    static constexpr bool FUSE = PASSES == 3 && BN <= 128;
    static constexpr int ACC_COLS = FUSE ? 2 * BN : BN;          // TMEM columns per accumulator stage
    // Epilogue organisation. Wide accumulators (256 columns): two TMEM stages, all 16 epilogue warps drain one tile
    // together (4 warps per TMEM lane quarter share its columns). Narrow accumulators (<= 128 columns): FOUR TMEM
    // stages and four independent epilogue groups of 4 warps (one per lane quarter); group i owns stage i and
    // handles tiles i, i+4, ... of this CTA, so four tiles are in flight in the epilogue and the per-tile fixed
    // cost (barrier wait, tile decode, row re-map) is paid by 4 warps instead of 16 - the narrow encoder GEMMs were
    // bound by exactly that instruction overhead (profiles/r01_enc_chunk_s3_summary.md).
    static constexpr int NGRP = ACC_COLS <= 128 ? 4 : 1;
    static constexpr int NACC = NGRP == 4 ? 4 : 2;
    static constexpr int TMEM_COLS = NACC * ACC_COLS < 32 ? 32 : NACC * ACC_COLS;  // power of two >= 32
    static constexpr int G = NEPI / 4 / NGRP;                    // epilogue warps per TMEM lane quarter per tile
    // epilogue chunk width (columns per tcgen05.ld) of the generic path; the GELU fast path below uses 32
    static constexpr int CW = BN / G >= 16 ? 16 : 8;
    // k-blocks of look-ahead for the L2 prefetch of the streamed (A) operand; the smem ring itself holds STAGES
    static constexpr int PF = STAGES + 2;
};

struct Maps {
    CUtensorMap a[2][2];  // [segment][plane]
    CUtensorMap w[2];     // [plane]; box of BN rows (BN / 2 rows in the 2-CTA cluster variant: each CTA loads one half)
};

// CL2: the kernel runs as clusters of two CTAs that work on two consecutive row tiles of the SAME column tile in
// lockstep. Each CTA loads its own A tiles and one HALF of every W k-block, multicast into both CTAs, so the W bytes
// cross L2 -> SM once per pair: the single-pass 128 x 256 tiles need 94 B/clk/SM of operands otherwise and were bound
// by that (profiles/r01_final_summary.md). Shared-memory stages are recycled when BOTH CTAs' MMAs have read them
// (tcgen05.commit multicast onto both CTAs' empty barriers).
// CLM = 2 (cta_group::2): same pairing of CTAs and tiles, but the pair runs ONE tcgen05.mma of M = 256 per step, issued by
// the leader CTA; each CTA keeps only its half of the W rows (no multicast copy), see the cta_group::2 helpers above.
// EPI selects a specialised epilogue instantiation: 0 generic (+ VQ argmin), 1 LSTM cell, 2 ConvNeXt GEMM-1 only (bias +
// erf-GELU -> fp16 plane). The specialised kernels do not carry the generic path's live state (row re-map, mirror rows,
// three output forms), which the 96-register cap otherwise turns into local-memory traffic inside the tile loop.
template <int BN, int PASSES, int EPI = 0, int CLM = 0>
// 18 warps are allocated as 20 (warp granularity 4): 65536 / (20 * 32) = 102 -> 96 registers per thread.
__global__ void __launch_bounds__(NUM_THREADS, 1)
tap_gemm_tc_kernel(const __grid_constant__ Maps maps, const TcGemm g) {
    constexpr bool LSTM_EPI = EPI == 1, GELU_EPI = EPI == 2;
    constexpr bool CL2 = CLM != 0;  // the kernel runs as clusters of two CTAs
    constexpr bool CG2 = CLM == 2;  // ... whose tensor cores execute one M = 256 MMA together
    using C = Cfg<BN, PASSES, CG2>;
    constexpr int CW = C::CW;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + C::STAGES * C::STAGE;
    // barriers: full[STAGES], empty[STAGES], tmem_full[NACC], tmem_empty[NACC], then the TMEM base address slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + C::NACC + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 2 * C::NACC);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // GELU-only epilogue: bias[0, N) lives in shared memory behind the barriers (the per-chunk global loads of the bias sat
    // on the epilogue's critical path: 18 % of its stall samples, profiles/r02_ncu_stalls.txt)
    const float* sbias = reinterpret_cast<const float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
    if (GELU_EPI) {
        float* sb = const_cast<float*>(sbias);
        for (int i = threadIdx.x; i < g.N; i += NUM_THREADS) sb[i] = g.bias[i];
    }
    // per-CTA timeline stamps (tools/gemm_timeline.py) exist only in builds with -DWT_TIMELINE=1 (WT_TIMELINE=1 in the
    // environment of _native.build): in the shipped kernels they cost registers the 96-register cap does not have
#if WT_TIMELINE
    long long* dbg = g.dbg ? g.dbg + (long long)blockIdx.x * 64 : nullptr;
    const long long t_begin = dbg ? clock64() : 0;
    auto stamp = [&](int slot) { if (dbg && slot < 64) dbg[slot] = clock64() - t_begin; };
#else
    auto stamp = [](int) {};
#endif
    const int m_tiles = (g.M + BM - 1) / BM;
    const int n_tiles = (g.N + BN - 1) / BN;
    const int per_batch = m_tiles * n_tiles;
    const int total_tiles = per_batch * g.batch;
    const int nkb0 = g.seg[0].num_kb;
    const int num_kb = nkb0 + (g.nseg > 1 ? g.seg[1].num_kb : 0);
    const int PF_DIST = g.prefetch ? C::PF : 0;
    const bool simple_tiles = !CL2 && total_tiles == m_tiles;  // one column tile, one batch: tile index = row tile
    // work distribution: CTA (or CTA pair) `wid` of `nworkers` takes work items wid, wid + nworkers, ...; an item is
    // one output tile, or in the cluster variant a pair of row tiles (2 mp, 2 mp + 1) x one column tile
    const uint32_t cta_rank = CL2 ? cluster_ctarank() : 0u;
    const int wid = CL2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nworkers = CL2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int total_items = CL2 ? ((m_tiles + 1) / 2) * n_tiles : total_tiles;
    auto decode = [&](int item, int& bz, int& mt, int& nt) {
        if (CL2) {
            const int mp = item / n_tiles;
            bz = 0; nt = item - mp * n_tiles; mt = 2 * mp + (int)cta_rank;
        } else if (simple_tiles) {
            bz = 0; mt = item; nt = 0;
        } else {
            bz = item / per_batch;
            const int rem = item - bz * per_batch;
            mt = rem / n_tiles; nt = rem - mt * n_tiles;
        }
    };
    // k-block width: 64 elements (128-byte rows, SWIZZLE_128B) or, for narrow operands, 32 / 16 (64- / 32-byte rows
    // under SWIZZLE_64B / _32B) so that TMA fetches exactly the bytes that exist instead of zero-filling 128-byte rows
    const int kw = g.kw;
    const uint32_t a_plane = (uint32_t)(BM * kw * 2), b_plane = (uint32_t)((CG2 ? BN / 2 : BN) * kw * 2);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            // multicast variant: both CTAs' MMAs must have read the stage; pair MMA: the leader's commit covers both
            mbar_init(empty_bar(s), (CL2 && !CG2) ? 2 : 1);
        }
        for (int s = 0; s < C::NACC; ++s) {
            mbar_init(tfull_bar(s), 1);
            // one arrive per epilogue warp of the group that drains it (pair MMA: of BOTH CTAs, on the leader's barrier)
            mbar_init(tempty_bar(s), (CG2 ? 2 : 1) * (NEPI / C::NGRP));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG2) {  // one warp of EACH CTA of the pair: the columns are allocated in both tensor memories
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"((uint32_t)C::TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"((uint32_t)C::TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL2) cluster_sync_all();  // the peer's barriers exist before anything is multicast at them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (threadIdx.x == 0) stamp(0);  // prologue done

    if (warp == 0) {
        // ===================== TMA producer: ONE elected thread runs the whole persistent loop =====================
        // (elect.sync lets ptxas treat the region as single-threaded: no per-instruction ELECT/R2UR waterfall around
        // UTMALDG, and the mbarrier is polled by one thread; the other 31 lanes wait at the teardown barrier)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = wid; tile < total_items; tile += nworkers) {
                int bz, mt, nt;
                decode(tile, bz, mt, nt);
                const int m0 = mt * BM + (int)(bz * g.a_brows), n0 = nt * BN + (int)(bz * g.w_brows);
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int si = kb < nkb0 ? 0 : 1;
                    const int kl = si ? kb - nkb0 : kb;
                    const int kpt = g.seg[si].kb_per_tap;
                    const int tap = kl / kpt;
                    const int c0 = (kl - tap * kpt) * kw;
                    const int r0 = m0 + g.seg[si].shift0 + tap;
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * C::STAGE;
                    const uint32_t sb = sa + C::PLANES * a_plane;
                    if (CG2) {
                        // pair MMA: both CTAs' loads complete on the LEADER's full barrier (its MMA thread is the only
                        // consumer); the leader arms it with the bytes of both CTAs. Each CTA keeps its own A rows and
                        // its own half of the W rows: nothing is multicast.
                        const uint32_t fb = mapa_rank(full_bar(stage), 0);
                        if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::PLANES * (a_plane + b_plane));
                        const int nh = n0 + (int)cta_rank * (BN / 2);
                        tma_load_2d_cg2(sa, &maps.a[si][0], c0, r0, fb);
                        if (PASSES == 3) tma_load_2d_cg2(sa + a_plane, &maps.a[si][1], c0, r0, fb);
                        tma_load_2d_cg2(sb, &maps.w[0], kb * kw, nh, fb);
                        if (PASSES == 3) tma_load_2d_cg2(sb + b_plane, &maps.w[1], kb * kw, nh, fb);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(full_bar(stage), C::PLANES * (a_plane + b_plane));
                    tma_load_2d(sa, &maps.a[si][0], c0, r0, full_bar(stage));
                    if (PASSES == 3) tma_load_2d(sa + a_plane, &maps.a[si][1], c0, r0, full_bar(stage));
                    if (CL2) {  // this CTA's half of the W rows, delivered to both CTAs of the pair
                        const uint32_t hoff = cta_rank * (b_plane / 2);
                        const int nh = n0 + (int)cta_rank * (BN / 2);
                        tma_load_2d_mc(sb + hoff, &maps.w[0], kb * kw, nh, full_bar(stage), (uint16_t)3);
                        if (PASSES == 3) tma_load_2d_mc(sb + b_plane + hoff, &maps.w[1], kb * kw, nh, full_bar(stage), (uint16_t)3);
                    } else {
                        tma_load_2d(sb, &maps.w[0], kb * kw, n0, full_bar(stage));
                        if (PASSES == 3) tma_load_2d(sb + b_plane, &maps.w[1], kb * kw, n0, full_bar(stage));
                    }
                    if (PF_DIST > 0 && kb + PF_DIST < num_kb) {  // A operand of a later k-block of this tile -> L2
                        const int kf = kb + PF_DIST;
                        const int sf = kf < nkb0 ? 0 : 1;
                        const int kfl = sf ? kf - nkb0 : kf;
                        const int kptf = g.seg[sf].kb_per_tap;
                        const int tapf = kfl / kptf;
                        tma_prefetch_2d(&maps.a[sf][0], (kfl - tapf * kptf) * kw, m0 + g.seg[sf].shift0 + tapf);
                        if (PASSES == 3) tma_prefetch_2d(&maps.a[sf][1], (kfl - tapf * kptf) * kw, m0 + g.seg[sf].shift0 + tapf);
                    }
                    if (tile == wid) stamp(1 + kb);  // slots 1..16: producer issued k-block kb (first tile)
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: ONE elected thread (pair MMA: of the leader CTA only) =====================
        if ((!CG2 || cta_rank == 0) && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16(BN, CG2 ? 2 * BM : BM);
            constexpr uint32_t idesc2 = umma_idesc_f16(C::FUSE ? 2 * BN : BN);  // A_hi x [W_hi; W_lo]
            // k16 steps that hold real columns in the LAST k-block of a segment row: narrow operands (16 channels,
            // 8 audio taps, k*C = 96) are zero-padded to 64 by TMA, and an SS-mode MMA costs ~110 cycles whatever it
            // multiplies, so the all-zero steps are simply not issued
            const int full16 = kw / UMMA_K;
            int tail0 = full16, tail1 = full16;
            {
                const long long r0 = g.seg[0].inner - (long long)(g.seg[0].kb_per_tap - 1) * kw;
                if (r0 < kw) tail0 = (int)((r0 + UMMA_K - 1) / UMMA_K);
                if (g.nseg > 1) {
                    const long long r1 = g.seg[1].inner - (long long)(g.seg[1].kb_per_tap - 1) * kw;
                    if (r1 < kw) tail1 = (int)((r1 + UMMA_K - 1) / UMMA_K);
                }
            }
            const int kpt0 = g.seg[0].kb_per_tap, kpt1 = g.nseg > 1 ? g.seg[1].kb_per_tap : 1;
            const bool any_tail = tail0 < full16 || tail1 < full16;
            const uint64_t dhi = umma_desc_hi(kw);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int ti = -1;
            for (int tile = wid; tile < total_items; tile += nworkers) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                ++ti;
                if (ti < 5) stamp(44 + 4 * ti);  // MMA may start tile ti (accumulator free)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * C::ACC_COLS);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (tile == wid) stamp(17 + kb);  // slots 17..32: data of k-block kb landed
                    const uint32_t sa = smem_base + stage * C::STAGE;
                    const uint32_t sb = sa + C::PLANES * a_plane;
                    int nk16 = full16;
                    if (any_tail) {
                        if (kb < nkb0) { if ((kb + 1) % kpt0 == 0) nk16 = tail0; }
                        else if ((kb - nkb0 + 1) % kpt1 == 0) nk16 = tail1;
                    }
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (k >= nk16) break;
                        const uint32_t koff = k * UMMA_K * 2;  // bytes inside the swizzle row
                        const uint64_t a_hi = umma_desc_at(dhi, sa + koff);
                        const uint64_t b_hi = umma_desc_at(dhi, sb + koff);
                        if (C::FUSE) {
                            const uint64_t a_lo = umma_desc_at(dhi, sa + a_plane + koff);
                            umma_f16(tmem_d, a_hi, b_hi, idesc2, (kb | k) != 0);  // [hh | hl], W_lo tile follows W_hi
                            umma_f16(tmem_d, a_lo, b_hi, idesc, 1);               // + lh into the first BN columns
                        } else if (CG2) {
                            // M = 256 over the pair: descriptors are this (leader) CTA's shared-memory offsets, the peer's
                            // tensor core uses the same offsets in ITS shared memory
                            umma_f16_cg2(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0);
                            if (PASSES == 3) {
                                const uint64_t a_lo = umma_desc_at(dhi, sa + a_plane + koff);
                                const uint64_t b_lo = umma_desc_at(dhi, sb + b_plane + koff);
                                umma_f16_cg2(tmem_d, a_hi, b_lo, idesc, 1);
                                umma_f16_cg2(tmem_d, a_lo, b_hi, idesc, 1);
                            }
                        } else {
                            umma_f16(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0);
                            if (PASSES == 3) {
                                const uint64_t a_lo = umma_desc_at(dhi, sa + a_plane + koff);
                                const uint64_t b_lo = umma_desc_at(dhi, sb + b_plane + koff);
                                umma_f16(tmem_d, a_hi, b_lo, idesc, 1);
                                umma_f16(tmem_d, a_lo, b_hi, idesc, 1);
                            }
                        }
                    }
                    if (CG2) umma_commit_cg2(empty_bar(stage), (uint16_t)3);  // frees the stage in both CTAs
                    else if (CL2) umma_commit_mc(empty_bar(stage), (uint16_t)3);  // ... in BOTH CTAs: the peer's W half lives here too
                    else umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
                    if (kb == num_kb - 1) {
                        if (CG2) umma_commit_cg2(tfull_bar(acc), (uint16_t)3);  // both CTAs' epilogues drain their 128 rows
                        else
                        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                        if (ti < 5) stamp(45 + 4 * ti);  // all MMAs of tile ti issued
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == C::NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..17) =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp may read (hardware: warp id % 4)
        const int grp = ((warp - 2) >> 2) / C::G;  // epilogue group: handles this CTA's tiles grp, grp + NGRP, ...
        const int cg = ((warp - 2) >> 2) % C::G;   // which share of the tile's columns this warp handles
        for (int ti = grp;; ti += C::NGRP) {
            const int tile = wid + ti * nworkers;
            if (tile >= total_items) break;
            const int acc = ti % C::NACC;
            const uint32_t acc_phase = (uint32_t)(ti / C::NACC) & 1u;
            int bz, mt, nt;
            decode(tile, bz, mt, nt);
            const int m_local = mt * BM + q * 32 + lane;
            const int m = m_local + (int)(bz * g.o_brows);
            const int n0 = nt * BN;
            mbar_wait(tfull_bar(acc), acc_phase);
            if (threadIdx.x == 64 && ti < 5) stamp(46 + 4 * ti);  // accumulator of tile ti ready
            if (tile == wid && threadIdx.x == 64) stamp(40);  // accumulator of the first tile ready
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::ACC_COLS);
            // accumulator chunk = columns [col, col+n) (+ the hl block BN columns further when the passes are fused)
            auto ld_acc8 = [&](uint32_t col, uint32_t (&r)[8]) {
                tmem_ld(tbase + col, r);
                if (C::FUSE) {
                    uint32_t t2[8];
                    tmem_ld(tbase + (uint32_t)BN + col, t2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(t2[i]));
                }
            };
            bool row_ok = m_local < g.M;
            // destination rows (identity, or re-mapped into the consumer's reflect-padded layout)
            long long dst = m, mir_l = -1, mir_r = -1;
            if (!GELU_EPI && g.map.Pin) {
                const int b = m / g.map.Pin, t = m - b * g.map.Pin;
                row_ok = row_ok && t < g.map.Tvalid;
                const long long sb = g.map.sb ? g.map.sb : g.map.Pout, st = g.map.st ? g.map.st : 1;
                const long long base = (long long)b * sb + (long long)g.map.off * st;
                dst = base + t * st;
                if (t >= 1 && t <= g.map.hl) mir_l = base - t * st;
                if (t <= g.map.Tvalid - 2 && t >= g.map.Tvalid - 1 - g.map.hr)
                    mir_r = base + (2 * (g.map.Tvalid - 1) - t) * st;
            }
            if constexpr (LSTM_EPI) {  // separate instantiation: keeps the cell's registers out of the generic kernels
                // LSTM cell (reference encoder/modules/lstm.py:20; gates i, f, g, o). Tile columns hold
                // [i | f | g | o] x 16 hidden units; the thread owns one batch row, two warps per lane quarter work.
#pragma unroll 1
                for (int hs = cg; hs < 2; hs += C::G) {  // two blocks of 8 hidden units per tile
                    const int u0 = nt * 16 + hs * 8;
                    uint32_t ri[8], rf[8], rg[8], ro[8];
                    __syncwarp();
                    ld_acc8(0 + hs * 8, ri);
                    ld_acc8(16 + hs * 8, rf);
                    ld_acc8(32 + hs * 8, rg);
                    ld_acc8(48 + hs * 8, ro);
                    if (row_ok) {
                        const float* xr = g.res + (long long)m * g.ldres + n0 + hs * 8;
                        float* cr = g.cell + (long long)m * g.hidden + u0;
                        float xi[8], xf[8], xg[8], xo[8], cv[8];
                        *reinterpret_cast<float4*>(xi) = *reinterpret_cast<const float4*>(xr);
                        *reinterpret_cast<float4*>(xi + 4) = *reinterpret_cast<const float4*>(xr + 4);
                        *reinterpret_cast<float4*>(xf) = *reinterpret_cast<const float4*>(xr + 16);
                        *reinterpret_cast<float4*>(xf + 4) = *reinterpret_cast<const float4*>(xr + 20);
                        *reinterpret_cast<float4*>(xg) = *reinterpret_cast<const float4*>(xr + 32);
                        *reinterpret_cast<float4*>(xg + 4) = *reinterpret_cast<const float4*>(xr + 36);
                        *reinterpret_cast<float4*>(xo) = *reinterpret_cast<const float4*>(xr + 48);
                        *reinterpret_cast<float4*>(xo + 4) = *reinterpret_cast<const float4*>(xr + 52);
                        *reinterpret_cast<float4*>(cv) = *reinterpret_cast<const float4*>(cr);
                        *reinterpret_cast<float4*>(cv + 4) = *reinterpret_cast<const float4*>(cr + 4);
                        float hv[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float ig = sigmoid1(__uint_as_float(ri[i]) + xi[i]);
                            const float fg = sigmoid1(__uint_as_float(rf[i]) + xf[i]);
                            const float gg = tanhf(__uint_as_float(rg[i]) + xg[i]);
                            const float og = sigmoid1(__uint_as_float(ro[i]) + xo[i]);
                            cv[i] = fg * cv[i] + ig * gg;
                            hv[i] = og * tanhf(cv[i]);
                        }
                        *reinterpret_cast<float4*>(cr) = *reinterpret_cast<float4*>(cv);
                        *reinterpret_cast<float4*>(cr + 4) = *reinterpret_cast<float4*>(cv + 4);
                        if (g.out_f32) {
                            float* o = g.out_f32 + (long long)m * g.ldo + u0;
                            *reinterpret_cast<float4*>(o) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                            *reinterpret_cast<float4*>(o + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                        }
                        if (g.out_hi) store_planes<8, false>(g.out_hi, g.out_lo, (long long)m * g.ldh + u0, hv);
                    }
                }
            } else if (!GELU_EPI && BN == 256 && PASSES == 3 && g.act == TC_ACT_ARGMIN) {
                // nearest code (reference encoder/quantization/core_vq.py:175-183): argmin_n ||x - c_n||^2 =
                // argmin_n (||c_n||^2 - 2 x.c_n); operands are centred on the codebook mean (distance-invariant),
                // first index wins ties (packed (distance, index) keys under a 64-bit atomicMin).
                float bd = INFINITY;
                int bi = 0;
#pragma unroll 1
                for (int c = cg; c < BN / CW; c += C::G) {
                    uint32_t r[CW];
                    __syncwarp();
                    tmem_ld(tbase + (uint32_t)(c * CW), r);
                    const int nb = n0 + c * CW;
#pragma unroll
                    for (int i = 0; i < CW; ++i) {
                        const float d = fmaf(-2.f, __uint_as_float(r[i]), g.bias[nb + i]);
                        if (d < bd) { bd = d; bi = nb + i; }
                    }
                }
                if (row_ok) {
                    uint32_t u = __float_as_uint(bd);
                    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving float -> uint
                    atomicMin(g.best + m, ((unsigned long long)u << 32) | (unsigned)bi);
                }
            } else if (GELU_EPI || (PASSES == 1 && BN == 256 && g.act == TC_ACT_GELU && g.out_hi && !g.out_lo && !g.out_f32 &&
                                    !g.elu_hi && !g.map.Pin && !g.gamma && !g.res && n0 + BN <= g.N)) {
                // ConvNeXt GEMM-1 fast path (reference decoder/modules.py:54-55): bias + exact-erf GELU -> fp16 plane,
                // 32 columns per tcgen05.ld so that the fixed per-chunk cost is paid half as often; nothing else live.
                static_assert(BN / C::G == 64 || BN != 256 || PASSES != 1, "two 32-column chunks per warp");
                if constexpr (GELU_EPI) {
                    // 16-column chunks, the tcgen05.ld of chunk c + 1 in flight while chunk c is computed (a chunk's load
                    // latency under 16 warps reading tensor memory is of the order of its ~250 instructions of math)
                    constexpr int NCH = BN / C::G / 16;
                    const int col0 = cg * (BN / C::G);
                    const float* bp0 = sbias + n0 + col0;
                    __half* op0 = g.out_hi + (long long)m * g.ldh + n0 + col0;
                    uint32_t ra[16], rb[16];
                    __syncwarp();
                    tmem_ld16_nowait(tbase + (uint32_t)col0, ra);
                    tmem_wait_ld(ra);
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        uint32_t(&cur)[16] = (c & 1) ? rb : ra;
                        uint32_t(&nxt)[16] = (c & 1) ? ra : rb;
                        if (c + 1 < NCH) tmem_ld16_nowait(tbase + (uint32_t)(col0 + 16 * (c + 1)), nxt);
                        if (row_ok) {
                            uint32_t w[8];
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 b = *reinterpret_cast<const float4*>(bp0 + 16 * c + i);
                                const float v0 = gelu_erf(__uint_as_float(cur[i]) + b.x);
                                const float v1 = gelu_erf(__uint_as_float(cur[i + 1]) + b.y);
                                const float v2 = gelu_erf(__uint_as_float(cur[i + 2]) + b.z);
                                const float v3 = gelu_erf(__uint_as_float(cur[i + 3]) + b.w);
                                const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
                                w[i / 2] = *reinterpret_cast<const uint32_t*>(&h01);
                                w[i / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                            }
                            st256(op0 + 16 * c, w);
                        }
                        __syncwarp();
                        if (c + 1 < NCH) tmem_wait_ld(nxt);
                    }
                } else {
#pragma unroll 1
                for (int c2 = 0; c2 < BN / C::G / 32; ++c2) {
                    const int col = cg * (BN / C::G) + c2 * 32;
                    uint32_t r[32];
                    __syncwarp();
                    tmem_ld(tbase + (uint32_t)col, r);
                    if (row_ok) {
                        const float* bp = (GELU_EPI ? sbias : g.bias) + n0 + col;
                        __half* op = g.out_hi + (long long)m * g.ldh + n0 + col;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            uint32_t w[8];
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 b = *reinterpret_cast<const float4*>(bp + 16 * j + i);
                                const float v0 = gelu_erf(__uint_as_float(r[16 * j + i]) + b.x);
                                const float v1 = gelu_erf(__uint_as_float(r[16 * j + i + 1]) + b.y);
                                const float v2 = gelu_erf(__uint_as_float(r[16 * j + i + 2]) + b.z);
                                const float v3 = gelu_erf(__uint_as_float(r[16 * j + i + 3]) + b.w);
                                const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
                                w[i / 2] = *reinterpret_cast<const uint32_t*>(&h01);
                                w[i / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                            }
                            st256(op + 16 * j, w);
                        }
                    }
                }
                }
            } else if constexpr (!GELU_EPI) {
#pragma unroll 1
                for (int c = cg; c < BN / CW; c += C::G) {
                    uint32_t r[CW];
                    __syncwarp();  // tcgen05.ld is .sync.aligned: re-converge after the predicated stores below
                    if constexpr (C::FUSE) {
                        uint32_t t2[CW];
                        tmem_ld_pair(tbase + (uint32_t)(c * CW), r, tbase + (uint32_t)(BN + c * CW), t2);
#pragma unroll
                        for (int i = 0; i < CW; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(t2[i]));
                    } else {
                        tmem_ld(tbase + (uint32_t)(c * CW), r);
                    }
                    const int nb = n0 + c * CW;
                    if (!row_ok || nb >= g.N) continue;
                    float v[CW];
#pragma unroll
                    for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]);
                    const bool full = nb + CW <= g.N;
                    if (full) {
                        if (g.bias) {
#pragma unroll
                            for (int i = 0; i < CW; i += 4) {
                                float4 b = *reinterpret_cast<const float4*>(g.bias + nb + i);
                                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                            }
                        }
                        if (g.act == TC_ACT_GELU) {
#pragma unroll
                            for (int i = 0; i < CW; ++i) v[i] = gelu_erf(v[i]);
                        }
                        if (g.gamma) {
#pragma unroll
                            for (int i = 0; i < CW; i += 4) {
                                float4 b = *reinterpret_cast<const float4*>(g.gamma + nb + i);
                                v[i] *= b.x; v[i + 1] *= b.y; v[i + 2] *= b.z; v[i + 3] *= b.w;
                            }
                        }
                        if (g.res) {
                            const float* rr = g.res + (long long)m * g.ldres + nb;
#pragma unroll
                            for (int i = 0; i < CW; i += 4) {
                                float4 b = *reinterpret_cast<const float4*>(rr + i);
                                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < CW; ++i) {
                            const int n = nb + i;
                            if (n < g.N) {
                                if (g.bias) v[i] += g.bias[n];
                                if (g.act == TC_ACT_GELU) v[i] = gelu_erf(v[i]);
                                if (g.gamma) v[i] *= g.gamma[n];
                                if (g.res) v[i] += g.res[(long long)m * g.ldres + n];
                            }
                        }
                    }
                    store_row<CW>(g, dst, nb, v, full);
                    if (mir_l >= 0) store_row<CW>(g, mir_l, nb, v, full);
                    if (mir_r >= 0) store_row<CW>(g, mir_r, nb, v, full);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (threadIdx.x == 64 && ti < 5) stamp(47 + 4 * ti);  // epilogue of tile ti done (this warp)
            if (tile == wid && threadIdx.x == 64) stamp(41);  // epilogue of the first tile done
            if (lane == 0) {
                if (CG2) mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0));  // the leader's MMA thread waits for both CTAs
                else mbar_arrive(tempty_bar(acc));
            }
        }
    }

    if (threadIdx.x == 64) stamp(43);  // this epilogue warp finished its last tile (a reliable "CTA done" time)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL2) cluster_sync_all();  // the peer may still multicast into this CTA's stages / arrive on its barriers
    if (threadIdx.x == 0) stamp(42);  // barrier issued (BAR.SYNC.DEFER_BLOCKING: not the release time)
    if (warp == 1) {
        if (CG2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"((uint32_t)C::TMEM_COLS)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"((uint32_t)C::TMEM_COLS)
                         : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// Persistent LSTM layer (reference encoder/modules/lstm.py:20, nn.LSTM recurrence; gates i, f, g, o).
//
// One cooperative launch runs all L time steps of one layer. CTA (ns, mg) owns the 64 gate columns of
// hidden units [16 ns, 16 ns + 16) (W_hh rows are permuted at load so that they are contiguous) and the batch
// tiles mg, mg + MG, ...: its slice of W_hh (hi + lo planes, 128 KB) is loaded into shared memory ONCE and
// stays resident for the whole layer. Per step it streams h_{t-1} of its batch tile through a 3-stage TMA
// ring, accumulates W_hh h in TMEM with 3-pass split-fp16 tcgen05 MMAs, and the epilogue warps apply the
// LSTM cell (adding the hoisted input projection) and publish h_t as fp32 rows and split-fp16 planes.
// The 32 CTAs that share a batch tile synchronise once per step through an arrival counter in global memory
// (release/acquire + proxy fences because h_t is written by threads and read back by TMA). All tensors are
// time-major: row = t*B + b.
// ---------------------------------------------------------------------------------------
struct LstmArgs {
    const float* xin;    // [L*B, 4D] W_ih x + b, gate-permuted ([i16 | f16 | g16 | o16] per 64 columns)
    float* y;            // [L*B, D] h_t fp32
    __half* h_hi;        // [L*B, D] h_t split planes (also this kernel's A operand through mapH)
    __half* h_lo;
    float* cell;         // [B, D]
    int* counters;       // [m_tiles * L * LSTM_KB * LSTM_CNT_PITCH], zeroed before launch
    int B, L, D, m_tiles;
    int t_begin, t_end;  // this launch runs steps [t_begin, t_end): h, c and the counters of earlier steps are in place
    int publish;         // 0: every lane fences (fence.proxy.async + membar.gl) before the release; 1: one release per warp;
                         // 2: one release per CTA behind a named barrier of the epilogue warps
    int poll_ns;         // back-off between two polls of the arrival counters (0: none)
    int keep_c;          // 1: c_t in registers between steps, y_t stored one step late (0: both stored every step)
    long long* dbg;      // optional timeline of CTA (0,0): 8 stamps per step for steps 4..7
};

constexpr int LSTM_EPI_WARPS = 16;  // 4 per TMEM lane quarter, 4 hidden units each
constexpr int LSTM_THREADS = 64 + LSTM_EPI_WARPS * 32;
constexpr int LSTM_KB = 8;                               // D = 512 = 8 k-blocks of 64
constexpr int LSTM_W_BYTES = LSTM_KB * 2 * 64 * 128;     // resident W slice: 8 kb x (hi, lo) x [64 rows x 128 B]
constexpr int LSTM_A_STAGE = 2 * BM * 128;               // hi + lo tile of h_{t-1}
#ifndef WT_LSTM_STAGES
#define WT_LSTM_STAGES 3  // measured: 2 and 3 ring stages give the same step time; with 2 the third stage's room holds the h staging tile of publish = 3
#endif
constexpr int LSTM_STAGES = WT_LSTM_STAGES;
constexpr int LSTM_HST = LSTM_STAGES <= 2 ? 2 * BM * 32 : 0;  // staging tile of h_t (hi + lo, [128 rows x 16 units]) for the TMA store
constexpr int LSTM_SMEM = LSTM_W_BYTES + LSTM_STAGES * LSTM_A_STAGE + LSTM_HST + 1024 + 256;
static_assert(LSTM_SMEM <= 227 * 1024, "LSTM shared memory budget");
constexpr int LSTM_CTAS_PER_KB = 4 * 16;  // arrivals per k-block of 64 hidden units: four CTAs x 16 epilogue warps
#ifndef WT_LSTM_CNT_PITCH
#define WT_LSTM_CNT_PITCH 8
#endif
constexpr int LSTM_CNT_PITCH = WT_LSTM_CNT_PITCH;     // ints between counters (8: one 32-byte sector each; 32: one line each)

// CL > 1: the kernel runs as clusters of CL CTAs along the gate-slice axis. The CL CTAs of a cluster share one batch tile,
// i.e. they all stream the SAME h_{t-1} tile every step: each loads 128 / CL of its rows per k-block and multicasts them
// into all CL shared memories, so the tile crosses L2 -> SM once per cluster instead of once per CTA (64 CTAs x 256 KB =
// 16 MB per step and layer otherwise, which is what the load phase of a step was bound by, twice over with both layers of
// the wavefront resident). A ring stage is recycled when the MMAs of all CL CTAs have read it (multicast commit).
template <int CL>
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_persistent_kernel(const __grid_constant__ CUtensorMap mapH_hi, const __grid_constant__ CUtensorMap mapH_lo,
                       const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
                       const __grid_constant__ CUtensorMap mapS_hi, const __grid_constant__ CUtensorMap mapS_lo,
                       const LstmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_base = smem_base;
    const uint32_t a_base = smem_base + LSTM_W_BYTES;
    const uint32_t hst_base = a_base + LSTM_STAGES * LSTM_A_STAGE;  // h_t staging tile (publish = 3)
    const uint32_t bar_base = hst_base + LSTM_HST;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (LSTM_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * LSTM_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * LSTM_STAGES + 2 + s); };
    const uint32_t w_bar = bar_base + 8u * (2 * LSTM_STAGES + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * LSTM_STAGES + 5);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ns = blockIdx.x;           // slice of 64 gate columns = 16 hidden units
    const int mg = blockIdx.y, MG = gridDim.y;
    long long* dbg = (a.dbg && blockIdx.x == 0 && blockIdx.y == 0) ? a.dbg : nullptr;
    const long long t_begin = dbg ? clock64() : 0;
    auto stamp = [&](int t, int slot) { if (dbg && t >= 4 && t < 8) dbg[(t - 4) * 8 + slot] = clock64() - t_begin; };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < LSTM_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), LSTM_EPI_WARPS); }
        mbar_init(w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // the peers' barriers exist before anything is multicast at them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr int QROWS = BM / CL;  // rows of a k-block this CTA loads (and multicasts)

    if (warp == 0) {
        // ---- producer: resident W slice once, then h_{t-1} tiles every step ----
        if (elect_one()) {
            mbar_expect_tx(w_bar, LSTM_W_BYTES);
            for (int kb = 0; kb < LSTM_KB; ++kb) {
                tma_load_2d(w_base + (kb * 2 + 0) * 8192, &mapW_hi, kb * BK, ns * 64, w_bar);
                tma_load_2d(w_base + (kb * 2 + 1) * 8192, &mapW_lo, kb * BK, ns * 64, w_bar);
            }
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int t = (a.t_begin > 1 ? a.t_begin : 1); t < a.t_end; ++t) {
            for (int mt = mg; mt < a.m_tiles; mt += MG) {
                // k-block kb of h_{t-1} (hidden units [64 kb, 64 kb + 64)) is produced by the four CTAs ns = 4 kb ..
                // 4 kb + 3 of this batch tile: wait for exactly those (one counter per k-block), so the loads of the
                // early k-blocks overlap the skew of the late publishers instead of starting after the last one
                const int r0 = (t - 1) * a.B + mt * BM;
                const int* cnt0 = a.counters + ((long long)mt * a.L + (t - 1)) * LSTM_KB * LSTM_CNT_PITCH;
                int issued = 0;
                uint32_t spins = 0;
                while (issued < LSTM_KB) {
                    // lanes 0..7 poll the eight k-block counters in parallel (one L2 round trip per poll round)
                    int v = LSTM_CTAS_PER_KB;
                    if (lane < LSTM_KB)
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(cnt0 + lane * LSTM_CNT_PITCH) : "memory");
                    const unsigned ready = __ballot_sync(0xffffffffu, v >= LSTM_CTAS_PER_KB);
                    __syncwarp();  // memory ordering between the polling lanes and the issuing lane
                    int n = issued;
                    while (n < LSTM_KB && ((ready >> n) & 1u)) ++n;
                    if (n == issued) {
                        if (++spins > (1u << 24)) asm volatile("trap;");
                        if (a.poll_ns) __nanosleep(a.poll_ns);  // back-off: the polls share L2 lines with the arrivals they wait for
                        continue;
                    }
                    if (issued == 0 && lane == 0) stamp(t, 0);  // first k-block(s) of h_{t-1} published
                    for (int kb = issued; kb < n; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        if (elect_one()) {
                            if (kb == issued) asm volatile("fence.proxy.async;" ::: "memory");
                            const uint32_t sa = a_base + stage * LSTM_A_STAGE;
                            mbar_expect_tx(full_bar(stage), LSTM_A_STAGE);
                            if (CL > 1) {
                                const uint32_t off = crank * (uint32_t)(QROWS * 128);
                                const int rq = r0 + (int)crank * QROWS;
                                tma_load_2d_mc(sa + off, &mapH_hi, kb * BK, rq, full_bar(stage), (uint16_t)((1u << CL) - 1));
                                tma_load_2d_mc(sa + BM * 128 + off, &mapH_lo, kb * BK, rq, full_bar(stage), (uint16_t)((1u << CL) - 1));
                            } else {
                                tma_load_2d(sa, &mapH_hi, kb * BK, r0, full_bar(stage));
                                tma_load_2d(sa + BM * 128, &mapH_lo, kb * BK, r0, full_bar(stage));
                            }
                            if (kb == LSTM_KB - 1) stamp(t, 1);  // last k-block issued
                        }
                        __syncwarp();
                        if (++stage == LSTM_STAGES) { stage = 0; phase ^= 1; }
                    }
                    issued = n;
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        constexpr uint32_t idesc = umma_idesc_f16(64);
        constexpr uint32_t idesc2 = umma_idesc_f16(128);  // h_hi x [W_hi; W_lo]: the lo tile follows the hi tile
        mbar_wait(w_bar, 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = (a.t_begin > 1 ? a.t_begin : 1); t < a.t_end; ++t) {
            for (int mt = mg; mt < a.m_tiles; mt += MG) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 128);
                for (int kb = 0; kb < LSTM_KB; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        if (kb == 0) stamp(t, 2);               // first k-block landed
                        if (kb == LSTM_KB - 1) stamp(t, 3);     // last k-block landed
                        const uint32_t sa = a_base + stage * LSTM_A_STAGE;
                        const uint32_t sw = w_base + kb * 2 * 8192;
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint32_t koff = k * UMMA_K * 2;
                            const uint64_t a_hi = umma_desc_sw128(sa + koff), a_lo = umma_desc_sw128(sa + BM * 128 + koff);
                            const uint64_t b_hi = umma_desc_sw128(sw + koff);
                            umma_f16(tmem_d, a_hi, b_hi, idesc2, (kb | k) != 0);  // [hh | hl] in one MMA (N = 128)
                            umma_f16(tmem_d, a_lo, b_hi, idesc, 1);               // + lh into columns [0, 64)
                        }
                        if (CL > 1) umma_commit_mc(empty_bar(stage), (uint16_t)((1u << CL) - 1));  // frees the stage in every CTA
                        else umma_commit(empty_bar(stage));
                        if (kb == LSTM_KB - 1) umma_commit(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == LSTM_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ---- epilogue: LSTM cell, 16 warps = 4 TMEM lane quarters x 4 groups of 4 hidden units ----
        const int q = warp & 3;
        const int hw = (warp - 2) >> 2;
        const int u0 = ns * 16 + hw * 4;
        int acc = 0;
        uint32_t acc_phase = 0;
        // One batch tile per CTA (the usual case): c_t stays in registers between the steps of a launch (it is read back only
        // by the thread that wrote it). y_t is stored one step late, after the next accumulator wait: the release that
        // publishes h_t waits for the outstanding writes of the SM, and only the h planes belong in front of it.
        const bool c_in_regs = a.keep_c && a.m_tiles <= MG;
        float4 c_keep = make_float4(0.f, 0.f, 0.f, 0.f), y_pend = c_keep;
        long long y_pend_row = -1;
        for (int t = a.t_begin; t < a.t_end; ++t) {
            for (int mt = mg; mt < a.m_tiles; mt += MG) {
                const int b = mt * BM + q * 32 + lane;
                const bool row_ok = b < a.B;
                const long long row = (long long)t * a.B + b;
                // operands that do not depend on h_{t-1} are fetched before waiting for the accumulator
                float4 xi = make_float4(0.f, 0.f, 0.f, 0.f), xf = xi, xg = xi, xo = xi, cv = xi;
                float* cr = a.cell + (long long)b * a.D + u0;
                if (row_ok) {
                    const float* xr = a.xin + row * 4 * a.D + ns * 64 + hw * 4;
                    xi = *reinterpret_cast<const float4*>(xr);
                    xf = *reinterpret_cast<const float4*>(xr + 16);
                    xg = *reinterpret_cast<const float4*>(xr + 32);
                    xo = *reinterpret_cast<const float4*>(xr + 48);
                    if (t > 0) cv = (c_in_regs && t > a.t_begin) ? c_keep : *reinterpret_cast<const float4*>(cr);  // c_{-1} = 0
                }
                uint32_t ri[4] = {0u, 0u, 0u, 0u}, rf[4] = {0u, 0u, 0u, 0u}, rg[4] = {0u, 0u, 0u, 0u},
                         ro[4] = {0u, 0u, 0u, 0u};  // h_{-1} = 0
                if (t > 0) {
                    mbar_wait(tfull_bar(acc), acc_phase);
                    if (threadIdx.x == 64) stamp(t, 4);  // accumulator ready
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128 + hw * 4);
                    __syncwarp();
                    // eight 4-column loads (hh + lh blocks and the hl blocks 64 columns on) in flight, ONE wait
                    uint32_t ti2[4], tf2[4], tg2[4], to2[4];
                    tmem_ld_nowait(tb + 0, ri);
                    tmem_ld_nowait(tb + 16, rf);
                    tmem_ld_nowait(tb + 32, rg);
                    tmem_ld_nowait(tb + 48, ro);
                    tmem_ld_nowait(tb + 64 + 0, ti2);
                    tmem_ld_nowait(tb + 64 + 16, tf2);
                    tmem_ld_nowait(tb + 64 + 32, tg2);
                    tmem_ld_nowait(tb + 64 + 48, to2);
                    // the registers pass THROUGH the wait ("+r") so that no use can be scheduled above it
                    asm volatile("tcgen05.wait::ld.sync.aligned;"
                                 : "+r"(ri[0]), "+r"(ri[1]), "+r"(ri[2]), "+r"(ri[3]), "+r"(rf[0]), "+r"(rf[1]), "+r"(rf[2]),
                                   "+r"(rf[3]), "+r"(rg[0]), "+r"(rg[1]), "+r"(rg[2]), "+r"(rg[3]), "+r"(ro[0]), "+r"(ro[1]),
                                   "+r"(ro[2]), "+r"(ro[3]), "+r"(ti2[0]), "+r"(ti2[1]), "+r"(ti2[2]), "+r"(ti2[3]),
                                   "+r"(tf2[0]), "+r"(tf2[1]), "+r"(tf2[2]), "+r"(tf2[3]), "+r"(tg2[0]), "+r"(tg2[1]),
                                   "+r"(tg2[2]), "+r"(tg2[3]), "+r"(to2[0]), "+r"(to2[1]), "+r"(to2[2]), "+r"(to2[3])
                                 :
                                 : "memory");
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ri[i] = __float_as_uint(__uint_as_float(ri[i]) + __uint_as_float(ti2[i]));
                        rf[i] = __float_as_uint(__uint_as_float(rf[i]) + __uint_as_float(tf2[i]));
                        rg[i] = __float_as_uint(__uint_as_float(rg[i]) + __uint_as_float(tg2[i]));
                        ro[i] = __float_as_uint(__uint_as_float(ro[i]) + __uint_as_float(to2[i]));
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));  // accumulator is in registers: free it early
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                if (y_pend_row >= 0) {  // y of the previous step (see above)
                    *reinterpret_cast<float4*>(a.y + y_pend_row * a.D + u0) = y_pend;
                    y_pend_row = -1;
                }
                float4 c_out = cv, y_out = cv;
                if (row_ok) {
                    const float xiv[4] = {xi.x, xi.y, xi.z, xi.w}, xfv[4] = {xf.x, xf.y, xf.z, xf.w};
                    const float xgv[4] = {xg.x, xg.y, xg.z, xg.w}, xov[4] = {xo.x, xo.y, xo.z, xo.w};
                    float cvv[4] = {cv.x, cv.y, cv.z, cv.w}, hv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float ig = sigmoid_fast(__uint_as_float(ri[i]) + xiv[i]);
                        const float fg = sigmoid_fast(__uint_as_float(rf[i]) + xfv[i]);
                        const float gg = tanh_fast(__uint_as_float(rg[i]) + xgv[i]);
                        const float og = sigmoid_fast(__uint_as_float(ro[i]) + xov[i]);
                        cvv[i] = fg * cvv[i] + ig * gg;
                        hv[i] = og * tanh_fast(cvv[i]);
                    }
                    // h planes first: they are what the other CTAs wait for
                    uint32_t h01, l01, h23, l23;
                    split2(hv[0], hv[1], h01, l01);
                    split2(hv[2], hv[3], h23, l23);
                    if (LSTM_HST > 0 && a.publish == 3 && (mt + 1) * BM <= a.B) {
                        // staged: [128 rows x 32 B] per plane under the 32-byte TMA swizzle (16-byte chunk ^= bit 2 of the row)
                        const int r = q * 32 + lane;
                        const uint32_t so = (uint32_t)(r * 32 + ((((hw >> 1) ^ (r >> 2)) & 1) << 4) + (hw & 1) * 8);
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(hst_base + so), "r"(h01), "r"(h23) : "memory");
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(hst_base + BM * 32 + so), "r"(l01), "r"(l23) : "memory");
                    } else {
                    *reinterpret_cast<uint2*>(a.h_hi + row * a.D + u0) = make_uint2(h01, h23);
                    *reinterpret_cast<uint2*>(a.h_lo + row * a.D + u0) = make_uint2(l01, l23);
                    }
                    c_out = make_float4(cvv[0], cvv[1], cvv[2], cvv[3]);
                    y_out = make_float4(hv[0], hv[1], hv[2], hv[3]);
                }
                // publish h_t of this (batch tile, step): generic writes -> async-proxy (TMA) readers in other CTAs.
                // Each warp arrives on its own (lanes fence, __syncwarp orders them before lane 0's release) on the
                // counter of the k-block this CTA's 16 hidden units belong to.
                if (threadIdx.x == 64) stamp(t, 5);  // cell math + stores issued
                // publish = 1: the stores of the warp's lanes are ordered before lane 0's release by __syncwarp (the release
                // is cumulative over writes that happen-before it), so ONE fence round trip per warp instead of two; the
                // generic -> async proxy ordering is established on the consumer side (fence.proxy.async after its acquire)
                if (a.publish >= 2) {
                    // ONE arrival per CTA: the 16 epilogue warps meet at a named barrier (which orders their h stores before
                    // the releasing thread), then one thread adds all 16 arrivals. 4 atomics per k-block counter instead of
                    // 64 (same-address atomics serialise in L2), one release fence per CTA instead of 16.
                    asm volatile("bar.sync 1, %0;" ::"r"(LSTM_EPI_WARPS * 32) : "memory");
                    if (threadIdx.x == 64) stamp(t, 7);  // all 16 epilogue warps have stored their h
                    if (threadIdx.x == 64) {
                        if (LSTM_HST > 0 && a.publish == 3 && (mt + 1) * BM <= a.B) {
                            // publish = 3: the CTA's [128 x 16] slice of h_t leaves as TWO bulk tensor stores from the staging
                            // tile instead of 1024 scattered 8-byte stores; the release then has nothing else to wait for
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            const int r0 = (int)((long long)t * a.B) + mt * BM;
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                         ::"l"(reinterpret_cast<uint64_t>(&mapS_hi)), "r"(ns * 16), "r"(r0), "r"(hst_base) : "memory");
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                         ::"l"(reinterpret_cast<uint64_t>(&mapS_lo)), "r"(ns * 16), "r"(r0), "r"(hst_base + BM * 32) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                            asm volatile("fence.proxy.async;" ::: "memory");
                        }
                        int* cnt = a.counters + (((long long)mt * a.L + t) * LSTM_KB + (ns >> 2)) * LSTM_CNT_PITCH;
                        asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(cnt), "r"(LSTM_EPI_WARPS) : "memory");
                    }
                } else {
                if (a.publish == 0) {
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __threadfence();
                }
                __syncwarp();
                if (lane == 0) {
                    int* cnt = a.counters + (((long long)mt * a.L + t) * LSTM_KB + (ns >> 2)) * LSTM_CNT_PITCH;
                    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(cnt) : "memory");
                }
                }
                if (threadIdx.x == 64) stamp(t, 6);  // published
                // c_t (read back by this same thread at step t + 1) and y_t (read after the kernel) are stored AFTER the
                // release: the fences above then only wait for the two h-plane stores the other CTAs are waiting for
                if (row_ok) {
                    if (c_in_regs) {
                        c_keep = c_out;
                        if (t == a.t_end - 1) *reinterpret_cast<float4*>(cr) = c_out;
                    } else {
                        *reinterpret_cast<float4*>(cr) = c_out;
                    }
                    if (a.keep_c) { y_pend = y_out; y_pend_row = row; }
                    else *reinterpret_cast<float4*>(a.y + row * a.D + u0) = y_out;
                }
            }
        }
        if (y_pend_row >= 0) *reinterpret_cast<float4*>(a.y + y_pend_row * a.D + u0) = y_pend;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // peers may still multicast into this CTA's stages / commit onto its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        WT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) throw Error(4, "cuTensorMapEncodeTiled is not available");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp16 tensor: `rows` rows of `inner` elements, row r starting r*stride elements after the base
// (stride < inner: overlapping rows) -> tensor map with a [box_rows, 64] box and 128 B swizzle.
// Descriptors are cached: the workspace arena hands out the same addresses call after call, and encoding
// one costs ~10 us of host time (six per launch would leave the GPU idle between the ~700 launches of a step).
struct MapKey {
    const void* base; long long rows, inner, stride; int box, kw, dev;
    bool operator==(const MapKey& o) const {
        return base == o.base && rows == o.rows && inner == o.inner && stride == o.stride && box == o.box && kw == o.kw &&
               dev == o.dev;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.base);
        auto mix = [&](size_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
        mix((size_t)k.rows); mix((size_t)k.inner); mix((size_t)k.stride); mix((size_t)k.box); mix((size_t)k.kw); mix((size_t)k.dev);
        return h;
    }
};

CUtensorMap encode_map(const __half* base, long long rows, long long inner, long long stride, int box_rows, int kw);

const CUtensorMap& make_map(const __half* base, long long rows, long long inner, long long stride, int box_rows,
                            int kw = BK) {
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    static std::mutex mu;
    if (rows < 1) rows = 1;
    std::lock_guard<std::mutex> lock(mu);
    MapKey key{base, rows, inner, stride, box_rows, kw, current_device()};  // device addresses repeat across GPUs
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    if (cache.size() > (1u << 16)) cache.clear();
    return cache.emplace(key, encode_map(base, rows, inner, stride, box_rows, kw)).first->second;
}

CUtensorMap encode_map(const __half* base, long long rows, long long inner, long long stride, int box_rows, int kw) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)stride * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)kw, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = kw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : kw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box,
                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw Error(4, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r) + " (rows " +
                           std::to_string(rows) + ", inner " + std::to_string(inner) + ", stride " +
                           std::to_string(stride) + ", kw " + std::to_string(kw) + ")");
    return m;
}

int num_sms() {
    static PerDevice<int> cache;
    int& n = cache.get();
    if (!n) {
        int dev = 0;
        WT_CUDA(cudaGetDevice(&dev));
        WT_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

template <int BN, int PASSES, int EPI = 0, int CLM = 0>
void launch_cfg(const TcGemm& g, cudaStream_t s) {
    constexpr bool CL2 = CLM != 0;
    using C = Cfg<BN, PASSES, CLM == 2>;
    auto kernel = tap_gemm_tc_kernel<BN, PASSES, EPI, CLM>;
    static PerDevice<bool> attr_dev;
    static PerDevice<int> max_ctas_dev;  // cluster variant: CTAs that can be co-resident as pairs
    bool& attr = attr_dev.get();
    int& max_ctas = max_ctas_dev.get();
    constexpr int SMEM = C::SMEM + (EPI == 2 ? C::BIAS_SMEM : 0);
    if (EPI == 2 && g.N * (int)sizeof(float) > C::BIAS_SMEM) throw Error(4, "gemm_tc: GELU-only epilogue needs N <= 3072");
    if (!attr) {
        WT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        if (CL2) {
            cudaLaunchConfig_t qc = {};
            qc.gridDim = dim3(num_sms() & ~1); qc.blockDim = dim3(NUM_THREADS); qc.dynamicSmemBytes = SMEM;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            int nc = 0;
            WT_CUDA(cudaOccupancyMaxActiveClusters(&nc, kernel, &qc));
            max_ctas = 2 * nc;
            if (max_ctas < 2) throw Error(4, "gemm_tc: no room for a 2-CTA cluster");
        }
        attr = true;
    }
    Maps maps;
    for (int si = 0; si < 2; ++si) {
        const TcSeg& sg = g.seg[si < g.nseg ? si : 0];
        maps.a[si][0] = make_map(sg.hi, sg.rows, sg.inner, sg.stride, BM, g.kw);
        maps.a[si][1] = make_map(PASSES == 3 ? sg.lo : sg.hi, sg.rows, sg.inner, sg.stride, BM, g.kw);
    }
    const long long ldw = g.ldw ? g.ldw : g.K, w_rows = g.w_rows ? g.w_rows : g.N;
    const int wbox = CL2 ? BN / 2 : BN;
    maps.w[0] = make_map(g.W_hi, w_rows, g.K, ldw, wbox, g.kw);
    maps.w[1] = make_map(PASSES == 3 ? g.W_lo : g.W_hi, w_rows, g.K, ldw, wbox, g.kw);
    const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
    last_launch_info().kern = BN * 10 + PASSES;
    last_launch_info().flops = 2.0 * g.M * g.N * g.K * g.batch;
    if (CL2) {
        const int items = ((m_tiles + 1) / 2) * n_tiles;
        int grid = std::min(2 * items, std::min(num_sms() & ~1, max_ctas));
        if (g.max_ctas > 0) grid = std::max(2, std::min(grid, g.max_ctas & ~1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        WT_CUDA(cudaLaunchKernelEx(&cfg, kernel, maps, g));
    } else {
        const int tiles = m_tiles * n_tiles * g.batch;
        int grid = tiles < num_sms() ? tiles : num_sms();
        if (g.max_ctas > 0) grid = std::max(1, std::min(grid, g.max_ctas));
        kernel<<<grid, NUM_THREADS, SMEM, s>>>(maps, g);
    }
    WT_CUDA(cudaGetLastError());
}

template <int PASSES>
void launch_bn(const TcGemm& g, cudaStream_t s) {
    if (g.act == TC_ACT_LSTM) return launch_cfg<64, PASSES, 1>(g, s);
    if (g.N <= 16) return launch_cfg<16, PASSES>(g, s);
    if (g.N <= 32) return launch_cfg<32, PASSES>(g, s);
    if (g.N <= 64) return launch_cfg<64, PASSES>(g, s);
    if (g.N <= 128) {
        // WT_TC_N128_MC=1: 128-column tiles with a long K (the level-1 strided conv: K = 512, W re-read per tile from L2,
        // 4 x the DRAM bytes) as CTA pairs with the W halves multicast into both CTAs
        if constexpr (PASSES == 3) {
            static const int mc = [] { const char* e = std::getenv("WT_TC_N128_MC"); return e ? std::atoi(e) : 0; }();
            const int mt = (g.M + BM - 1) / BM;
            if (mc && g.N == 128 && g.K >= 256 && g.batch == 1 && g.kw == 64 && mt >= 2) return launch_cfg<128, 3, 0, 1>(g, s);
        }
        return launch_cfg<128, PASSES>(g, s);
    }
    const bool wide = g.N % 256 == 0 || g.N > 1024;
    if (!wide) return launch_cfg<128, PASSES>(g, s);
    // wide tiles run as CTA pairs: WT_TC_CLUSTER=2 one tcgen05.mma.cta_group::2 of M = 256 per pair (each CTA holds half of
    // W), =1 two cta_group::1 MMAs with the W halves multicast into both CTAs, =0 single CTAs
    static const int cluster_mode = [] { const char* e = std::getenv("WT_TC_CLUSTER"); return e ? std::atoi(e) : 2; }();
    const int m_tiles = (g.M + BM - 1) / BM;
    const bool pair_ok = g.batch == 1 && g.kw == 64 && m_tiles >= 2;
    if constexpr (PASSES == 1) {
        // ConvNeXt GEMM-1 (bias + erf-GELU -> one fp16 plane, whole column tiles): the specialised epilogue instantiation
        const bool gelu_only = g.act == TC_ACT_GELU && g.out_hi && !g.out_lo && !g.out_f32 && !g.elu_hi && !g.map.Pin &&
                               !g.gamma && !g.res && g.bias && g.N % 256 == 0;
        if (gelu_only && pair_ok && cluster_mode == 2) return launch_cfg<256, 1, 2, 2>(g, s);
    }
    if (pair_ok && cluster_mode == 2) launch_cfg<256, PASSES, 0, 2>(g, s);
    else if (pair_ok && cluster_mode == 1) launch_cfg<256, PASSES, 0, 1>(g, s);
    else launch_cfg<256, PASSES>(g, s);
}

// fp32 -> split fp16 planes (weights at load time; activations whose producer is not fused yet)
__global__ void split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                                 long long rows, int cols, long long ld_in, long long ld_out) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = rows * (ld_out / 2);
    if (gid >= total) return;
    long long r = gid / (ld_out / 2);
    int c = (int)(gid - r * (ld_out / 2)) * 2;
    float a = c < cols ? x[r * ld_in + c] : 0.f;
    float b = c + 1 < cols ? x[r * ld_in + c + 1] : 0.f;
    __half ha = __float2half_rn(a), hb = __float2half_rn(b);
    *reinterpret_cast<__half2*>(hi + r * ld_out + c) = __halves2half2(ha, hb);
    if (lo)
        *reinterpret_cast<__half2*>(lo + r * ld_out + c) =
            __halves2half2(__float2half_rn(a - __half2float(ha)), __float2half_rn(b - __half2float(hb)));
}

}  // namespace

LaunchInfo& last_launch_info() {
    static thread_local LaunchInfo info;
    return info;
}

static long long* g_debug_timeline = nullptr;
void set_debug_timeline(long long* dev_buf) { g_debug_timeline = dev_buf; }

void launch_tap_gemm_tc(const TcGemm& g_in, cudaStream_t s) {
    TcGemm g = g_in;
    if (!g.dbg) g.dbg = g_debug_timeline;
    if (g.M <= 0 || g.N <= 0) return;
    int kbs = 0;
    for (int si = 0; si < g.nseg; ++si) {
        const TcSeg& sg = g.seg[si];
        if (sg.stride % 8 != 0) throw Error(4, "gemm_tc: A row stride must be a multiple of 8 (16-byte TMA pitch)");
        if (sg.num_kb < 1 || sg.kb_per_tap < 1) throw Error(4, "gemm_tc: empty A segment");
        if ((reinterpret_cast<uintptr_t>(sg.hi) & 15) || (reinterpret_cast<uintptr_t>(sg.lo) & 15))
            throw Error(4, "gemm_tc: A planes must be 16-byte aligned");
        kbs += sg.num_kb;
    }
    if (g.kw != 64 && g.kw != 32 && g.kw != 16) throw Error(4, "gemm_tc: k-block width must be 64, 32 or 16");
    if (g.kw != 64 && (g.batch != 1 || g.act == TC_ACT_LSTM || g.act == TC_ACT_ARGMIN))
        throw Error(4, "gemm_tc: narrow k-blocks are for plain conv GEMMs");
    if (g.nseg < 1 || g.nseg > 2 || g.K != kbs * g.kw) throw Error(4, "gemm_tc: K must equal kw * total k-blocks");
    if ((g.out_f32 && g.ldo % 4) || (g.res && g.ldres % 4) || (g.out_hi && g.ldh % 8) || (g.elu_hi && g.ldh2 % 8))
        throw Error(4, "gemm_tc: output pitches must keep 16-byte alignment");
    if (g.passes != 1 && g.passes != 3) throw Error(4, "gemm_tc: passes must be 1 or 3");
    if (g.act == TC_ACT_LSTM && (g.N % 64 || !g.cell || !g.res || g.map.Pin))
        throw Error(4, "gemm_tc: LSTM epilogue needs N % 64 == 0, a cell state and the input projection");
    if (g.act == TC_ACT_ARGMIN && (g.N % 256 || !g.best || !g.bias || g.map.Pin))
        throw Error(4, "gemm_tc: argmin epilogue needs N % 256 == 0, ||c||^2 and the packed best[] buffer");
    if (g.batch < 1 || (g.ldw && g.ldw % 8)) throw Error(4, "gemm_tc: bad batch / W pitch");
    if (g.passes == 3) launch_bn<3>(g, s); else launch_bn<1>(g, s);
}

size_t lstm_counter_ints(int B, int L) { return (size_t)((B + BM - 1) / BM) * L * LSTM_KB * LSTM_CNT_PITCH; }

void launch_lstm_persistent(const float* xin, float* y, __half* h_hi, __half* h_lo, float* cell, int* counters,
                            const __half* w_hi, const __half* w_lo, int B, int L, int D, cudaStream_t s, int t_begin,
                            int t_end) {
    if (D != 512) throw Error(4, "lstm_persistent: hidden size must be 512");
    // WT_LSTM_CLUSTER = 1 (default) / 2 / 4: CTAs per cluster sharing the multicast h tile (1: every CTA loads its own copy)
    static const int cl = [] { const char* e = std::getenv("WT_LSTM_CLUSTER"); const int v = e ? std::atoi(e) : 1; return (v == 1 || v == 2 || v == 4) ? v : 1; }();
    const void* kernel = cl == 4 ? (const void*)lstm_persistent_kernel<4> : cl == 2 ? (const void*)lstm_persistent_kernel<2>
                                                                                       : (const void*)lstm_persistent_kernel<1>;
    static PerDevice<bool> attr_dev;
    bool& attr = attr_dev.get();
    if (!attr) {
        WT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LSTM_SMEM));
        attr = true;
    }
    LstmArgs a;
    a.xin = xin; a.y = y; a.h_hi = h_hi; a.h_lo = h_lo; a.cell = cell; a.counters = counters;
    a.B = B; a.L = L; a.D = D; a.m_tiles = (B + BM - 1) / BM;
    if (t_end < 0) t_end = L;
    if (t_begin < 0 || t_begin >= t_end || t_end > L) throw Error(4, "lstm_persistent: bad step range");
    a.t_begin = t_begin; a.t_end = t_end;
    // measured: publish 0 vs 1: 30.05 vs 30.20 ms per step (noise); 2 (one arrival per CTA): LSTM category 7.2-7.5 -> 6.3-6.6 ms
    // 3 (needs -DWT_LSTM_STAGES=2: bulk tensor store of the staged h slice behind the one release): LSTM kernel 4.76 vs 4.65 ms, not kept
    static const int publish = [] { const char* e = std::getenv("WT_LSTM_PUBLISH"); return e ? std::atoi(e) : 2; }();
    a.publish = publish;
    static const int poll_ns = [] { const char* e = std::getenv("WT_LSTM_POLL_NS"); return e ? std::atoi(e) : 0; }();
    a.poll_ns = poll_ns;
    static const int keep_c = [] { const char* e = std::getenv("WT_LSTM_KEEP_C"); return e ? std::atoi(e) : 1; }();
    a.keep_c = keep_c;
    a.dbg = g_debug_timeline ? g_debug_timeline + 148 * 64 : nullptr;  // after the generic GEMMs' per-CTA slots
    const int n_slices = 4 * D / 64;
    int mgroups = a.m_tiles;
    while (n_slices * mgroups > num_sms()) --mgroups;  // every CTA must be co-resident (they wait on each other)
    if (mgroups < 1) throw Error(4, "lstm_persistent: device too small");
    if (4 * D / 64 != LSTM_KB * 4 || LSTM_CTAS_PER_KB != 4 * LSTM_EPI_WARPS) throw Error(4, "lstm_persistent: slice / k-block mapping");
    CUtensorMap mw_hi = make_map(w_hi, 4LL * D, D, D, 64);
    CUtensorMap mw_lo = make_map(w_lo, 4LL * D, D, D, 64);
    if (t_begin == 0) WT_CUDA(cudaMemsetAsync(counters, 0, lstm_counter_ints(B, L) * sizeof(int), s));
    CUtensorMap mh_hi = make_map(h_hi, (long long)L * B, D, D, BM / cl);
    CUtensorMap mh_lo = make_map(h_lo, (long long)L * B, D, D, BM / cl);
    CUtensorMap ms_hi = make_map(h_hi, (long long)L * B, D, D, BM, 16);  // store maps: [128 rows x 16 units] boxes
    CUtensorMap ms_lo = make_map(h_lo, (long long)L * B, D, D, BM, 16);
    void* args[] = {&mh_hi, &mh_lo, &mw_hi, &mw_lo, &ms_hi, &ms_lo, &a};
    // cooperative: every CTA must be co-resident (they wait on each other); clusters along the gate-slice axis
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_slices, mgroups); cfg.blockDim = dim3(LSTM_THREADS); cfg.dynamicSmemBytes = LSTM_SMEM; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = cl; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = cl > 1 ? 2 : 1;
    WT_CUDA(cudaLaunchKernelExC(&cfg, kernel, args));
}

int lstm_ctas(int B, int D) {
    const int n_slices = 4 * D / 64;
    int mgroups = (B + BM - 1) / BM;
    while (mgroups > 1 && n_slices * mgroups > num_sms()) --mgroups;
    return n_slices * mgroups;
}

void launch_split_f16(const float* x, __half* hi, __half* lo, long long rows, int cols, long long ld_in,
                      long long ld_out, cudaStream_t s) {
    if (rows <= 0) return;
    if (ld_out % 2) throw Error(4, "split_f16: ld_out must be even");
    long long total = rows * (ld_out / 2);
    split_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, hi, lo, rows, cols, ld_in, ld_out);
    WT_CUDA(cudaGetLastError());
}

const CUtensorMap& tc_make_map(const __half* base, long long rows, long long inner, long long stride, int box_rows, int kw) {
    return make_map(base, rows, inner, stride, box_rows, kw);
}
int tc_num_sms() { return num_sms(); }
long long* tc_debug_timeline() { return g_debug_timeline; }

}  // namespace wt
