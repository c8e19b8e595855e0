// tcgen05 / TMA / mbarrier PTX wrappers and epilogue helpers shared by the tensor-core kernels of this library
// (gemm_tc.cu: the generic tap GEMM and the persistent LSTM; enc_fused.cu: the fused encoder level). sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

#ifndef WT_TIMELINE
#define WT_TIMELINE 0
#endif

namespace wt {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // fp16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NEPI = 16;                       // epilogue warps: 4 per TMEM lane quarter, splitting the columns
constexpr int NUM_THREADS = 64 + NEPI * 32;    // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue
constexpr uint32_t SPIN_LIMIT = 1u << 28;  // a wedged pipeline traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > SPIN_LIMIT) asm volatile("trap;");
    } while (!ok);
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

// L2 prefetch of a tensor-map box (no shared-memory destination, no barrier): pulls the A rows of a k-block that is
// still a few stages away from HBM into L2, so that the TMA load that later fills the freed stage is an L2 hit.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1)
                 : "memory");
}

// K-major, 128-byte swizzled operand tile (rows of 64 fp16 = 128 B, 8-row groups 1024 B apart):
// start address >> 4 | LBO (ignored for swizzled K-major) | SBO = 1024 B | version 1 | SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Same, for a k-block of kw = 64 / 32 / 16 elements: rows of 128 / 64 / 32 bytes under the matching TMA swizzle
// (layout type 2 / 4 / 6, 8-row groups 8 * row bytes apart). `hi` carries every field but the start address.
__device__ __forceinline__ uint64_t umma_desc_hi(int kw) {
    const uint64_t layout = kw == 64 ? 2 : kw == 32 ? 4 : 6;
    const uint64_t sbo = (uint64_t)(8 * kw * 2);
    return ((uint64_t)1 << 16) | ((sbo >> 4) << 32) | ((uint64_t)1 << 46) | (layout << 61);
}
__device__ __forceinline__ uint64_t umma_desc_at(uint64_t hi, uint32_t smem_addr) {
    return hi | (uint64_t)((smem_addr & 0x3FFFF) >> 4);
}

// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = BN.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n, int m = BM) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// elect.sync: exactly one lane of a converged warp gets true. ptxas recognises an elect-guarded region as
// single-threaded, which a `lane == 0` test does not give it.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- 2-CTA cluster helpers (W tiles are loaded once per CTA pair and multicast into both CTAs) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose box lands at the same shared-memory offset of every CTA in `mask`, each CTA's mbarrier at the same
// offset receiving the complete_tx for the bytes written into it.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, "
        "%3}], [%4], %5;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
        : "memory");
}
// commit whose mbarrier arrive is delivered to the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}

// ---- CTA-pair MMA (cta_group::2): one tcgen05.mma of M = 256 runs on the tensor cores of BOTH SMs of a pair. Each CTA
// holds its own 128 A rows and HALF of the B (weight) rows in shared memory; each SM reads its A tile and its B half
// locally and receives the other B half from the peer, so the shared-memory reads per MMA drop from 12 KB to 8 KB per
// SM (N = 256) and a k-block occupies 32 KB instead of 48 KB per SM. The leader CTA (rank 0) issues the MMAs and owns
// the barriers the MMA thread waits on. ----
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// TMA load into THIS CTA's shared memory whose complete_tx goes to an mbarrier that may live in the peer CTA
// (`bar` is a shared::cluster address, e.g. the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit of the pair's MMAs: the arrive is delivered to the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}
// arrive on an mbarrier of another CTA of the cluster (`bar` is a shared::cluster address)
// Default semantics (.release at CTA scope): what is handed over is the TMEM accumulator stage, already ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync. A .release.cluster here costs MEMBAR.ALL.GPU + ERRBAR per tile,
// i.e. a wait for the tile's global stores (ncu: 14.6 % of the ConvNeXt GEMM-1 stall samples).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Exact-erf GELU (reference decoder/modules.py:35, nn.GELU()): see gelu_erf below.

// Two accumulator chunks (hh + lh columns and the hl columns BN further) in flight behind ONE wait; the registers
// pass through the wait ("+r") so that no use can be scheduled above it.
__device__ __forceinline__ void tmem_ld_pair(uint32_t a0, uint32_t (&r)[16], uint32_t a1, uint32_t (&t)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(a0));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
                 : "r"(a1));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_nowait(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}

__device__ __forceinline__ float gelu_erf(float x) {
    // erfc(z) = 2^P(z) on [0, 4], P of degree 7 without constant term (weighted minimax fit, |erf error| <= 1.6e-7
    // evaluated in fp32; erfc(4) = 1.5e-8 so z is clamped there): one MUFU.EX2, no reciprocal, no branch.
    const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.f);
    float p = fmaf(1.00181106e-04f, z, -4.61322709e-04f);
    p = fmaf(p, z, -2.30282884e-03f);
    p = fmaf(p, z, 2.94531747e-02f);
    p = fmaf(p, z, -1.48964027e-01f);
    p = fmaf(p, z, -9.18328559e-01f);
    p = fmaf(p, z, -1.62791374e+00f);
    float e2;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(p * z));  // bare MUFU.EX2 (exp2f adds a denormal-range fix-up)
    const float erfa = 1.f - e2;  // erf(|x| / sqrt(2))
    const float h = 0.5f * x;
    return fmaf(fabsf(h), erfa, h);         // 0.5 x (1 + sign(x) erf(|x|/sqrt 2))
}
// ELU(alpha = 1) (reference encoder/modules/seanet.py:37). Negative branch exp(x) - 1 through MUFU.EX2: absolute
// error <= ~1.2e-7 (one ulp of the exponential near 1), i.e. the 2^-22 resolution the value is then stored with
// in split-fp16 planes at the O(0.1..1) activation scale of the encoder; five instructions, no branch.
__device__ __forceinline__ float elu1(float x) {
    float e;  // bare MUFU.EX2 (__expf adds a denormal-range fix-up: FSETP + two predicated FMULs per call)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return x > 0.f ? x : e - 1.f;
}
__device__ __forceinline__ float sigmoid1(float x) { return 1.f / (1.f + expf(-x)); }
// LSTM gates on the recurrent critical path: MUFU.EX2 + MUFU.RCP forms (absolute error ~1e-7, the resolution of the
// split-fp16 planes h_t is published in) instead of the ~25-instruction expf / IEEE-divide / tanhf sequences.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
    const float xc = fminf(fmaxf(x, -15.f), 15.f);
    return fmaf(-2.f, __fdividef(1.f, 1.f + __expf(2.f * xc)), 1.f);
}

// Split a pair into packed hi / lo halves. Packed conversions (F2FP.F16.F32.PACK_AB, ALU pipe) instead of scalar
// F2F (XU pipe, 16 lanes/clk/SM): the narrow-N epilogues were conversion-throughput bound.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 256-bit global store (sm_100a STG.E.256): one full 32-byte sector per lane per instruction.
__device__ __forceinline__ void st256(void* p, const uint32_t* w) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

// Store CW consecutive values of one row as split planes (32-byte stores when the row is 32-byte aligned).
template <int CW, bool ELU>
__device__ __forceinline__ void store_planes(__half* hi_p, __half* lo_p, long long off, const float (&v)[CW]) {
    uint32_t hi[CW / 2], lo[CW / 2];
    if (lo_p) {
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) {
            float a = v[2 * i], b = v[2 * i + 1];
            if (ELU) { a = elu1(a); b = elu1(b); }
            split2(a, b, hi[i], lo[i]);
        }
    } else {  // single-pass consumer: only the hi plane is read
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) {
            float a = v[2 * i], b = v[2 * i + 1];
            if (ELU) { a = elu1(a); b = elu1(b); }
            const __half2 h2 = __floats2half2_rn(a, b);
            hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
    }
    if (CW % 16 == 0 && (off & 15) == 0) {
#pragma unroll
        for (int j = 0; j < CW / 16; ++j) {
            st256(hi_p + off + 16 * j, hi + 8 * j);
            if (lo_p) st256(lo_p + off + 16 * j, lo + 8 * j);
        }
        return;
    }
    uint4* oh = reinterpret_cast<uint4*>(hi_p + off);
#pragma unroll
    for (int i = 0; i < CW / 8; ++i) oh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
    if (lo_p) {
        uint4* ol = reinterpret_cast<uint4*>(lo_p + off);
#pragma unroll
        for (int i = 0; i < CW / 8; ++i) ol[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
    }
}

// ---- asynchronous TMEM loads: issue now, wait later (tcgen05.wait::ld waits for ALL outstanding loads of the thread) ----
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
// tcgen05.wait::ld with the loaded registers passing THROUGH it, so that no use can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void pin16(uint32_t (&r)[16]) {  // orders the uses of r[] after the preceding volatile asm
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                      "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
// non-blocking test of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// tcgen05.mma with the A operand in tensor memory (TS form): rows = TMEM lanes, the 16 K elements of a k-step packed two
// per 32-bit column (low half = even k), i.e. 8 columns per k-step; B from a shared-memory descriptor as usual
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

}  // namespace
}  // namespace wt
