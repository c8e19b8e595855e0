// tcgen05 tap-GEMM for sm_100a (plan 1): the tensor-core contraction behind every Conv1d / Linear
// of the decoder (reference decoder/models.py:58-127,177; decoder/modules.py:43-60; decoder/heads.py:53),
// the LSTM input projection and the final encoder conv.
//
//   out[m, n] = epi( sum_{j<taps} sum_{c<Cin} A[m + j - center, c] * W[n, j*Cin + c] )
//
// * Operands are fp16 "split" planes: x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi) (~22 mantissa
//   bits). PASSES = 3 issues hi*hi + hi*lo + lo*hi per k-step (fp32 accumulate in TMEM), which keeps the
//   path inside the parity bars of BASELINE.json; PASSES = 1 uses the hi planes only.
// * A rows live in a "padded row space": clip b owns rows [b*Lp, b*Lp + L) followed by Lp - L zero rows, so
//   a k-tap Conv1d is k row-shifted TMA loads of the same 2-D tensor (out-of-range rows are zero-filled by
//   TMA); no im2col and no per-tap index arithmetic in the mainloop.
// * One CTA per SM, persistent over output tiles (128 x BN). Warp 0 = TMA producer, warp 1 = MMA issuer
//   (single thread, tcgen05.mma cta_group::1 kind::f16) + TMEM allocator, warps 2..5 = epilogue
//   (tcgen05.ld 32x32b -> bias / GELU / layer-scale / residual -> fp32 and/or split-fp16 stores).
//   Two TMEM accumulator stages let the epilogue of tile i overlap the mainloop of tile i+1.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace wt {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // fp16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
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

// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = BN.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
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

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int BN, int PASSES>
struct Cfg {
    static constexpr int A_PLANE = BM * BK * 2;  // bytes
    static constexpr int B_PLANE = BN * BK * 2;
    static constexpr int PLANES = PASSES == 3 ? 2 : 1;
    static constexpr int STAGE = PLANES * (A_PLANE + B_PLANE);
    static constexpr int STAGES = (200 * 1024) / STAGE > 8 ? 8 : (200 * 1024) / STAGE;
    static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages; 256 or 512 (power of two)
};

template <int BN, int PASSES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tap_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                   const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
                   const TcGemm g) {
    using C = Cfg<BN, PASSES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + C::STAGES * C::STAGE;
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base address slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
    uint32_t* tmem_slot_ptr =
        reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (g.M + BM - 1) / BM;
    const int n_tiles = (g.N + BN - 1) / BN;
    const int total_tiles = m_tiles * n_tiles;
    const int kb_per_tap = g.Cin / BK;
    const int num_kb = g.taps * kb_per_tap;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
                const int m0 = mt * BM, n0 = nt * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / kb_per_tap;
                    const int c0 = (kb - tap * kb_per_tap) * BK;
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * C::STAGE;
                    mbar_expect_tx(full_bar(stage), C::STAGE);
                    tma_load_2d(sa, &mapA_hi, c0, m0 + tap - g.center, full_bar(stage));
                    if (PASSES == 3) tma_load_2d(sa + C::A_PLANE, &mapA_lo, c0, m0 + tap - g.center, full_bar(stage));
                    const uint32_t sb = sa + C::PLANES * C::A_PLANE;
                    tma_load_2d(sb, &mapW_hi, kb * BK, n0, full_bar(stage));
                    if (PASSES == 3) tma_load_2d(sb + C::B_PLANE, &mapW_lo, kb * BK, n0, full_bar(stage));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_base + stage * C::STAGE;
                    const uint32_t sb = sa + C::PLANES * C::A_PLANE;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 2;  // bytes inside the 128 B swizzle row
                        const uint64_t a_hi = umma_desc_sw128(sa + koff);
                        const uint64_t b_hi = umma_desc_sw128(sb + koff);
                        umma_f16(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0);
                        if (PASSES == 3) {
                            const uint64_t a_lo = umma_desc_sw128(sa + C::A_PLANE + koff);
                            const uint64_t b_lo = umma_desc_sw128(sb + C::B_PLANE + koff);
                            umma_f16(tmem_d, a_hi, b_lo, idesc, 1);
                            umma_f16(tmem_d, a_lo, b_hi, idesc, 1);
                        }
                    }
                    umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
            const int m = mt * BM + q * 32 + lane;
            const int n0 = nt * BN;
            mbar_wait(tfull_bar(acc), acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool row_ok = m < g.M;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                __syncwarp();  // tcgen05.ld is .sync.aligned: re-converge after the predicated stores below
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
                const int nb = n0 + c * 32;
                if (!row_ok || nb >= g.N) continue;
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                const bool full = nb + 32 <= g.N;
                if (full) {
                    if (g.bias) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 b = *reinterpret_cast<const float4*>(g.bias + nb + i);
                            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                        }
                    }
                    if (g.act == ACT_GELU) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
                    }
                    if (g.gamma) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 b = *reinterpret_cast<const float4*>(g.gamma + nb + i);
                            v[i] *= b.x; v[i + 1] *= b.y; v[i + 2] *= b.z; v[i + 3] *= b.w;
                        }
                    }
                    if (g.res) {
                        const float* rr = g.res + (long long)m * g.ldres + nb;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 b = *reinterpret_cast<const float4*>(rr + i);
                            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                        }
                    }
                    if (g.out_f32) {
                        float* o = g.out_f32 + (long long)m * g.ldo + nb;
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    }
                    if (g.out_hi) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            __half h0 = __float2half_rn(v[2 * i]), h1 = __float2half_rn(v[2 * i + 1]);
                            __half l0 = __float2half_rn(v[2 * i] - __half2float(h0));
                            __half l1 = __float2half_rn(v[2 * i + 1] - __half2float(h1));
                            hi[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                            lo[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                        }
                        uint4* oh = reinterpret_cast<uint4*>(g.out_hi + (long long)m * g.ldh + nb);
#pragma unroll
                        for (int i = 0; i < 4; ++i) oh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                        if (g.out_lo) {
                            uint4* ol = reinterpret_cast<uint4*>(g.out_lo + (long long)m * g.ldh + nb);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                ol[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                        }
                    }
                } else {
                    for (int i = 0; i < 32 && nb + i < g.N; ++i) {
                        float x = v[i];
                        const int n = nb + i;
                        if (g.bias) x += g.bias[n];
                        if (g.act == ACT_GELU) x = gelu_erf(x);
                        if (g.gamma) x *= g.gamma[n];
                        if (g.res) x += g.res[(long long)m * g.ldres + n];
                        if (g.out_f32) g.out_f32[(long long)m * g.ldo + n] = x;
                        if (g.out_hi) {
                            __half h = __float2half_rn(x);
                            g.out_hi[(long long)m * g.ldh + n] = h;
                            if (g.out_lo) g.out_lo[(long long)m * g.ldh + n] = __float2half_rn(x - __half2float(h));
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
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

// 2-D fp16 row-major [rows, cols] (row pitch ld elements) -> tensor map with a [box_rows, 64] box, 128 B swizzle.
CUtensorMap make_map(const __half* base, long long rows, long long cols, long long ld, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box,
                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(4, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        WT_CUDA(cudaGetDevice(&dev));
        WT_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

template <int BN, int PASSES>
void launch_cfg(const TcGemm& g, cudaStream_t s) {
    using C = Cfg<BN, PASSES>;
    static bool attr = false;
    if (!attr) {
        WT_CUDA(cudaFuncSetAttribute(tap_gemm_tc_kernel<BN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM));
        attr = true;
    }
    CUtensorMap a_hi = make_map(g.A_hi, g.rowsA, g.Cin, g.lda, BM);
    CUtensorMap a_lo = make_map(PASSES == 3 ? g.A_lo : g.A_hi, g.rowsA, g.Cin, g.lda, BM);
    CUtensorMap w_hi = make_map(g.W_hi, g.N, g.K, g.K, BN);
    CUtensorMap w_lo = make_map(PASSES == 3 ? g.W_lo : g.W_hi, g.N, g.K, g.K, BN);
    const int tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    tap_gemm_tc_kernel<BN, PASSES><<<grid, NUM_THREADS, C::SMEM, s>>>(a_hi, a_lo, w_hi, w_lo, g);
    WT_CUDA(cudaGetLastError());
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

void launch_tap_gemm_tc(const TcGemm& g, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return;
    if (g.Cin % BK != 0 || g.K != g.taps * g.Cin) throw Error(4, "tap_gemm_tc: Cin must be a multiple of 64");
    if (g.lda % 8 != 0) throw Error(4, "tap_gemm_tc: lda must be a multiple of 8 (16-byte TMA pitch)");
    if ((g.out_f32 && g.ldo % 4) || (g.res && g.ldres % 4) || (g.out_hi && g.ldh % 8))
        throw Error(4, "tap_gemm_tc: output pitches must keep 16-byte alignment");
    if (g.passes != 1 && g.passes != 3) throw Error(4, "tap_gemm_tc: passes must be 1 or 3");
    const bool wide = g.N % 256 == 0 || g.N > 1024;
    if (g.passes == 3) {
        if (wide) launch_cfg<256, 3>(g, s); else launch_cfg<128, 3>(g, s);
    } else {
        if (wide) launch_cfg<256, 1>(g, s); else launch_cfg<128, 1>(g, s);
    }
}

void launch_split_f16(const float* x, __half* hi, __half* lo, long long rows, int cols, long long ld_in,
                      long long ld_out, cudaStream_t s) {
    if (rows <= 0) return;
    if (ld_out % 2) throw Error(4, "split_f16: ld_out must be even");
    long long total = rows * (ld_out / 2);
    split_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, hi, lo, rows, cols, ld_in, ld_out);
    WT_CUDA(cudaGetLastError());
}

}  // namespace wt
