// fp32 CUDA-core tap-GEMM (plan 0): bit-conservative contraction used for every Conv1d / Linear
// of the path and as the on-device yardstick for the tcgen05 kernels.
// Replaces ATen conv1d / addmm call sites of reference encoder/modules/conv.py:195-211,
// decoder/models.py:58-127,177, decoder/modules.py:43-60, decoder/heads.py:53.
#include "common.cuh"

namespace wt {

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NT = 256;

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int TN>
__global__ void __launch_bounds__(NT) tap_gemm_simt_kernel(const TapGemm g) {
    constexpr int BN = 16 * TN;
    constexpr int LDA_S = BM + 4;
    constexpr int LDB_S = BN + 4;
    __shared__ __align__(16) float As[2][BK][LDA_S];
    __shared__ __align__(16) float Bs[2][BK][LDB_S];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- A loader state: two (row, k-quad) slots per thread ----
    const float* a_base[2];
    int a_t0[2];
    bool a_ok[2];
    int a_row[2], a_kq[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int idx = tid + i * NT;
        a_row[i] = idx >> 2;
        a_kq[i] = idx & 3;
        int m = m0 + a_row[i];
        a_ok[i] = m < g.M;
        int b = a_ok[i] ? m / g.Tout : 0;
        int t = a_ok[i] ? m - b * g.Tout : 0;
        a_base[i] = g.A + (long long)b * g.Tin * g.lda;
        a_t0[i] = t * g.stride - g.pad_left;
    }
    constexpr int B_SLOTS = (BN * 4 + NT - 1) / NT;

    float4 a_reg[2];
    float4 b_reg[B_SLOTS];

    auto load_tile = [&](int kt) {
        const int k0 = kt * BK;
        const int tap = (g.taps == 1) ? 0 : k0 / g.Cin;
        const int c0 = k0 - tap * g.Cin;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int ti = a_t0[i] + tap;
            bool ok = a_ok[i];
            if (g.pad_mode == PAD_REFLECT) {
                if (ti < 0) ti = -ti;
                if (ti >= g.Trefl) ti = 2 * (g.Trefl - 1) - ti;
                ok = ok && ti < g.Tin && ti >= 0;
            } else {
                ok = ok && ti >= 0 && ti < g.Tin;
            }
            if (ok) {
                v = *reinterpret_cast<const float4*>(a_base[i] + (long long)ti * g.lda + c0 + a_kq[i] * 4);
                if (g.pro == PRO_ELU) {
                    v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w);
                }
            }
            a_reg[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_SLOTS; ++i) {
            int idx = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < BN * 4) {
                int n = n0 + (idx >> 2);
                if (n < g.N) v = *reinterpret_cast<const float4*>(g.W + (long long)n * g.K + k0 + (idx & 3) * 4);
            }
            b_reg[i] = v;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int kk = a_kq[i] * 4;
            As[buf][kk + 0][a_row[i]] = a_reg[i].x;
            As[buf][kk + 1][a_row[i]] = a_reg[i].y;
            As[buf][kk + 2][a_row[i]] = a_reg[i].z;
            As[buf][kk + 3][a_row[i]] = a_reg[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_SLOTS; ++i) {
            int idx = tid + i * NT;
            if (idx < BN * 4) {
                int kk = (idx & 3) * 4, nn = idx >> 2;
                Bs[buf][kk + 0][nn] = b_reg[i].x;
                Bs[buf][kk + 1][nn] = b_reg[i].y;
                Bs[buf][kk + 2][nn] = b_reg[i].z;
                Bs[buf][kk + 3][nn] = b_reg[i].w;
            }
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int KT = g.K / BK;
    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) load_tile(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if constexpr (TN == 8) {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            } else if constexpr (TN == 4) {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            } else if constexpr (TN == 2) {
                float2 b0 = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
                b[0] = b0.x; b[1] = b0.y;
            } else {
                b[0] = Bs[buf][kk][tx];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < KT) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue: bias, activation, layer-scale, residual ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int nl;
            if constexpr (TN == 8) nl = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            else if constexpr (TN == 4) nl = tx * 4 + j;
            else if constexpr (TN == 2) nl = tx * 2 + j;
            else nl = tx;
            const int n = n0 + nl;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += g.bias[n];
            if (g.act == ACT_GELU) v = gelu_erf(v);
            if (g.gamma) v *= g.gamma[n];
            if (g.res) v += g.res[(long long)m * g.ldres + n];
            g.out[(long long)m * g.ldo + n] = v;
        }
    }
}

}  // namespace

void launch_tap_gemm_simt(const TapGemm& g, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return;
    if (g.K % BK != 0 || (g.taps > 1 && g.Cin % BK != 0) || g.lda % 4 != 0)
        throw Error(4, "tap_gemm_simt: K and Cin must be multiples of 16 and lda of 4");
    dim3 block(NT);
    auto grid = [&](int bn) { return dim3((g.M + BM - 1) / BM, (g.N + bn - 1) / bn); };
    if (g.N > 64) tap_gemm_simt_kernel<8><<<grid(128), block, 0, s>>>(g);
    else if (g.N > 32) tap_gemm_simt_kernel<4><<<grid(64), block, 0, s>>>(g);
    else if (g.N > 16) tap_gemm_simt_kernel<2><<<grid(32), block, 0, s>>>(g);
    else tap_gemm_simt_kernel<1><<<grid(16), block, 0, s>>>(g);
    WT_CUDA(cudaGetLastError());
}

}  // namespace wt
