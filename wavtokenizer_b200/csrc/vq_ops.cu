// Vector quantiser kernels (plan 0, fp32 CUDA cores) and codebook gathers.
// Replaces EuclideanCodebook.quantize / dequantize (reference encoder/quantization/core_vq.py:175-190)
// and WavTokenizer.codes_to_features (decoder/pretrained.py:209-239).
#include <cuda_fp16.h>

#include "common.cuh"

namespace wt {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

// One block owns 128 frames and sweeps the whole codebook: the [frames, bins] score matrix
// (and the one-hot the reference builds from it, core_vq.py:213) is never materialised.
// d = (||x||^2 - 2 x.c) + ||c||^2 in the reference's term order; argmin, first index on ties.
__global__ void __launch_bounds__(NT) vq_simt_kernel(const float* __restrict__ x, const float* __restrict__ cb,
                                                     const float* __restrict__ cnorm, long long N, int D, int bins,
                                                     long long* __restrict__ codes) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    __shared__ float xn[BM];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lane = tid & 31, warp = tid >> 5;
    const long long m0 = (long long)blockIdx.x * BM;

    // row norms: warp w handles rows w, w+8, ...
    for (int r = warp; r < BM; r += NT / 32) {
        long long m = m0 + r;
        float s = 0.f;
        if (m < N)
            for (int c = lane; c < D; c += 32) {
                float v = x[m * D + c];
                s = fmaf(v, v, s);
            }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) xn[r] = s;
    }

    int a_row[2], a_kq[2];
    bool a_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int idx = tid + i * NT;
        a_row[i] = idx >> 2;
        a_kq[i] = idx & 3;
        a_ok[i] = (m0 + a_row[i]) < N;
    }
    const int KT = D / BK;
    const int NTILES = bins / BN;
    const int total = KT * NTILES;
    float4 a_reg[2], b_reg[2];
    auto load_tile = [&](int it) {
        int nt = it / KT, kt = it - nt * KT;
        int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            a_reg[i] = a_ok[i] ? *reinterpret_cast<const float4*>(x + (m0 + a_row[i]) * D + k0 + a_kq[i] * 4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            int n = nt * BN + a_row[i];
            b_reg[i] = *reinterpret_cast<const float4*>(cb + (long long)n * D + k0 + a_kq[i] * 4);
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int kk = a_kq[i] * 4;
            As[buf][kk + 0][a_row[i]] = a_reg[i].x; As[buf][kk + 1][a_row[i]] = a_reg[i].y;
            As[buf][kk + 2][a_row[i]] = a_reg[i].z; As[buf][kk + 3][a_row[i]] = a_reg[i].w;
            Bs[buf][kk + 0][a_row[i]] = b_reg[i].x; Bs[buf][kk + 1][a_row[i]] = b_reg[i].y;
            Bs[buf][kk + 2][a_row[i]] = b_reg[i].z; Bs[buf][kk + 3][a_row[i]] = b_reg[i].w;
        }
    };

    float best[8];
    int besti[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { best[i] = INFINITY; besti[i] = 0; }
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int buf = it & 1;
        if (it + 1 < total) load_tile(it + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        const int nt = it / KT, kt = it - nt * KT;
        if (kt == KT - 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = nt * BN + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                const float cn = cnorm[n];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
                    float d = (xn[r] - 2.f * acc[i][j]) + cn;
                    if (d < best[i]) { best[i] = d; besti[i] = n; }
                    acc[i][j] = 0.f;
                }
            }
        }
        if (it + 1 < total) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }
    // reduce over the 16 threads (tx) that share each row; ties -> lower index
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = best[i];
        int bi = besti[i];
#pragma unroll
        for (int o = 8; o; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < v || (ov == v && oi < bi)) { v = ov; bi = oi; }
        }
        if (tx == 0) {
            long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (m < N) codes[m] = bi;
        }
    }
}

// quantized [N, D] = codebook[codes]  (row-major both sides)
__global__ void gather_rows_kernel(const float4* __restrict__ cb, const long long* __restrict__ codes,
                                   float4* __restrict__ out, long long N, int D4, int bins, int* err) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= N * D4) return;
    long long m = gid / D4;
    int c = (int)(gid - m * D4);
    long long code = codes[m];
    if (code < 0 || code >= bins) {
        if (err) atomicExch(err, 1);
        return;
    }
    out[gid] = cb[code * D4 + c];
}

// features[b, c, t] = sum_k codebooks[k*bins + codes[k, b, t]][c]; tile of 32 frames per block,
// transposed through shared memory so both the codebook reads (float4 per lane, the 32 codes of the tile staged
// once in shared memory, four independent row gathers in flight per thread) and the [B, D, L] writes coalesce.
__global__ void __launch_bounds__(256) codes_to_features_kernel(const float* __restrict__ cbs,
                                                                const long long* __restrict__ codes,
                                                                float* __restrict__ out, int K, int B, int L, int D,
                                                                int bins, int* err) {
    extern __shared__ float tile[];  // [32][D + 1]
    __shared__ int scode[32];
    const int b = blockIdx.y, t0 = blockIdx.x * 32;
    const int ldt = D + 1;
    const int D4 = D >> 2;
    for (int k = 0; k < K; ++k) {
        if (k) __syncthreads();
        if (threadIdx.x < 32) {
            const int t = t0 + threadIdx.x;
            long long code = t < L ? codes[((long long)k * B + b) * L + t] : 0;
            if (code < 0 || code >= bins) {
                if (err) atomicExch(err, 1);
                code = 0;
            }
            scode[threadIdx.x] = (int)code;
        }
        __syncthreads();
#pragma unroll 4
        for (int i = threadIdx.x; i < 32 * D4; i += blockDim.x) {
            const int tt = i / D4, c4 = i - tt * D4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t0 + tt < L)
                v = *reinterpret_cast<const float4*>(cbs + ((long long)k * bins + scode[tt]) * D + c4 * 4);
            float* tp = tile + tt * ldt + c4 * 4;
            if (k == 0) { tp[0] = v.x; tp[1] = v.y; tp[2] = v.z; tp[3] = v.w; }
            else { tp[0] += v.x; tp[1] += v.y; tp[2] += v.z; tp[3] += v.w; }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * D; i += blockDim.x) {
        int c = i >> 5, tt = i & 31;
        int t = t0 + tt;
        if (t < L) out[((long long)b * D + c) * L + t] = tile[tt * ldt + c];
    }
}

// frames [N, D] fp32 -> split-fp16 planes of (x - mu): the tcgen05 VQ reads frames centred on the codebook mean
__global__ void center_split_kernel(const float* __restrict__ x, const float* __restrict__ mu, __half* __restrict__ hi,
                                    __half* __restrict__ lo, long long n2, int D) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n2) return;
    const int c = (int)((gid * 2) % D);
    const float2 v = *reinterpret_cast<const float2*>(x + gid * 2);
    const float a = v.x - mu[c], b = v.y - mu[c + 1];
    const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
    *reinterpret_cast<__half2*>(hi + gid * 2) = __halves2half2(ha, hb);
    *reinterpret_cast<__half2*>(lo + gid * 2) =
        __halves2half2(__float2half_rn(a - __half2float(ha)), __float2half_rn(b - __half2float(hb)));
}

// packed (distance, index) keys of the argmin epilogue -> int64 codes
__global__ void best_to_codes_kernel(const unsigned long long* __restrict__ best, long long* __restrict__ codes,
                                     long long n) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) codes[gid] = (long long)(best[gid] & 0xFFFFFFFFull);
}

}  // namespace

void launch_center_split(const float* x, const float* mu, __half* hi, __half* lo, long long N, int D, cudaStream_t s) {
    if (N <= 0) return;
    long long n2 = N * D / 2;
    center_split_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, s>>>(x, mu, hi, lo, n2, D);
    WT_CUDA(cudaGetLastError());
}

void launch_best_to_codes(const unsigned long long* best, long long* codes, long long N, cudaStream_t s) {
    if (N <= 0) return;
    best_to_codes_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(best, codes, N);
    WT_CUDA(cudaGetLastError());
}

void launch_vq_simt(const float* x, const float* codebook, const float* cnorm, long long N, int D, int bins,
                    long long* codes, cudaStream_t s) {
    if (N <= 0) return;
    if (D % BK || bins % BN) throw Error(1, "vq: dimension must be a multiple of 16 and bins of 128");
    vq_simt_kernel<<<(unsigned)((N + BM - 1) / BM), NT, 0, s>>>(x, codebook, cnorm, N, D, bins, codes);
    WT_CUDA(cudaGetLastError());
}

void launch_gather_rows(const float* codebook, const long long* codes, float* out, long long N, int D, int bins,
                        int* err_flag, cudaStream_t s) {
    if (N <= 0) return;
    long long n = N * (D / 4);
    gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const float4*)codebook, codes, (float4*)out, N,
                                                                   D / 4, bins, err_flag);
    WT_CUDA(cudaGetLastError());
}

void launch_codes_to_features(const float* codebooks, const long long* codes, float* out, int K, int B, int L, int D,
                              int bins, int* err_flag, cudaStream_t s) {
    if (B <= 0 || L <= 0) return;
    if (D % 4) throw Error(1, "codes_to_features: codebook dim must be a multiple of 4");
    size_t smem = (size_t)32 * (D + 1) * sizeof(float);
    static PerDevice<bool> attr_dev;
    bool& attr_set = attr_dev.get();
    if (!attr_set) {
        WT_CUDA(cudaFuncSetAttribute(codes_to_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    dim3 grid((L + 31) / 32, B);
    codes_to_features_kernel<<<grid, 256, smem, s>>>(codebooks, codes, out, K, B, L, D, bins, err_flag);
    WT_CUDA(cudaGetLastError());
}

namespace {
// Codebook rows of a batch of codes straight into the decoder's padded row planes (what codes_to_features followed by
// features_to_rows produce, without the [B, D, L] fp32 round trip): row b*Lp + t <- planes of codebook[codes[b*L + t]] for
// t < L, zeros for the halo rows. One thread per 16 bytes of a plane row.
__global__ void codes_to_row_planes_kernel(const __half* __restrict__ cb_hi, const __half* __restrict__ cb_lo,
                                           const long long* __restrict__ codes, __half* __restrict__ hi,
                                           __half* __restrict__ lo, int L, int Lp, int D, int bins, long long total) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int per_row = D / 8;
    const long long row = gid / per_row;
    const int c8 = (int)(gid - row * per_row) * 8;
    const long long b = row / Lp;
    const int t = (int)(row - b * Lp);
    uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
    if (t < L) {
        long long code = codes[b * L + t];
        code = code < 0 ? 0 : (code >= bins ? bins - 1 : code);  // codes come from this library's own argmin
        vh = *reinterpret_cast<const uint4*>(cb_hi + code * D + c8);
        vl = *reinterpret_cast<const uint4*>(cb_lo + code * D + c8);
    }
    *reinterpret_cast<uint4*>(hi + row * D + c8) = vh;
    *reinterpret_cast<uint4*>(lo + row * D + c8) = vl;
}
}  // namespace

void launch_codes_to_row_planes(const __half* cb_hi, const __half* cb_lo, const long long* codes, __half* hi, __half* lo,
                                int B, int L, int Lp, int D, int bins, cudaStream_t s) {
    const long long total = (long long)B * Lp * (D / 8);
    if (total <= 0) return;
    codes_to_row_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cb_hi, cb_lo, codes, hi, lo, L, Lp, D, bins, total);
    WT_CUDA(cudaGetLastError());
}

}  // namespace wt
