// Encoder-side kernels that are not tap-GEMMs: the 1->C first convolution, the LSTM cell
// pointwise update and layout transposes.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace wt {

namespace {

// SConv1d(1 -> C, k=7, stride 1, reflect pad 3/3) (reference encoder/modules/seanet.py:107-110,
// conv.py:195-211). K = 7 is far too small for an MMA: HBM-bound, one thread per output sample,
// C channels written as float4 (channels-last).
template <int C>
__global__ void __launch_bounds__(256) conv0_kernel(const float* __restrict__ wav, const float* __restrict__ w,
                                                    const float* __restrict__ bias, float* __restrict__ out, int B,
                                                    int T, int Trefl) {
    __shared__ float ws[C * 7];
    __shared__ float bs[C];
    for (int i = threadIdx.x; i < C * 7; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) bs[i] = bias[i];
    __syncthreads();
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * T) return;
    int b = (int)(gid / T);
    int t = (int)(gid - (long long)b * T);
    const float* x = wav + (long long)b * T;
    float xv[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        int ti = t - 3 + j;
        if (ti < 0) ti = -ti;
        if (ti >= Trefl) ti = 2 * (Trefl - 1) - ti;
        xv[j] = (ti >= 0 && ti < T) ? x[ti] : 0.f;
    }
    float4* o = reinterpret_cast<float4*>(out + gid * C);
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
        float r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int c = c4 * 4 + u;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 7; ++j) acc = fmaf(ws[c * 7 + j], xv[j], acc);
            r[u] = acc + bs[c];
        }
        o[c4] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

// fast ELU, see gemm_tc.cu (same formula so that every producer of ELU planes agrees)
__device__ __forceinline__ float elu_fast(float x) {
    const float e = __expf(x) - 1.f;
    return x > 0.f ? x : e;
}

__device__ __forceinline__ void split_store8(__half* hi, __half* lo, long long off, const float (&v)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // packed conversions (ALU pipe) rather than scalar F2F (XU pipe)
        const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 f = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// conv0 for the tcgen05 encoder: same arithmetic as conv0_kernel, one thread per (PADDED position, group of 8
// channels) of the next layer's input layout (clip pitch T + 2, data at offset 1, reflect halo of 1 on both
// sides). Writes the split-fp16 planes of ELU(x) (k3-conv operand) and, instead of the planes of x itself, the
// 8-wide window of raw audio samples around the position: the ResBlock shortcut conv1x1(conv0(wav)) is a single
// k7 conv of the audio with composed weights (model.cu), which saves 96 B of HBM traffic per sample.
template <int C>
__global__ void __launch_bounds__(256) conv0_planes_kernel(const float* __restrict__ wav, const float* __restrict__ w,
                                                           const float* __restrict__ bias, __half* __restrict__ win_hi,
                                                           __half* __restrict__ win_lo, __half* __restrict__ elu_hi,
                                                           __half* __restrict__ elu_lo, int B, int T) {
    // One thread = 8 channels x 4 consecutive padded positions of one clip (blockIdx.y): its 56 taps stay in
    // registers, the 10 audio samples under the 4 windows are loaded once, and there is no index division.
    const int P = T + 2;
    constexpr int G8 = C / 8;  // 16-byte stores coalesce across the G8 threads of a position
    const int c8 = threadIdx.x % G8;
    const int p0 = (blockIdx.x * (256 / G8) + threadIdx.x / G8) * 4;
    const int b = blockIdx.y;
    if (p0 >= P) return;
    float wr[8][7], br[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        br[u] = bias[c8 * 8 + u];
#pragma unroll
        for (int j = 0; j < 7; ++j) wr[u][j] = w[(c8 * 8 + u) * 7 + j];
    }
    const float* x = wav + (long long)b * T;
    const long long row0 = (long long)b * P + p0;
    auto emit = [&](int q, const float (&xv)[8]) {  // position p0 + q from its 7-sample window
        float e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float acc = br[u];
#pragma unroll
            for (int j = 0; j < 7; ++j) acc = fmaf(wr[u][j], xv[j], acc);
            e[u] = elu_fast(acc);
        }
        split_store8(elu_hi, elu_lo, ((row0 + q) * G8 + c8) * 8, e);
        if (c8 == 0) split_store8(win_hi, win_lo, (row0 + q) * 8, xv);
    };
    if (p0 >= 4 && p0 + 6 <= T) {  // interior: samples t0 - 3 .. t0 + 6 with t0 = p0 - 1, no reflection
        float xs[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) xs[j] = x[p0 - 4 + j];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float xv[8];
#pragma unroll
            for (int j = 0; j < 7; ++j) xv[j] = xs[q + j];
            xv[7] = 0.f;
            emit(q, xv);
        }
    } else {  // clip edges: reflect the position (halo rows), then the taps (reference conv.py:79-96)
        for (int q = 0; q < 4 && p0 + q < P; ++q) {
            int t = p0 + q - 1;
            if (t < 0) t = -t;
            if (t >= T) t = 2 * (T - 1) - t;
            float xv[8];
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                int ti = t - 3 + j;
                if (ti < 0) ti = -ti;
                if (ti >= T) ti = 2 * (T - 1) - ti;
                xv[j] = x[ti];
            }
            xv[7] = 0.f;
            emit(q, xv);
        }
    }
}

// SLSTM skip connection y + x (reference encoder/modules/lstm.py:38; y and x in time-major rows) fused with the ELU in front of the last
// encoder conv: writes fp32 rows [B*L, D] (tap) and the split planes of ELU(y + x) in the reflect-padded layout
// of the k7 conv (clip pitch L + 6, data at offset 3).
__global__ void lstm_skip_elu_pad_kernel(const float* __restrict__ y, const float* __restrict__ x,
                                         float* __restrict__ out_f32, __half* __restrict__ elu_hi,
                                         __half* __restrict__ elu_lo, int B, int L, int D) {
    const int P = L + 6;
    const int d8 = D / 8;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * P * d8) return;
    int c8 = (int)(gid % d8);
    long long row = gid / d8;
    int b = (int)(row / P), p = (int)(row - (long long)b * P);
    int t = p - 3;
    const bool interior = t >= 0 && t < L;
    if (t < 0) t = -t;
    if (t >= L) t = 2 * (L - 1) - t;
    const long long src = ((long long)t * B + b) * D + c8 * 8;   // y, x: time-major rows [t*B + b]
    const long long dst = ((long long)b * L + t) * D + c8 * 8;   // fp32 copy: batch-major rows
    float v[8], e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = y[src + i] + x[src + i];
        e[i] = elu_fast(v[i]);
    }
    if (interior && out_f32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) out_f32[dst + i] = v[i];
    }
    split_store8(elu_hi, elu_lo, row * D + c8 * 8, e);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// LSTM cell update (reference encoder/modules/lstm.py:20 -> nn.LSTM; gate order i, f, g, o).
// gates [B, 4H] already hold W_ih x + b_ih + b_hh + W_hh h_{t-1}.
__global__ void lstm_pointwise_kernel(const float* __restrict__ gates, float* __restrict__ c, float* __restrict__ y,
                                      int B, int H, long long ldg, long long ldy) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * H) return;
    int b = gid / H, u = gid - b * H;
    const float* gr = gates + (long long)b * ldg;
    float ig = sigmoidf_(gr[u]);
    float fg = sigmoidf_(gr[H + u]);
    float gg = tanhf(gr[2 * H + u]);
    float og = sigmoidf_(gr[3 * H + u]);
    float cn = fg * c[gid] + ig * gg;
    c[gid] = cn;
    y[(long long)b * ldy + u] = og * tanhf(cn);
}

__global__ void add_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o,
                           long long n4) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 x = a[i], y = b[i];
    o[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

// [B, R, C] -> [B, C, R] tiled transpose (both API<->internal directions use it).
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
    __shared__ float tile[32][33];
    const float* ib = in + (long long)blockIdx.z * R * C;
    float* ob = out + (long long)blockIdx.z * R * C;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? ib[(long long)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) ob[(long long)c * R + r] = tile[threadIdx.x][i];
    }
}

}  // namespace

void launch_conv0(const float* wav, const float* w, const float* bias, float* out, int B, int T, int C,
                  cudaStream_t s) {
    if (C != 32) throw Error(1, "conv0: n_filters must be 32");
    long long n = (long long)B * T;
    int Trefl = T > 4 ? T : 4;  // reflect needs length > pad (3): zero-extend short signals (conv.py:86-94)
    conv0_kernel<32><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(wav, w, bias, out, B, T, Trefl);
    WT_CUDA(cudaGetLastError());
}

void launch_conv0_planes(const float* wav, const float* w, const float* bias, __half* win_hi, __half* win_lo,
                         __half* elu_hi, __half* elu_lo, int B, int T, int C, cudaStream_t s) {
    if (C != 32) throw Error(1, "conv0: n_filters must be 32");
    if (T < 4) throw Error(4, "conv0_planes: clip too short for the tcgen05 encoder layout");
    dim3 grid((unsigned)((T + 2 + 255) / 256), (unsigned)B);  // 256 padded positions per block
    conv0_planes_kernel<32><<<grid, 256, 0, s>>>(wav, w, bias, win_hi, win_lo, elu_hi, elu_lo, B, T);
    WT_CUDA(cudaGetLastError());
}

void launch_lstm_skip_elu_pad(const float* y, const float* x, float* out_f32, __half* elu_hi, __half* elu_lo, int B,
                              int L, int D, cudaStream_t s) {
    long long n = (long long)B * (L + 6) * (D / 8);
    lstm_skip_elu_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y, x, out_f32, elu_hi, elu_lo, B, L, D);
    WT_CUDA(cudaGetLastError());
}

void launch_lstm_pointwise(const float* gates, float* c, float* y, int B, int H, long long ldg, long long ldy,
                           cudaStream_t s) {
    int n = B * H;
    lstm_pointwise_kernel<<<(n + 255) / 256, 256, 0, s>>>(gates, c, y, B, H, ldg, ldy);
    WT_CUDA(cudaGetLastError());
}

void launch_add(const float* a, const float* b, float* out, long long n, cudaStream_t s) {
    if (n % 4) throw Error(4, "add: n % 4 != 0");
    long long n4 = n / 4;
    add_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>((const float4*)a, (const float4*)b, (float4*)out, n4);
    WT_CUDA(cudaGetLastError());
}

void launch_transpose_bcl_to_blc(const float* in, float* out, int B, int C, int L, cudaStream_t s) {
    // in [B, C, L]: rows = C, cols = L
    dim3 grid((L + 31) / 32, (C + 31) / 32, B), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(in, out, C, L);
    WT_CUDA(cudaGetLastError());
}

void launch_transpose_blc_to_bcl(const float* in, float* out, int B, int L, int C, cudaStream_t s) {
    dim3 grid((C + 31) / 32, (L + 31) / 32, B), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(in, out, L, C);
    WT_CUDA(cudaGetLastError());
}

}  // namespace wt
