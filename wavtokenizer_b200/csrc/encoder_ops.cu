// Encoder-side kernels that are not tap-GEMMs: the 1->C first convolution, the LSTM cell
// pointwise update and layout transposes.
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
