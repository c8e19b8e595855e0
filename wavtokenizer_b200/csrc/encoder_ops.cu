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
    float e;  // bare MUFU.EX2 (__expf adds a denormal-range fix-up: FSETP + two predicated FMULs per call)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return x > 0.f ? x : e - 1.f;
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

// ---------------------------------------------------------------------------------------
// Level 0 of the SEANet encoder in ONE kernel, fp32 on the CUDA cores:
//   x0 = conv0(wav) (1 -> 32, k7, reflect)                       reference encoder/modules/seanet.py:107-110
//   h1 = conv_k3(ELU(x0)) (32 -> 16, reflect pad 1)              seanet.py:45-63 (block), conv.py:195-211
//   y  = conv1x1(ELU(h1)) + shortcut1x1(x0)  (-> 32)             seanet.py:61-63 (true_skip = False)
//   out = ELU(y) as split-fp16 planes in the padded layout of the strided conv that follows (seanet.py:123-129)
// The contractions are K = 7 / 96 / 16 + 7 wide with N = 32 / 16 / 32: as tcgen05 tiles they are bound by the
// per-tile epilogue latency, not by the MMA (profiles/r01_enc_chunk_s3_summary.md), and their operands (2.4 GB per
// 256 clips each) round-trip through HBM. Here a block owns 512 consecutive samples of a clip: ELU(x0) and ELU(h1)
// live in shared memory only, weights are staged in shared memory once per block, every thread register-tiles
// 4 positions x 8 (16) channels so that one 128-bit shared load feeds >= 8 FMAs, and HBM sees 4 B in and 128 B out
// per sample. The shortcut conv1x1(conv0(wav)) is the composed k7 conv of the raw audio (weights composed in fp64
// at load time). Bound: fp32 FMA issue (2.5 kFMA per sample).
// ---------------------------------------------------------------------------------------
constexpr int RB0_TILE = 512, RB0_THREADS = 256, RB0_LD = RB0_TILE + 4;
// packed weights (floats): w0t[7][32] b0[32] w1t[96][16] b1[16] w2t[16][32] wsct[8][32] b2[32]
constexpr int RB0_W0 = 0, RB0_B0 = 224, RB0_W1 = 256, RB0_B1 = 1792, RB0_W2 = 1808, RB0_WSC = 2320, RB0_B2 = 2576,
              RB0_PACK = 2608, RB0_PACK_PAD = 2624, RB0_A = RB0_TILE + 8, RB0_A_PAD = RB0_TILE + 16;
constexpr int RB0_SMEM = (RB0_PACK_PAD + RB0_A_PAD + 32 * RB0_LD + 16 * RB0_TILE) * 4;

__global__ void __launch_bounds__(RB0_THREADS, 2)
resblock0_fused_kernel(const float* __restrict__ wav, const float* __restrict__ pack, __half* __restrict__ ye_hi,
                       __half* __restrict__ ye_lo, float* __restrict__ y_f32, int T, int Py, int left, int hr) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm;
    float* sA = sW + RB0_PACK_PAD;
    float* sE0 = sA + RB0_A_PAD;          // [32][RB0_LD], column col <-> sample t = t0 - 1 + col
    float* sH1 = sE0 + 32 * RB0_LD;       // [16][RB0_TILE], ELU(h1)
    const int tid = threadIdx.x;
    const int b = blockIdx.y, t0 = blockIdx.x * RB0_TILE;
    const float* x = wav + (long long)b * T;

    // ---- stage 0: weights and the audio samples t0 - 4 .. t0 + TILE + 3 (tap reflection, conv.py:79-96) ----
    for (int i = tid; i < RB0_PACK / 4; i += RB0_THREADS)
        reinterpret_cast<float4*>(sW)[i] = reinterpret_cast<const float4*>(pack)[i];
    for (int i = tid; i < RB0_A; i += RB0_THREADS) {
        int idx = t0 - 4 + i;
        if (idx < 0) idx = -idx;
        if (idx >= T) idx = 2 * (T - 1) - idx;
        sA[i] = (idx >= 0 && idx < T) ? x[idx] : 0.f;
    }
    __syncthreads();

    // ---- stage A: ELU(conv0) for columns 0 .. TILE + 1: two columns per thread share every weight vector ----
    {
        float xa[7], xb[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) { xa[j] = sA[tid + j]; xb[j] = sA[tid + RB0_THREADS + j]; }
#pragma unroll 1  // (fully unrolled, the 56 weight vectors get hoisted: 224 registers)
        for (int c4 = 0; c4 < 8; ++c4) {
            const float4 bb = *reinterpret_cast<const float4*>(sW + RB0_B0 + c4 * 4);
            float4 a = bb, c = bb;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const float4 w = *reinterpret_cast<const float4*>(sW + RB0_W0 + j * 32 + c4 * 4);
                a.x = fmaf(w.x, xa[j], a.x); a.y = fmaf(w.y, xa[j], a.y);
                a.z = fmaf(w.z, xa[j], a.z); a.w = fmaf(w.w, xa[j], a.w);
                c.x = fmaf(w.x, xb[j], c.x); c.y = fmaf(w.y, xb[j], c.y);
                c.z = fmaf(w.z, xb[j], c.z); c.w = fmaf(w.w, xb[j], c.w);
            }
            float* e = sE0 + (c4 * 4) * RB0_LD + tid;
            e[0] = elu_fast(a.x); e[RB0_LD] = elu_fast(a.y); e[2 * RB0_LD] = elu_fast(a.z); e[3 * RB0_LD] = elu_fast(a.w);
            e += RB0_THREADS;
            e[0] = elu_fast(c.x); e[RB0_LD] = elu_fast(c.y); e[2 * RB0_LD] = elu_fast(c.z); e[3 * RB0_LD] = elu_fast(c.w);
        }
        if (tid < 64) {  // the last two columns (TILE, TILE + 1): one (column, channel) per thread of the first two warps
            const int col = RB0_TILE + (tid >> 5), ch = tid & 31;
            float acc = sW[RB0_B0 + ch];
#pragma unroll
            for (int j = 0; j < 7; ++j) acc = fmaf(sW[RB0_W0 + j * 32 + ch], sA[col + j], acc);
            sE0[ch * RB0_LD + col] = elu_fast(acc);
        }
    }
    __syncthreads();
    // reflect padding of ELU(x0) for the k3 conv: sample -1 mirrors sample 1, sample T mirrors sample T - 2
    if (tid < 32) {
        if (t0 == 0) sE0[tid * RB0_LD] = sE0[tid * RB0_LD + 2];
        const int cT = T - t0 + 1;
        if (cT >= 2 && cT <= RB0_TILE + 1) sE0[tid * RB0_LD + cT] = sE0[tid * RB0_LD + cT - 2];
    }
    __syncthreads();

    const int half = tid >> 7;         // warp-uniform channel half
    const int pos0 = (tid & 127) * 4;  // four consecutive samples t0 + pos0 .. + 3
    // ---- stage B: h1 = conv_k3(ELU(x0)), channels half*8 .. +7 ----
    {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float bb = sW[RB0_B1 + half * 8 + i];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q][i] = bb;
        }
#pragma unroll 2
        for (int c = 0; c < 32; ++c) {
            const float4 ea = *reinterpret_cast<const float4*>(sE0 + c * RB0_LD + pos0);
            const float2 eb = *reinterpret_cast<const float2*>(sE0 + c * RB0_LD + pos0 + 4);
            const float e[6] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y};
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float4 w0 = *reinterpret_cast<const float4*>(sW + RB0_W1 + (j * 32 + c) * 16 + half * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(sW + RB0_W1 + (j * 32 + c) * 16 + half * 8 + 4);
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[q][i] = fmaf(w[i], e[q + j], acc[q][i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(sH1 + (half * 8 + i) * RB0_TILE + pos0) =
                make_float4(elu_fast(acc[0][i]), elu_fast(acc[1][i]), elu_fast(acc[2][i]), elu_fast(acc[3][i]));
    }
    __syncthreads();

    // ---- stage C: y = conv1x1(ELU(h1)) + composed shortcut k7 of the raw audio ----
    // Two passes over POSITIONS (256 each); a thread owns 4 positions x 8 channels and adjacent lanes own the two
    // 8-channel halves of the same rows, so that the 16-byte plane stores of a lane pair fill whole 32-byte sectors
    // (channel passes left every sector half empty: ncu, 16 of 32 bytes per sector used).
    const int cgrp = tid & 1;
    const int ch0 = half * 16 + cgrp * 8;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int pos0 = pass * 256 + ((tid & 127) >> 1) * 4;
        float a10[10];  // samples t - 3 .. t + 6 of the first position = sA[pos0 + 1 ..]
#pragma unroll
        for (int j = 0; j < 10; ++j) a10[j] = sA[pos0 + 1 + j];
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float bb = sW[RB0_B2 + ch0 + i];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q][i] = bb;
        }
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const float4 hv = *reinterpret_cast<const float4*>(sH1 + k * RB0_TILE + pos0);
            const float hq[4] = {hv.x, hv.y, hv.z, hv.w};
            const float4 w0 = *reinterpret_cast<const float4*>(sW + RB0_W2 + k * 32 + ch0);
            const float4 w1 = *reinterpret_cast<const float4*>(sW + RB0_W2 + k * 32 + ch0 + 4);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[q][i] = fmaf(w[i], hq[q], acc[q][i]);
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float4 w0 = *reinterpret_cast<const float4*>(sW + RB0_WSC + j * 32 + ch0);
            const float4 w1 = *reinterpret_cast<const float4*>(sW + RB0_WSC + j * 32 + ch0 + 4);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[q][i] = fmaf(w[i], a10[q + j], acc[q][i]);
        }
        // output: ELU(y) planes, row b*Py + left + t plus the reflect halo rows of the strided conv's layout
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = t0 + pos0 + q;
            if (t < T) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float ea = elu_fast(acc[q][2 * i]), eb = elu_fast(acc[q][2 * i + 1]);
                    const __half2 hh = __floats2half2_rn(ea, eb);
                    const float2 f = __half22float2(hh);
                    const __half2 ll = __floats2half2_rn(ea - f.x, eb - f.y);
                    hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
                    lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                const long long base = (long long)b * Py + left;
                const long long r0 = base + t;
                const long long r1 = (t >= 1 && t <= left) ? base - t : -1;
                const long long r2 = (t <= T - 2 && t >= T - 1 - hr) ? base + 2 * (T - 1) - t : -1;
                auto put_row = [&](long long row) {
                    const long long off = row * 32 + ch0;
                    *reinterpret_cast<uint4*>(ye_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(ye_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    if (y_f32) {
                        *reinterpret_cast<float4*>(y_f32 + off) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
                        *reinterpret_cast<float4*>(y_f32 + off + 4) = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
                    }
                };
                put_row(r0);
                if (r1 >= 0) put_row(r1);
                if (r2 >= 0) put_row(r2);
            }
        }
    }
}

// SLSTM skip connection y + x (reference encoder/modules/lstm.py:38; y and x in time-major rows) fused with the ELU in front of the last
// encoder conv: writes fp32 rows [B*L, D] (tap) and the split planes of ELU(y + x) in the reflect-padded layout
// of the k7 conv (clip pitch L + 6, data at offset 3).
__global__ void lstm_skip_elu_pad_kernel(const float* __restrict__ y, const float* __restrict__ x,
                                         float* __restrict__ out_f32, __half* __restrict__ elu_hi,
                                         __half* __restrict__ elu_lo, int B, int Lmax, int D,
                                         const int* __restrict__ len_tab) {
    const int P = Lmax + 6;  // clip pitch of the padded rows; a ragged batch keeps ONE pitch, clip b is L = len_tab[b] long
    const int d8 = D / 8;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * P * d8) return;
    int c8 = (int)(gid % d8);
    long long row = gid / d8;
    int b = (int)(row / P), p = (int)(row - (long long)b * P);
    const int L = len_tab ? len_tab[b] : Lmax;
    float v[8], e[8];
    if (p >= L + 6) {  // past this clip's own padded rows (ragged batch only): finite filler, never read by a valid frame
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = 0.f;
        split_store8(elu_hi, elu_lo, row * D + c8 * 8, e);
        return;
    }
    int t = p - 3;
    const bool interior = t >= 0 && t < L;
    if (t < 0) t = -t;
    if (t >= L) t = 2 * (L - 1) - t;
    const long long src = ((long long)t * B + b) * D + c8 * 8;   // y, x: time-major rows [t*B + b]
    const long long dst = ((long long)b * Lmax + t) * D + c8 * 8;   // fp32 copy: batch-major rows
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

int resblock0_pack_floats() { return RB0_PACK; }

void launch_resblock0_fused(const float* wav, const float* pack, __half* ye_hi, __half* ye_lo, float* y_f32, int B, int T,
                            int Py, int left, int hr, cudaStream_t s) {
    if (B <= 0) return;
    if (T < 8) throw Error(4, "resblock0_fused: clip too short");
    static PerDevice<bool> attr_dev;
    bool& attr = attr_dev.get();
    if (!attr) {
        WT_CUDA(cudaFuncSetAttribute(resblock0_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB0_SMEM));
        attr = true;
    }
    dim3 grid((unsigned)((T + RB0_TILE - 1) / RB0_TILE), (unsigned)B);
    resblock0_fused_kernel<<<grid, RB0_THREADS, RB0_SMEM, s>>>(wav, pack, ye_hi, ye_lo, y_f32, T, Py, left, hr);
    WT_CUDA(cudaGetLastError());
}

void launch_lstm_skip_elu_pad(const float* y, const float* x, float* out_f32, __half* elu_hi, __half* elu_lo, int B,
                              int L, int D, cudaStream_t s, const int* len_tab) {
    long long n = (long long)B * (L + 6) * (D / 8);
    lstm_skip_elu_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y, x, out_f32, elu_hi, elu_lo, B, L, D, len_tab);
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

namespace {

// Transposed conv as a GEMM + this gather (reference encoder/modules/conv.py:232-253, SConvTranspose1d non-causal):
// g[b, t, j*Cout + co] = sum_ci W[ci, co, j] x[b, t, ci] (one GEMM, N = k*Cout, k = 2*stride), then
// y[b, m, co] = bias[co] + g[b, n/s, n%s] + g[b, n/s - 1, n%s + s] with n = m + left, for m in [0, L*stride): the full
// transposed conv trimmed by left = ceil((k - s) / 2) in front and (k - s) / 2 at the end.
__global__ void convtr_gather_kernel(const float* __restrict__ g, const float* __restrict__ bias, float* __restrict__ y,
                                     int L, int Cout, int stride, int left, long long total) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int co = (int)(gid % Cout);
    const long long bm = gid / Cout;
    const long long Tout = (long long)L * stride;
    const long long b = bm / Tout;
    const int m = (int)(bm - b * Tout);
    const int n = m + left;
    const int t0 = n / stride, j0 = n - t0 * stride;
    const long long ldg = 2LL * stride * Cout;
    float v = bias[co];
    if (t0 < L) v += g[(b * L + t0) * ldg + (long long)j0 * Cout + co];
    if (t0 >= 1 && t0 - 1 < L) v += g[(b * L + t0 - 1) * ldg + (long long)(j0 + stride) * Cout + co];
    y[gid] = v;
}
}  // namespace

void launch_convtr_gather(const float* g, const float* bias, float* y, int B, int L, int Cout, int stride, int left,
                          cudaStream_t s) {
    const long long total = (long long)B * L * stride * Cout;
    if (total <= 0) return;
    convtr_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(g, bias, y, L, Cout, stride, left, total);
    WT_CUDA(cudaGetLastError());
}

}  // namespace wt
