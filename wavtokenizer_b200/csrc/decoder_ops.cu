// Memory-bound decoder kernels: GroupNorm(+swish), LayerNorm / AdaLayerNorm, depthwise conv fused
// with AdaLayerNorm, single-head attention, spectral (mag/phase -> re/im) and overlap-add.
//
// Activations are channels-last rows. Clip b owns rows [b*Lp, b*Lp + L); with Lp > L the Lp - L rows
// after each clip are the zero halo of the padded row space the tcgen05 tap-GEMM reads (gemm_tc.cu).
// Every producer can write fp32 rows or the split-fp16 operand planes (hi = fp16(x), lo = fp16(x - hi))
// of the next GEMM directly, so no separate conversion pass touches HBM.
#include <cuda_fp16.h>
#include <algorithm>

#include <cstdlib>

#include "common.cuh"

namespace wt {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NW>
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += red[i];
    return t;
}

__device__ __forceinline__ void put(const RowOut& o, long long idx, float v) {
    if (o.f32) o.f32[idx] = v;
    if (o.hi) {
        __half h = __float2half_rn(v);
        o.hi[idx] = h;
        if (o.lo) o.lo[idx] = __float2half_rn(v - __half2float(h));
    }
}

// 8 consecutive channels of one row -> fp32 and/or split planes (16-byte stores).
__device__ __forceinline__ void put8(const RowOut& o, long long idx, const float (&v)[8]) {
    if (o.f32) {
        *reinterpret_cast<float4*>(o.f32 + idx) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o.f32 + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (o.hi) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // packed conversions (ALU pipe) rather than scalar F2F (XU pipe)
            const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            const float2 f = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
            h[i] = *reinterpret_cast<const uint32_t*>(&hh);
            l[i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(o.hi + idx) = make_uint4(h[0], h[1], h[2], h[3]);
        if (o.lo) *reinterpret_cast<uint4*>(o.lo + idx) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// Normalize = GroupNorm(32, C, eps=1e-6, affine) (+ x*sigmoid(x)) (reference decoder/models.py:10-16,
// 58-78, 107-110). Statistics span all L frames of a clip. One block per (clip, slab of GN_GPB = 2 groups = 48
// channels): 12 lanes x float4 read 192-byte row slabs (six full sectors), 32 row phases (384 threads, four
// independent loads in flight each), 16 x B blocks so that the grid spreads evenly over the SMs. Pass 1
// accumulates sum / sum-of-squares (combined in fp64), pass 2 re-reads the slab (L2-resident), normalises and
// stores fp32 rows or split-fp16 planes with 8-byte vectors. Halo rows of the padded row space are written as zeros.
// RES > 0: L <= RES * GN_PH, the thread's <= RES rows stay in registers between the two passes (all RES loads in
// flight at once, no second read); RES == 0: any L, pass 2 re-reads the slab.
constexpr int GN_GPB = 2, GN_LANES = GN_GPB * 6, GN_PH = 32, GN_THREADS = GN_LANES * GN_PH, GN_RES = 8;
template <int RES, int MINB = 3>  // MINB = 3: 56 registers, three blocks (36 warps) per SM instead of two at 66 registers
__global__ void __launch_bounds__(GN_THREADS, MINB) groupnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bsh, RowOut out, int L, int Lp,
                                                               int C, float eps, int swish, Ragged rg) {
    __shared__ double part[GN_THREADS / 32][4];  // per warp: (sum, sum of squares) of group 0 and of group 1
    const int b = blockIdx.y, slab = blockIdx.x;
    if (rg.len) L = rg.len[b];  // ragged: statistics over the clip's own frames, zero rows from there to the pitch
    const int lane = threadIdx.x % GN_LANES, ph = threadIdx.x / GN_LANES;
    const int c = slab * (GN_GPB * 24) + lane * 4;
    const long long base = (long long)b * Lp * C + c;
    float s = 0.f, q = 0.f;
    float4 keep[RES > 0 ? RES : 1];
    if (RES > 0) {
#pragma unroll
        for (int i = 0; i < RES; ++i) {
            const int t = ph + i * GN_PH;
            keep[i] = t < L ? *reinterpret_cast<const float4*>(x + base + (long long)t * C) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < RES; ++i) {  // zero rows add nothing to either sum
            const float4 v = keep[i];
            s += (v.x + v.y) + (v.z + v.w);
            q = fmaf(v.x, v.x, q); q = fmaf(v.y, v.y, q); q = fmaf(v.z, v.z, q); q = fmaf(v.w, v.w, q);
        }
    } else {
#pragma unroll 4
        for (int t = ph; t < L; t += GN_PH) {
            const float4 v = *reinterpret_cast<const float4*>(x + base + (long long)t * C);
            s += (v.x + v.y) + (v.z + v.w);
            q = fmaf(v.x, v.x, q); q = fmaf(v.y, v.y, q); q = fmaf(v.z, v.z, q); q = fmaf(v.w, v.w, q);
        }
    }
    // group sums: lanes 6g..6g+5 of every row phase belong to group g of the slab. Warp-level sums of the two groups in
    // fp64 (shuffles), one partial per warp in shared memory, then every thread adds the 12 partials of ITS group: one
    // barrier and no serial tail (the column-sum loops of the first version cost about as much as the loads)
    {
        const int wl = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const bool g1 = lane >= 6;
        double a0 = g1 ? 0.0 : (double)s, b0 = g1 ? 0.0 : (double)q, a1 = g1 ? (double)s : 0.0, b1 = g1 ? (double)q : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, off);
            b0 += __shfl_xor_sync(0xffffffffu, b0, off);
            a1 += __shfl_xor_sync(0xffffffffu, a1, off);
            b1 += __shfl_xor_sync(0xffffffffu, b1, off);
        }
        if (wl == 0) { part[wid][0] = a0; part[wid][1] = b0; part[wid][2] = a1; part[wid][3] = b1; }
    }
    __syncthreads();
    float mean, rstd;
    {
        const int g = lane >= 6 ? 1 : 0;
        double ds = 0, dq = 0;
#pragma unroll
        for (int w_ = 0; w_ < GN_THREADS / 32; ++w_) { ds += part[w_][2 * g]; dq += part[w_][2 * g + 1]; }
        // 1 / n and 1 / sqrt(var + eps) from the fp32 special-function unit + one Newton step in fp64 (relative error
        // ~1e-14: the fp32 results are the correctly rounded ones) instead of three fp64 divisions and a square root,
        // which every thread of the block would execute (~100 issue slots of ~800)
        const double n = (double)L * 24.0;
        const double r0 = (double)__frcp_rn((float)n);
        const double rn = r0 * (2.0 - n * r0);
        const double m_ = ds * rn;
        double var = dq * rn - m_ * m_;
        if (var < 0) var = 0;
        const double ve = var + (double)eps;
        const double y0 = (double)rsqrtf((float)ve);
        mean = (float)m_;
        rstd = (float)(y0 * (1.5 - 0.5 * ve * y0 * y0));
    }
    // y = (v - mean) * rstd * w + b as ONE packed FMA per channel pair: y = v * a + d, a = rstd * w, d = b - mean * a
    const float4 wv = *reinterpret_cast<const float4*>(w + c), bv = *reinterpret_cast<const float4*>(bsh + c);
    const float2 a01 = make_float2(rstd * wv.x, rstd * wv.y), a23 = make_float2(rstd * wv.z, rstd * wv.w);
    const float2 d01 = make_float2(fmaf(-mean, a01.x, bv.x), fmaf(-mean, a01.y, bv.y));
    const float2 d23 = make_float2(fmaf(-mean, a23.x, bv.z), fmaf(-mean, a23.y, bv.w));
    auto emit = [&](int t, const float4& v) {  // row t of the slab: normalised (+ swish) or, past L, a zero halo row
        float y[4] = {0.f, 0.f, 0.f, 0.f};
        if (t < L) {
            const float2 y01 = __ffma2_rn(make_float2(v.x, v.y), a01, d01), y23 = __ffma2_rn(make_float2(v.z, v.w), a23, d23);
            y[0] = y01.x; y[1] = y01.y; y[2] = y23.x; y[3] = y23.y;
            if (swish) {
#pragma unroll
                for (int i = 0; i < 4; ++i) y[i] = __fdividef(y[i], 1.f + __expf(-y[i]));
            }
        }
        const long long idx = base + (long long)t * C;
        if (out.f32) *reinterpret_cast<float4*>(out.f32 + idx) = make_float4(y[0], y[1], y[2], y[3]);
        if (out.hi) {
            uint32_t h[2], l[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const __half2 hh = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
                const float2 f = __half22float2(hh);
                const __half2 ll = __floats2half2_rn(y[2 * i] - f.x, y[2 * i + 1] - f.y);
                h[i] = *reinterpret_cast<const uint32_t*>(&hh);
                l[i] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            *reinterpret_cast<uint2*>(out.hi + idx) = make_uint2(h[0], h[1]);
            if (out.lo) *reinterpret_cast<uint2*>(out.lo + idx) = make_uint2(l[0], l[1]);
        }
    };
    if (RES > 0) {
#pragma unroll
        for (int i = 0; i < RES; ++i) {
            const int t = ph + i * GN_PH;
            if (t < Lp) emit(t, keep[i]);
        }
        const int t = ph + RES * GN_PH;  // halo rows past RES * GN_PH (Lp <= L + GN_PH)
        if (t < Lp) emit(t, make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
#pragma unroll 4
        for (int t = ph; t < Lp; t += GN_PH)
            emit(t, t < L ? *reinterpret_cast<const float4*>(x + base + (long long)t * C) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

// layer_norm over C = 768 (eps 1e-6) then * w + b, one warp per row; lane owns 3 x 8 consecutive channels
// (c = seg*256 + lane*8 + k) so that loads are 32-byte and stores 16-byte vectors. With w = scale[id],
// b = shift[id] this is AdaLayerNorm (reference decoder/modules.py:81-86); with the affine parameters it is
// the final nn.LayerNorm (decoder/models.py:195, 234).
__device__ __forceinline__ void ln_finish(float (&v)[3][8], const float* __restrict__ w, const float* __restrict__ b,
                                          const RowOut& out, long long row, int lane, float eps) {
    constexpr int C = 768;
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[g][k];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = v[g][k] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const int c = g * 256 + lane * 8;
        float wv[8], bv[8], y[8];
        load8(w + c, wv);
        load8(b + c, bv);
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = (v[g][k] - mean) * rstd * wv[k] + bv[k];
        put8(out, row * C + c, y);
    }
}

__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, RowOut out, long long M,
                                                        float eps) {
    const int lane = threadIdx.x & 31;
    long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    float v[3][8];
#pragma unroll
    for (int g = 0; g < 3; ++g) load8(x + row * 768 + g * 256 + lane * 8, v[g]);
    ln_finish(v, w, b, out, row, lane, eps);
}

// ConvNeXt front half: depthwise Conv1d(k=7, pad 3 zeros, groups=C) -> AdaLayerNorm
// (reference decoder/modules.py:30-33, 49-53). One block per (clip, run of DW_TT frames); a thread owns 4
// consecutive channels for the whole run: its 28 taps live in registers, the DW_TT + 6 input rows are read once
// each as coalesced float4 (all loads issued up front), the conv slides over them in registers, and the
// LayerNorm statistics of the DW_TT frames are reduced together (a 16-value butterfly: 16 shuffles instead of
// 80 per pass, then one shared-memory exchange between the 6 warps). Two-pass variance (mean, then centred sum
// of squares) as in the per-row LayerNorm kernel. HBM traffic: one read of x, one write of the operand planes.
constexpr int DW_TT = 16, DW_THREADS = 192;

// Sums v[i] over the 32 lanes for 16 values at once. On return v[0] of lane l holds the total of value (l >> 1).
__device__ __forceinline__ void warp_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int o = 16, n = 16; n > 1; o >>= 1, n >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__global__ void __launch_bounds__(DW_THREADS, 2) dwconv_ln_kernel(const float* __restrict__ x, const float* __restrict__ dwT,
                                                                  const float* __restrict__ db, const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, RowOut out, int L, int Lp,
                                                                  float eps, Ragged rg) {
    constexpr int C = 768, NW = DW_THREADS / 32;
    __shared__ __align__(16) float red[2][DW_TT][8];  // [pass][frame][warp] partial sums (6 used)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, t0 = blockIdx.x * DW_TT;
    if (rg.len) L = rg.len[b];  // ragged: rows past the clip's own end read as the conv's zero padding and are written as zeros
    const int c = threadIdx.x * 4;
    const float* xb = x + (long long)b * Lp * C + c;
    float4 xr[DW_TT + 6];
#pragma unroll
    for (int r = 0; r < DW_TT + 6; ++r) {
        const int t = t0 - 3 + r;
        xr[r] = (t >= 0 && t < L) ? *reinterpret_cast<const float4*>(xb + (long long)t * C) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 w[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) w[j] = *reinterpret_cast<const float4*>(dwT + j * C + c);
    const float4 bias = *reinterpret_cast<const float4*>(db + c);
    float4 v[DW_TT];
    float part[16];
#pragma unroll
    for (int i = 0; i < DW_TT; ++i) {
        float4 a = bias;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            a.x = fmaf(w[j].x, xr[i + j].x, a.x);
            a.y = fmaf(w[j].y, xr[i + j].y, a.y);
            a.z = fmaf(w[j].z, xr[i + j].z, a.z);
            a.w = fmaf(w[j].w, xr[i + j].w, a.w);
        }
        v[i] = a;
        part[i] = (a.x + a.y) + (a.z + a.w);
    }
    warp_sum16(part, lane);
    if ((lane & 1) == 0) red[0][lane >> 1][warp] = part[0];
    __syncthreads();
    float mean[DW_TT];
#pragma unroll
    for (int i = 0; i < DW_TT; ++i) {
        const float4 p0 = *reinterpret_cast<const float4*>(&red[0][i][0]);
        const float2 p1 = *reinterpret_cast<const float2*>(&red[0][i][4]);
        static_assert(NW == 6, "six warps cover the 768 channels");
        mean[i] = (((p0.x + p0.y) + (p0.z + p0.w)) + (p1.x + p1.y)) * (1.f / (float)C);
        const float dx = v[i].x - mean[i], dy = v[i].y - mean[i], dz = v[i].z - mean[i], dw = v[i].w - mean[i];
        part[i] = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
    }
    warp_sum16(part, lane);
    if ((lane & 1) == 0) red[1][lane >> 1][warp] = part[0];
    __syncthreads();
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
#pragma unroll
    for (int i = 0; i < DW_TT; ++i) {
        const int t = t0 + i;
        if (t >= Lp) break;
        float y[4] = {0.f, 0.f, 0.f, 0.f};
        if (t < L) {
            const float4 p0 = *reinterpret_cast<const float4*>(&red[1][i][0]);
            const float2 p1 = *reinterpret_cast<const float2*>(&red[1][i][4]);
            const float var = (((p0.x + p0.y) + (p0.z + p0.w)) + (p1.x + p1.y)) * (1.f / (float)C);
            const float rstd = rsqrtf(var + eps);
            y[0] = (v[i].x - mean[i]) * rstd * sc.x + sh.x;
            y[1] = (v[i].y - mean[i]) * rstd * sc.y + sh.y;
            y[2] = (v[i].z - mean[i]) * rstd * sc.z + sh.z;
            y[3] = (v[i].w - mean[i]) * rstd * sc.w + sh.w;
        }  // else: halo row of the padded row space, written as zeros
        const long long idx = ((long long)b * Lp + t) * C + c;
        if (out.f32) *reinterpret_cast<float4*>(out.f32 + idx) = make_float4(y[0], y[1], y[2], y[3]);
        if (out.hi) {
            const __half2 h01 = __floats2half2_rn(y[0], y[1]), h23 = __floats2half2_rn(y[2], y[3]);
            *reinterpret_cast<uint2*>(out.hi + idx) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            if (out.lo) {
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(y[0] - f01.x, y[1] - f01.y);
                const __half2 l23 = __floats2half2_rn(y[2] - f23.x, y[3] - f23.y);
                *reinterpret_cast<uint2*>(out.lo + idx) =
                    make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            }
        }
    }
}

// Same operator for the single-plane consumer (plan 2: the ConvNeXt GEMM-1 reads ONE fp16 plane, so the output carries 11
// bits): runs of 8 frames, 14 input rows per thread, mean and E[x^2] reduced TOGETHER (one 16-value butterfly, one barrier;
// var = E[x^2] - mean^2 in fp32 is exact to ~1e-6 relative here, far below the fp16 rounding of the result), <= 112
// registers so that three blocks share an SM (the 16-frame kernel above holds 168 registers: two blocks, 18 % occupancy,
// and its two statistics passes cost a second barrier).
constexpr int DW8_TT = 8;
__global__ void __launch_bounds__(DW_THREADS, 3) dwconv_ln_hi_kernel(const float* __restrict__ x, const float* __restrict__ dwT,
                                                                     const float* __restrict__ db, const float* __restrict__ scale,
                                                                     const float* __restrict__ shift, __half* __restrict__ out_hi,
                                                                     int L, int Lp, float eps, Ragged rg) {
    constexpr int C = 768, NW = DW_THREADS / 32;
    static_assert(NW == 6, "six warps cover the 768 channels");
    __shared__ __align__(16) float red[2 * DW8_TT][8];  // [sum of frame i | sum of squares of frame i][warp]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, t0 = blockIdx.x * DW8_TT;
    if (rg.len) L = rg.len[b];
    const int c = threadIdx.x * 4;
    const float* xb = x + (long long)b * Lp * C + c;
    float4 xr[DW8_TT + 6];
#pragma unroll
    for (int r = 0; r < DW8_TT + 6; ++r) {
        const int t = t0 - 3 + r;
        xr[r] = (t >= 0 && t < L) ? *reinterpret_cast<const float4*>(xb + (long long)t * C) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 bias = *reinterpret_cast<const float4*>(db + c);
    float4 v[DW8_TT];
#pragma unroll
    for (int i = 0; i < DW8_TT; ++i) v[i] = bias;
#pragma unroll
    for (int j = 0; j < 7; ++j) {  // tap-major: one weight vector live at a time
        const float4 w = *reinterpret_cast<const float4*>(dwT + j * C + c);
#pragma unroll
        for (int i = 0; i < DW8_TT; ++i) {
            v[i].x = fmaf(w.x, xr[i + j].x, v[i].x);
            v[i].y = fmaf(w.y, xr[i + j].y, v[i].y);
            v[i].z = fmaf(w.z, xr[i + j].z, v[i].z);
            v[i].w = fmaf(w.w, xr[i + j].w, v[i].w);
        }
    }
    float part[16];
#pragma unroll
    for (int i = 0; i < DW8_TT; ++i) {
        part[i] = (v[i].x + v[i].y) + (v[i].z + v[i].w);
        part[DW8_TT + i] = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, v[i].w * v[i].w)));
    }
    warp_sum16(part, lane);
    if ((lane & 1) == 0) red[lane >> 1][warp] = part[0];
    __syncthreads();
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
#pragma unroll
    for (int i = 0; i < DW8_TT; ++i) {
        const int t = t0 + i;
        if (t >= Lp) break;
        float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
        if (t < L) {
            const float4 a0 = *reinterpret_cast<const float4*>(&red[i][0]);
            const float2 a1 = *reinterpret_cast<const float2*>(&red[i][4]);
            const float4 q0 = *reinterpret_cast<const float4*>(&red[DW8_TT + i][0]);
            const float2 q1 = *reinterpret_cast<const float2*>(&red[DW8_TT + i][4]);
            const float mean = (((a0.x + a0.y) + (a0.z + a0.w)) + (a1.x + a1.y)) * (1.f / (float)C);
            const float ex2 = (((q0.x + q0.y) + (q0.z + q0.w)) + (q1.x + q1.y)) * (1.f / (float)C);
            const float rstd = rsqrtf(fmaxf(ex2 - mean * mean, 0.f) + eps);
            y0 = (v[i].x - mean) * rstd * sc.x + sh.x;
            y1 = (v[i].y - mean) * rstd * sc.y + sh.y;
            y2 = (v[i].z - mean) * rstd * sc.z + sh.z;
            y3 = (v[i].w - mean) * rstd * sc.w + sh.w;
        }  // else: halo row of the padded row space, written as zeros
        const __half2 h01 = __floats2half2_rn(y0, y1), h23 = __floats2half2_rn(y2, y3);
        *reinterpret_cast<uint2*>(out_hi + ((long long)b * Lp + t) * C + c) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    }
}

// The same arithmetic with half the issue slots: the kernel above issues ~1000
// instructions per thread and run (31 per output value) and is bound by exactly that (2.1 IPC at 18 warps per SM, 3.2 TB/s).
// Here the 224 conv FMAs and the normalisation run as packed fp32 pairs (FFMA2 / FADD2 / FMUL2: two channels per
// instruction), and the cross-warp sums of the 16 statistics are finished by 16 lanes (one statistic each) and handed
// round by shuffles instead of every thread adding up all 16 x 6 partials.
template <bool EDGE>  // EDGE = false: all 14 input rows and all 8 output rows of the run lie inside the clip (no masks)
__device__ __forceinline__ void dwconv_ln_hi2_run(const float* __restrict__ xb, const float* __restrict__ dwT,
                                                  const float* __restrict__ db, const float* __restrict__ scale,
                                                  const float* __restrict__ shift, __half* __restrict__ ob, int t0, int L, int Lp,
                                                  float eps, float (*red)[8]) {
    constexpr int C = 768;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = threadIdx.x * 4;
    float4 xr[DW8_TT + 6];
#pragma unroll
    for (int r = 0; r < DW8_TT + 6; ++r) {
        const int t = t0 - 3 + r;
        xr[r] = (!EDGE || (t >= 0 && t < L)) ? *reinterpret_cast<const float4*>(xb + (long long)t * C + c)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 bias = *reinterpret_cast<const float4*>(db + c);
    float2 va[DW8_TT], vb[DW8_TT];  // channels (c, c+1) and (c+2, c+3)
#pragma unroll
    for (int i = 0; i < DW8_TT; ++i) { va[i] = make_float2(bias.x, bias.y); vb[i] = make_float2(bias.z, bias.w); }
#pragma unroll
    for (int j = 0; j < 7; ++j) {  // tap-major: one weight vector live at a time
        const float4 w = *reinterpret_cast<const float4*>(dwT + j * C + c);
        const float2 wa = make_float2(w.x, w.y), wb = make_float2(w.z, w.w);
#pragma unroll
        for (int i = 0; i < DW8_TT; ++i) {
            va[i] = __ffma2_rn(wa, make_float2(xr[i + j].x, xr[i + j].y), va[i]);
            vb[i] = __ffma2_rn(wb, make_float2(xr[i + j].z, xr[i + j].w), vb[i]);
        }
    }
    float part[16];
#pragma unroll
    for (int i = 0; i < DW8_TT; ++i) {
        part[i] = (va[i].x + va[i].y) + (vb[i].x + vb[i].y);
        part[DW8_TT + i] = fmaf(va[i].x, va[i].x, fmaf(va[i].y, va[i].y, fmaf(vb[i].x, vb[i].x, vb[i].y * vb[i].y)));
    }
    warp_sum16(part, lane);
    if ((lane & 1) == 0) red[lane >> 1][warp] = part[0];
    __syncthreads();
    // lane l < 16 finishes statistic l (same order of additions as above); lanes 0..7 then hold mean and rstd of frame l
    float tot = 0.f;
    if (lane < 2 * DW8_TT) {
        const float4 a0 = *reinterpret_cast<const float4*>(&red[lane][0]);
        const float2 a1 = *reinterpret_cast<const float2*>(&red[lane][4]);
        tot = (((a0.x + a0.y) + (a0.z + a0.w)) + (a1.x + a1.y)) * (1.f / (float)C);
    }
    const float ex2 = __shfl_down_sync(0xffffffffu, tot, DW8_TT);
    const float rstd_l = rsqrtf(fmaxf(ex2 - tot * tot, 0.f) + eps);
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
    const float2 sca = make_float2(sc.x, sc.y), scb = make_float2(sc.z, sc.w);
    const float2 sha = make_float2(sh.x, sh.y), shb = make_float2(sh.z, sh.w);
#pragma unroll
    for (int i = 0; i < DW8_TT; ++i) {
        const int t = t0 + i;
        if (EDGE && t >= Lp) break;
        const float nmean = -__shfl_sync(0xffffffffu, tot, i), rstd = __shfl_sync(0xffffffffu, rstd_l, i);
        float2 ya = make_float2(0.f, 0.f), yb = ya;
        if (!EDGE || t < L) {
            const float2 nm = make_float2(nmean, nmean), rs = make_float2(rstd, rstd);
            ya = __ffma2_rn(__fmul2_rn(__fadd2_rn(va[i], nm), rs), sca, sha);
            yb = __ffma2_rn(__fmul2_rn(__fadd2_rn(vb[i], nm), rs), scb, shb);
        }  // else: halo row of the padded row space, written as zeros
        const __half2 h01 = __floats2half2_rn(ya.x, ya.y), h23 = __floats2half2_rn(yb.x, yb.y);
        *reinterpret_cast<uint2*>(ob + (long long)t * C + c) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    }
}

__global__ void __launch_bounds__(DW_THREADS, 3) dwconv_ln_hi2_kernel(const float* __restrict__ x, const float* __restrict__ dwT,
                                                                      const float* __restrict__ db, const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, __half* __restrict__ out_hi,
                                                                      int L, int Lp, float eps, Ragged rg) {
    static_assert(DW_THREADS / 32 == 6, "six warps cover the 768 channels");
    __shared__ __align__(16) float red[2 * DW8_TT][8];  // [sum of frame i | sum of squares of frame i][warp]
    const int b = blockIdx.y, t0 = blockIdx.x * DW8_TT;
    if (rg.len) L = rg.len[b];
    const float* xb = x + (long long)b * Lp * 768;
    __half* ob = out_hi + (long long)b * Lp * 768;
    if (t0 >= 3 && t0 + DW8_TT + 3 <= L) dwconv_ln_hi2_run<false>(xb, dwT, db, scale, shift, ob, t0, L, Lp, eps, red);
    else dwconv_ln_hi2_run<true>(xb, dwT, db, scale, shift, ob, t0, L, Lp, eps, red);
}

// AttnBlock core (reference decoder/models.py:115-123): softmax(q k^T * C^-0.5) v, one head of width C
// over the L frames of a clip. One warp per query row, 8 queries per block; scores live in shared memory.
template <int PER>
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, RowOut out, int L, int Lp,
                                                        float scale) {
    constexpr int C = PER * 32;
    extern __shared__ float sm[];  // [8][L] scores
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int qi = blockIdx.x * 8 + warp;
    const float* base = qkv + (long long)b * Lp * 3 * C;
    float* sc = sm + warp * L;
    if (qi >= Lp) return;
    const long long orow = ((long long)b * Lp + qi) * C;
    if (qi >= L) {
#pragma unroll
        for (int i = 0; i < PER; ++i) put(out, orow + lane + 32 * i, 0.f);
        return;
    }
    float q[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) q[i] = base[(long long)qi * 3 * C + lane + 32 * i];
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) {
        const float* kr = base + (long long)j * 3 * C + C;
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) d = fmaf(q[i], kr[lane + 32 * i], d);
        d = warp_sum(d) * scale;
        if (lane == 0) sc[j] = d;
        mx = fmaxf(mx, d);
    }
    __syncwarp();
    float den = 0.f;
    for (int j = lane; j < L; j += 32) {
        float e = expf(sc[j] - mx);
        sc[j] = e;
        den += e;
    }
    den = warp_sum(den);
    __syncwarp();
    const float inv = 1.f / den;
    float o[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) o[i] = 0.f;
    for (int j = 0; j < L; ++j) {
        const float p = sc[j] * inv;
        const float* vr = base + (long long)j * 3 * C + 2 * C;
#pragma unroll
        for (int i = 0; i < PER; ++i) o[i] = fmaf(p, vr[lane + 32 * i], o[i]);
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) put(out, orow + lane + 32 * i, o[i]);
}

// ISTFTHead spectral step (reference decoder/heads.py:55-65): z = [log-mag | phase] (row pitch ldz) ->
// S = [min(exp(m), 100) cos p | min(exp(m), 100) sin p], zero-filled to ldS columns (the K of the iDFT GEMM).
// One thread per (row, bin): both outputs of a bin share one exp and one sincos; threads past the last bin zero the pad.
__global__ void __launch_bounds__(256) spectral_kernel(const float* __restrict__ z, int ldz, RowOut S, long long M, int half,
                                                       int ldS) {
    const int W = ldS - half;  // bins, then the zero pad columns [2*half, ldS)
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= M * W) return;
    const long long m = gid / W;
    const int i = (int)(gid - m * W);
    if (i < half) {
        const float mag = fminf(expf(z[m * ldz + i]), 100.f);
        float sn, cs;
        sincosf(z[m * ldz + half + i], &sn, &cs);
        put(S, m * ldS + i, mag * cs);
        put(S, m * ldS + half + i, mag * sn);
    } else {
        put(S, m * ldS + half + i, 0.f);
    }
}

// Same step, one WARP per row: no 64-bit index division per element, four bins per lane in flight (the loads of a run of
// 128 bins are issued before the exp / sincos of the first one). Same expf / sincosf, so the values are identical.
__global__ void __launch_bounds__(256) spectral_rows_kernel(const float* __restrict__ z, int ldz, RowOut S, long long M,
                                                            int half, int ldS) {
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= M) return;
    const float* zr = z + m * ldz;
    const long long o = m * ldS;
    for (int i0 = 0; i0 < half; i0 += 128) {
        float mg[4], ph[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + lane;
            mg[u] = i < half ? zr[i] : 0.f;
            ph[u] = i < half ? zr[half + i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + lane;
            if (i < half) {
                const float mag = fminf(expf(mg[u]), 100.f);
                float sn, cs;
                sincosf(ph[u], &sn, &cs);
                put(S, o + i, mag * cs);
                put(S, o + half + i, mag * sn);
            }
        }
    }
    for (int i = 2 * half + lane; i < ldS; i += 32) put(S, o + i, 0.f);
}

// ISTFT "same" overlap-add + envelope normalisation (reference decoder/spectral_ops.py:58-73).
// frames already carry the window (folded into the iDFT basis). Each output sample sums the <= n_fft/hop
// frames that cover it and divides by the matching sum of squared window samples.
__global__ void overlap_add_kernel(const float* __restrict__ frames, const float* __restrict__ wsq,
                                   float* __restrict__ audio, int L, int Lp, int n_fft, int hop, int pad, Ragged rg) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const long long out0 = rg.off ? rg.off[b] * hop : (long long)b * L * hop;  // ragged: outputs packed back to back
    if (rg.len) L = rg.len[b];
    const int len = L * hop;
    if (n >= len) return;
    const int p = n + pad;
    int t_hi = p / hop;
    int t_lo = (p - n_fft + hop) / hop;  // ceil((p - n_fft + 1) / hop) for p >= n_fft - 1; clipped below
    if (p - n_fft + 1 <= 0) t_lo = 0;
    if (t_hi > L - 1) t_hi = L - 1;
    float acc = 0.f, env = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) {
        int k = p - t * hop;
        acc += frames[((long long)b * Lp + t) * n_fft + k];
        env += wsq[k];
    }
    audio[out0 + n] = acc / env;
}

// Four consecutive samples per thread (hop, n_fft and the padding are multiples of 4: the four samples share their frame
// range and sit in one 16-byte group of every frame and of the envelope table). Same summation order as above.
__global__ void __launch_bounds__(256) overlap_add4_kernel(const float* __restrict__ frames, const float* __restrict__ wsq,
                                                           float* __restrict__ audio, int L, int Lp, int n_fft, int hop, int pad,
                                                           Ragged rg) {
    const int b = blockIdx.y;
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const long long out0 = rg.off ? rg.off[b] * hop : (long long)b * L * hop;
    if (rg.len) L = rg.len[b];
    if (n >= L * hop) return;
    const int p = n + pad;
    int t_hi = p / hop;
    int t_lo = p < n_fft ? 0 : (p - n_fft + hop) / hop;
    if (t_hi > L - 1) t_hi = L - 1;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), env = acc;
    const float* fb = frames + (long long)b * Lp * n_fft;
    for (int t = t_lo; t <= t_hi; ++t) {
        const int k = p - t * hop;
        const float4 f = *reinterpret_cast<const float4*>(fb + (long long)t * n_fft + k);
        const float4 w = *reinterpret_cast<const float4*>(wsq + k);
        acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
        env.x += w.x; env.y += w.y; env.z += w.z; env.w += w.w;
    }
    *reinterpret_cast<float4*>(audio + out0 + n) = make_float4(acc.x / env.x, acc.y / env.y, acc.z / env.z, acc.w / env.w);
}

// features [B, C, L] (API layout) -> rows [B*Lp, C] (fp32 or split planes), halo rows zeroed.
__global__ void features_to_rows_kernel(const float* __restrict__ in, RowOut out, int C, int L, int Lp, Ragged rg) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const float* ib = in + (rg.off ? rg.off[b] * C : (long long)b * C * L);  // ragged: clip b is [C, len[b]] at off[b] * C
    if (rg.len) L = rg.len[b];
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < L) ? ib[(long long)c * L + t] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int t = t0 + i, c = c0 + threadIdx.x;
        if (t < Lp && c < C) put(out, ((long long)b * Lp + t) * C + c, tile[threadIdx.x][i]);
    }
}

// softmax(q k^T * C^-0.5) rows for the tensor-core attention (reference decoder/models.py:117-120):
// S [B*Lp, ldS] fp32 raw scores -> split-fp16 planes of the probabilities [B*Lp, Lpad], zero for keys >= L
// (they are the K dimension of the P.V GEMM) and for halo query rows. One warp per query row.
__global__ void __launch_bounds__(256) softmax_planes_kernel(const float* __restrict__ S, int ldS,
                                                             __half* __restrict__ p_hi, __half* __restrict__ p_lo,
                                                             int Lpad, int B, int L, int Lp, float scale, Ragged rg) {
    const int lane = threadIdx.x & 31;
    long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long long)B * Lp) return;
    const int i = (int)(row % Lp);
    if (rg.len) L = rg.len[row / Lp];  // ragged: keys and queries of clip b stop at len[b]
    __half* oh = p_hi + row * Lpad;
    __half* ol = p_lo + row * Lpad;
    if (i >= L) {
        for (int j = lane; j < Lpad; j += 32) { oh[j] = __float2half_rn(0.f); ol[j] = __float2half_rn(0.f); }
        return;
    }
    const float* sr = S + row * ldS;
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) mx = fmaxf(mx, sr[j] * scale);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float den = 0.f;
    for (int j = lane; j < L; j += 32) den += expf(sr[j] * scale - mx);
    den = warp_sum(den);
    const float inv = 1.f / den;
    for (int j = lane; j < Lpad; j += 32) {
        float p = j < L ? expf(sr[j] * scale - mx) * inv : 0.f;
        __half h = __float2half_rn(p);
        oh[j] = h;
        ol[j] = __float2half_rn(p - __half2float(h));
    }
}

// V^T planes for the P.V GEMM: qkv planes [B*Lp, 3C] (v = columns 2C..3C) -> vT [B*C, Lpad], zero for keys >= L.
__global__ void vt_planes_kernel(const __half* __restrict__ q_hi, const __half* __restrict__ q_lo,
                                 __half* __restrict__ vt_hi, __half* __restrict__ vt_lo, int L, int Lp, int C,
                                 int Lpad, Ragged rg) {
    __shared__ __half th[32][34], tl[32][34];
    const int b = blockIdx.z;
    if (rg.len) L = rg.len[b];
    const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, c = c0 + threadIdx.x;
        __half h = __float2half_rn(0.f), l = h;
        if (j < L) {
            const long long src = ((long long)b * Lp + j) * 3 * C + 2 * C + c;
            h = q_hi[src];
            l = q_lo[src];
        }
        th[r][threadIdx.x] = h;
        tl[r][threadIdx.x] = l;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, j = j0 + threadIdx.x;
        if (j < Lpad) {
            const long long dst = ((long long)b * C + c) * Lpad + j;
            vt_hi[dst] = th[threadIdx.x][r];
            vt_lo[dst] = tl[threadIdx.x][r];
        }
    }
}

// Same transpose on 64 x 64 tiles with two halves per thread on both sides (128-byte rows per warp instead of 64).
__global__ void __launch_bounds__(256) vt_planes2_kernel(const __half* __restrict__ q_hi, const __half* __restrict__ q_lo,
                                                         __half* __restrict__ vt_hi, __half* __restrict__ vt_lo, int L, int Lp,
                                                         int C, int Lpad, Ragged rg) {
    __shared__ __half th[64][66], tl[64][66];
    const int b = blockIdx.z;
    if (rg.len) L = rg.len[b];
    const int j0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    for (int r = threadIdx.y; r < 64; r += 8) {
        const int j = j0 + r, c = c0 + 2 * threadIdx.x;
        __half2 h = __floats2half2_rn(0.f, 0.f), l = h;
        if (j < L) {
            const long long src = ((long long)b * Lp + j) * 3 * C + 2 * C + c;
            h = *reinterpret_cast<const __half2*>(q_hi + src);
            l = *reinterpret_cast<const __half2*>(q_lo + src);
        }
        *reinterpret_cast<__half2*>(&th[r][2 * threadIdx.x]) = h;
        *reinterpret_cast<__half2*>(&tl[r][2 * threadIdx.x]) = l;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 64; r += 8) {
        const int c = c0 + r, j = j0 + 2 * threadIdx.x;
        if (j < Lpad) {
            const long long dst = ((long long)b * C + c) * Lpad + j;
            *reinterpret_cast<__half2*>(vt_hi + dst) = __halves2half2(th[2 * threadIdx.x][r], th[2 * threadIdx.x + 1][r]);
            *reinterpret_cast<__half2*>(vt_lo + dst) = __halves2half2(tl[2 * threadIdx.x][r], tl[2 * threadIdx.x + 1][r]);
        }
    }
}

// WT_MEM_V1=1 selects the one-element-per-thread forms of the spectral, overlap-add and V-transpose kernels.
inline bool mem_v1() {
    static const bool v = [] { const char* e = std::getenv("WT_MEM_V1"); return e && std::atoi(e) != 0; }();
    return v;
}

}  // namespace

void launch_softmax_planes(const float* S, int ldS, __half* p_hi, __half* p_lo, int Lpad, int B, int L, int Lp,
                           float scale, cudaStream_t s, Ragged rg) {
    if (B <= 0 || L <= 0) return;
    long long rows = (long long)B * Lp;
    softmax_planes_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(S, ldS, p_hi, p_lo, Lpad, B, L, Lp, scale, rg);
    WT_CUDA(cudaGetLastError());
}

void launch_vt_planes(const __half* q_hi, const __half* q_lo, __half* vt_hi, __half* vt_lo, int B, int L, int Lp, int C,
                      int Lpad, cudaStream_t s, Ragged rg) {
    if (B <= 0 || L <= 0) return;
    dim3 block(32, 8);
    if (!mem_v1() && C % 64 == 0 && Lpad % 2 == 0) {
        dim3 grid((Lpad + 63) / 64, C / 64, B);
        vt_planes2_kernel<<<grid, block, 0, s>>>(q_hi, q_lo, vt_hi, vt_lo, L, Lp, C, Lpad, rg);
    } else {
        dim3 grid((Lpad + 31) / 32, C / 32, B);
        vt_planes_kernel<<<grid, block, 0, s>>>(q_hi, q_lo, vt_hi, vt_lo, L, Lp, C, Lpad, rg);
    }
    WT_CUDA(cudaGetLastError());
}

void launch_groupnorm(const float* x, const float* w, const float* b, RowOut out, int B, int L, int Lp, int C,
                      int groups, float eps, int swish, cudaStream_t s, Ragged rg) {
    if (B <= 0 || L <= 0) return;
    if (groups != 32 || C != 768) throw Error(1, "groupnorm: expected GroupNorm(32, 768)");
    dim3 grid(C / (GN_GPB * 24), B);
    static const bool minb2 = [] { const char* e = std::getenv("WT_GN_MINB"); return e && std::atoi(e) == 2; }();
    if (L <= GN_RES * GN_PH && Lp <= L + GN_PH) {
        if (minb2) groupnorm_kernel<GN_RES, 2><<<grid, GN_THREADS, 0, s>>>(x, w, b, out, L, Lp, C, eps, swish, rg);
        else groupnorm_kernel<GN_RES><<<grid, GN_THREADS, 0, s>>>(x, w, b, out, L, Lp, C, eps, swish, rg);
    }
    else groupnorm_kernel<0><<<grid, GN_THREADS, 0, s>>>(x, w, b, out, L, Lp, C, eps, swish, rg);
    WT_CUDA(cudaGetLastError());
}

void launch_layernorm(const float* x, const float* w, const float* b, RowOut out, long long M, int C, float eps,
                      cudaStream_t s) {
    if (M <= 0) return;
    if (C != 768) throw Error(1, "layernorm: backbone dim must be 768");
    layernorm_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(x, w, b, out, M, eps);
    WT_CUDA(cudaGetLastError());
}

void launch_dwconv_ln(const float* x, const float* dw, const float* db, const float* scale, const float* shift,
                      RowOut out, int B, int L, int Lp, int C, float eps, cudaStream_t s, Ragged rg) {
    if (B <= 0 || L <= 0) return;
    if (C != 768) throw Error(1, "dwconv_ln: backbone dim must be 768");
    static const bool hi8 = [] { const char* e = std::getenv("WT_DW_HI8"); return !e || std::atoi(e) != 0; }();
    if (hi8 && out.hi && !out.lo && !out.f32) {  // single-plane consumer (plan 2)
        dim3 grid8((Lp + DW8_TT - 1) / DW8_TT, B);
        if (mem_v1()) dwconv_ln_hi_kernel<<<grid8, DW_THREADS, 0, s>>>(x, dw, db, scale, shift, out.hi, L, Lp, eps, rg);
        else dwconv_ln_hi2_kernel<<<grid8, DW_THREADS, 0, s>>>(x, dw, db, scale, shift, out.hi, L, Lp, eps, rg);
        WT_CUDA(cudaGetLastError());
        return;
    }
    dim3 grid((Lp + DW_TT - 1) / DW_TT, B);
    dwconv_ln_kernel<<<grid, DW_THREADS, 0, s>>>(x, dw, db, scale, shift, out, L, Lp, eps, rg);
    WT_CUDA(cudaGetLastError());
}

void launch_attention(const float* qkv, RowOut out, int B, int L, int Lp, int C, cudaStream_t s) {
    if (B <= 0 || L <= 0) return;
    if (C != 768) throw Error(1, "attention: backbone dim must be 768");
    size_t smem = (size_t)8 * L * sizeof(float);
    if (smem > 200 * 1024) throw Error(1, "attention: clip too long for the on-chip score buffer (L <= 6400)");
    static PerDevice<size_t> attr_dev;
    size_t& attr = attr_dev.get();
    if (smem > 48 * 1024 && smem > attr) {
        WT_CUDA(cudaFuncSetAttribute(attention_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = 200 * 1024;
    }
    dim3 grid((Lp + 7) / 8, B);
    attention_kernel<24><<<grid, 256, smem, s>>>(qkv, out, L, Lp, 1.0f / sqrtf((float)C));
    WT_CUDA(cudaGetLastError());
}

void launch_spectral(const float* z, int ldz, RowOut S, long long M, int half, int ldS, cudaStream_t s) {
    if (M <= 0) return;
    if (ldS < 2 * half) throw Error(1, "spectral: ldS must cover both halves");
    if (!mem_v1()) {
        spectral_rows_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(z, ldz, S, M, half, ldS);
    } else {
        long long n = M * (ldS - half);
        spectral_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(z, ldz, S, M, half, ldS);
    }
    WT_CUDA(cudaGetLastError());
}

void launch_overlap_add(const float* frames, const float* wsq, float* audio, int B, int L, int Lp, int n_fft, int hop,
                        cudaStream_t s, Ragged rg) {
    if (B <= 0 || L <= 0) return;
    const int pad = (n_fft - hop) / 2;
    const bool vec4 = !mem_v1() && hop % 4 == 0 && n_fft % 4 == 0 && pad % 4 == 0 &&
                      (((uintptr_t)frames | (uintptr_t)wsq | (uintptr_t)audio) & 15) == 0;
    if (vec4) {
        dim3 grid((L * hop / 4 + 255) / 256, B);
        overlap_add4_kernel<<<grid, 256, 0, s>>>(frames, wsq, audio, L, Lp, n_fft, hop, pad, rg);
    } else {
        dim3 grid((L * hop + 255) / 256, B);
        overlap_add_kernel<<<grid, 256, 0, s>>>(frames, wsq, audio, L, Lp, n_fft, hop, pad, rg);
    }
    WT_CUDA(cudaGetLastError());
}

void launch_features_to_rows(const float* in, RowOut out, int B, int C, int L, int Lp, cudaStream_t s, Ragged rg) {
    if (B <= 0 || L <= 0) return;
    dim3 grid((Lp + 31) / 32, (C + 31) / 32, B), block(32, 8);
    features_to_rows_kernel<<<grid, block, 0, s>>>(in, out, C, L, Lp, rg);
    WT_CUDA(cudaGetLastError());
}

}  // namespace wt
