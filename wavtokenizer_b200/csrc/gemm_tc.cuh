// tcgen05 tap-GEMM launch descriptor (see gemm_tc.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace wt {

struct TcGemm {
    // A: split-fp16 planes [rowsA, Cin] (row pitch lda elements) in padded row space
    const __half* A_hi = nullptr;
    const __half* A_lo = nullptr;
    long long rowsA = 0;
    int Cin = 0, lda = 0;
    int taps = 1, center = 0;  // out row m reads A rows m + j - center, j < taps (rows outside [0, rowsA) are zero)
    // W: split-fp16 planes [N, K], K = taps*Cin, K index = tap*Cin + c
    const __half* W_hi = nullptr;
    const __half* W_lo = nullptr;
    int M = 0, N = 0, K = 0;
    int passes = 3;  // 3: hi*hi + hi*lo + lo*hi; 1: hi*hi only
    // epilogue: v = acc + bias; act; v *= gamma; v += res; store fp32 and/or split fp16
    const float* bias = nullptr;
    const float* gamma = nullptr;
    const float* res = nullptr;
    int ldres = 0;
    int act = 0;
    float* out_f32 = nullptr;
    int ldo = 0;
    __half* out_hi = nullptr;
    __half* out_lo = nullptr;
    int ldh = 0;
};

void launch_tap_gemm_tc(const TcGemm& g, cudaStream_t s);
void launch_split_f16(const float* x, __half* hi, __half* lo, long long rows, int cols, long long ld_in,
                      long long ld_out, cudaStream_t s);

}  // namespace wt
