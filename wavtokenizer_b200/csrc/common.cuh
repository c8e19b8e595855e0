// Shared declarations for the wavtok_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace wt {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// thread-local message behind wt_last_error() (model.cu)
void set_last_error(const std::string& msg);

#define WT_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw ::wt::Error(4, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

// Per-device launch state. cudaFuncSetAttribute, occupancy queries and the SM count belong to ONE device: a cache kept
// in a plain static would leave the second GPU of a process without its shared-memory opt-in (every launch there
// would then fail with "invalid argument"). State is kept per device ordinal instead; a benign race at most repeats
// an idempotent call.
constexpr int WT_MAX_DEVICES = 64;
inline int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < WT_MAX_DEVICES) ? d : 0;
}
template <typename T>
struct PerDevice {
    T v[WT_MAX_DEVICES] = {};
    T& get() { return v[current_device()]; }
};
// Makes `dev` current for a scope and restores the caller's device afterwards (the C ABI must not change the current
// device of the calling thread: torch and other libraries in the process rely on it).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) {
            cudaError_t e = cudaSetDevice(dev);
            if (e != cudaSuccess) throw Error(4, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
            switched = true;
        }
    }
    ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// ---------------------------------------------------------------------------------------
// tap-GEMM: the one contraction shape of this path.
//   out[m, n] = epi( sum_{j<taps} sum_{c<Cin} pro(A[src(m, j), c]) * W[n, j*Cin + c] )
// Activations are channels-last rows [B*T, C]; a Conv1d with k taps and stride s is a GEMM
// whose A row for output (b, t) is the k rows  t*s - pad_left + j  of clip b, resolved with
// zero or reflect padding in the loader (encoder/modules/conv.py:79-96, 195-211).
// ---------------------------------------------------------------------------------------
enum : int { PAD_ZERO = 0, PAD_REFLECT = 1 };
enum : int { PRO_NONE = 0, PRO_ELU = 1 };
enum : int { ACT_NONE = 0, ACT_GELU = 1 };

struct TapGemm {
    const float* A = nullptr;     // [B*Tin, lda]
    const float* W = nullptr;     // [N, K]   K = taps*Cin
    const float* bias = nullptr;  // [N] or null
    const float* gamma = nullptr; // [N] or null: out = res + gamma*(acc + bias)
    const float* res = nullptr;   // [M, ldres] or null
    float* out = nullptr;         // [M, ldo]
    int M = 0, N = 0, K = 0;
    int Cin = 0, taps = 1, stride = 1, pad_left = 0;
    int Tin = 0, Tout = 0;        // rows per clip on the input / output side
    int Trefl = 0;                // reflect length T' = max(Tin, max_pad + 1) (conv.py:86-94)
    int pad_mode = PAD_ZERO;
    int pro = PRO_NONE;
    int act = ACT_NONE;
    int lda = 0, ldo = 0, ldres = 0;
};

void launch_tap_gemm_simt(const TapGemm& g, cudaStream_t s);

// encoder
void launch_conv0(const float* wav, const float* w /*[C,7]*/, const float* bias, float* out, int B, int T, int C,
                  cudaStream_t s);
void launch_lstm_pointwise(const float* gates /*row b*ldg, 4H wide*/, float* c /*[B,H]*/, float* y /*row b*ldy*/,
                           int B, int H, long long ldg, long long ldy, cudaStream_t s);
void launch_add(const float* a, const float* b, float* out, long long n, cudaStream_t s);
// tcgen05 encoder helpers (padded, split-fp16 layouts; see model.cu encoder_front_tc)
void launch_conv0_planes(const float* wav, const float* w, const float* bias, __half* win_hi /*[B*(T+2), 8]*/,
                         __half* win_lo, __half* elu_hi /*[B*(T+2), C]*/, __half* elu_lo, int B, int T, int C,
                         cudaStream_t s);
// level 0 (conv0 + ResBlock 0) fused, fp32 CUDA cores (encoder_ops.cu). pack: w0t[7][32] b0[32] w1t[96][16] b1[16]
// w2t[16][32] wsct[8][32] b2[32]; output ELU(y) planes [B*Py, 32] with data at row `left` and reflect halo left / hr.
int resblock0_pack_floats();
void launch_resblock0_fused(const float* wav, const float* pack, __half* ye_hi, __half* ye_lo, float* y_f32, int B, int T,
                            int Py, int left, int hr, cudaStream_t s);
// len_tab (device, [B], optional): per-clip frame counts of a ragged batch; L is then the longest clip (the row pitch)
void launch_lstm_skip_elu_pad(const float* y, const float* x, float* out_f32, __half* elu_hi, __half* elu_lo, int B,
                              int L, int D, cudaStream_t s, const int* len_tab = nullptr);

// SEANet decoder: gather of a transposed conv's per-tap GEMM output g [B*L, 2*stride*Cout] into
// y [B, L*stride, Cout] (+ bias), trimmed by `left` samples in front (reference encoder/modules/conv.py:232-253)
void launch_convtr_gather(const float* g, const float* bias, float* y, int B, int L, int Cout, int stride, int left,
                          cudaStream_t s);

// vq
void launch_vq_simt(const float* x, const float* codebook, const float* cnorm, long long N, int D, int bins,
                    long long* codes, cudaStream_t s);
void launch_gather_rows(const float* codebook, const long long* codes, float* out, long long N, int D, int bins,
                        int* err_flag, cudaStream_t s);
// features [B, D, L] (channel-major) = sum_k codebook[k][codes[k, b, t]]
void launch_codes_to_features(const float* codebooks, const long long* codes, float* out, int K, int B, int L, int D,
                              int bins, int* err_flag, cudaStream_t s);

// codebook rows (split planes) of codes [B, L] -> decoder row planes [B*Lp, D] with zero halo rows
void launch_codes_to_row_planes(const __half* cb_hi, const __half* cb_lo, const long long* codes, __half* hi, __half* lo,
                                int B, int L, int Lp, int D, int bins, cudaStream_t s);

// tcgen05 VQ helpers
void launch_center_split(const float* x, const float* mu, __half* hi, __half* lo, long long N, int D, cudaStream_t s);
void launch_best_to_codes(const unsigned long long* best, long long* codes, long long N, cudaStream_t s);

// layout
void launch_transpose_bcl_to_blc(const float* in, float* out, int B, int C, int L, cudaStream_t s);
void launch_transpose_blc_to_bcl(const float* in, float* out, int B, int L, int C, cudaStream_t s);

// Row-wise output of a producer kernel: fp32 rows and/or the split-fp16 planes of the next GEMM's A operand.
struct RowOut {
    float* f32 = nullptr;
    __half* hi = nullptr;
    __half* lo = nullptr;
};
inline RowOut out_f32(float* p) { RowOut o; o.f32 = p; return o; }
inline RowOut out_split(__half* hi, __half* lo, float* also_f32 = nullptr) {
    RowOut o; o.hi = hi; o.lo = lo; o.f32 = also_f32; return o;
}

// Ragged decoder batch: every clip keeps the common row pitch Lp = Lmax + 3, clip b holds len[b]
// <= Lmax frames and everything from there to the pitch is halo (zeros wherever a k-tap conv reads it). off[b] = frames of
// the clips before b: where clip b sits in the PACKED API tensors (features in, audio out). Null pointers: uniform batch.
struct Ragged {
    const int* len = nullptr;
    const long long* off = nullptr;
};

// decoder (clip b owns rows [b*Lp, b*Lp + L); Lp - L halo rows after each clip are written as zeros)
void launch_features_to_rows(const float* in /*[B,C,L]*/, RowOut out, int B, int C, int L, int Lp, cudaStream_t s,
                             Ragged rg = Ragged{});
void launch_groupnorm(const float* x, const float* w, const float* b, RowOut out, int B, int L, int Lp, int C,
                      int groups, float eps, int swish, cudaStream_t s, Ragged rg = Ragged{});
void launch_layernorm(const float* x, const float* w, const float* b, RowOut out, long long M, int C, float eps,
                      cudaStream_t s);
void launch_dwconv_ln(const float* x, const float* dw /*[7,C] (taps transposed at load)*/, const float* db, const float* scale,
                      const float* shift, RowOut out, int B, int L, int Lp, int C, float eps, cudaStream_t s,
                      Ragged rg = Ragged{});
void launch_attention(const float* qkv /*[B*Lp, 3C]*/, RowOut out /*[B*Lp, C]*/, int B, int L, int Lp, int C,
                      cudaStream_t s);
// tensor-core attention helpers (scores and P.V run as batched tcgen05 GEMMs)
void launch_softmax_planes(const float* S, int ldS, __half* p_hi, __half* p_lo, int Lpad, int B, int L, int Lp,
                           float scale, cudaStream_t s, Ragged rg = Ragged{});
void launch_vt_planes(const __half* q_hi, const __half* q_lo, __half* vt_hi, __half* vt_lo, int B, int L, int Lp, int C,
                      int Lpad, cudaStream_t s, Ragged rg = Ragged{});
void launch_spectral(const float* z /*[M, ldz]*/, int ldz, RowOut S /*[M, ldS]*/, long long M, int half, int ldS,
                     cudaStream_t s);
void launch_overlap_add(const float* frames /*[B*Lp, n_fft]*/, const float* wsq /*[n_fft]*/, float* audio, int B, int L,
                        int Lp, int n_fft, int hop, cudaStream_t s, Ragged rg = Ragged{});

}  // namespace wt
