// tcgen05 GEMM launch descriptor (see gemm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace wt {

// One A-operand segment: a 2-D fp16 tensor (split planes hi/lo) whose row r starts `stride` elements
// after row r-1 and exposes `inner` contiguous elements. stride < inner gives OVERLAPPING rows: the
// im2col matrix of a Conv1d over channels-last data without materialising it (row stride = conv stride
// * channels, inner = kernel * channels). Columns past `inner` and rows outside [0, rows) read as zero
// (TMA out-of-bounds fill). The segment contributes num_kb k-blocks of 64; k-block kb loads columns
// (kb % kb_per_tap) * 64 of row  m + shift0 + kb / kb_per_tap  ("tap" mode: kb_per_tap = Cin / 64;
// "window" mode: kb_per_tap = num_kb).
struct TcSeg {
    const __half* hi = nullptr;
    const __half* lo = nullptr;
    long long rows = 0;
    long long inner = 0;
    long long stride = 0;
    int num_kb = 0;
    int kb_per_tap = 1;
    int shift0 = 0;
};

// Output row remap: GEMM row m = b*Pin + t (valid iff t < Tvalid) is stored at row b*Pout + off + t, and
// mirrored into the reflect halo rows [off - hl, off) and [off + Tvalid, off + Tvalid + hr) of the padded
// layout its consumer reads (reference encoder/modules/conv.py:79-96 reflect padding). Pin == 0: identity.
// Destination row = b*sb + (off + t)*st with sb = Pout, st = 1 by default (batch-major); sb = 1, st = #clips
// gives the time-major rows [t*B + b] the LSTM steps read as one contiguous [B, D] block per time step.
struct RowMap {
    int Pin = 0, Tvalid = 0, Pout = 0, off = 0, hl = 0, hr = 0;
    int sb = 0, st = 0;
};

enum : int { TC_ACT_NONE = 0, TC_ACT_GELU = 1, TC_ACT_LSTM = 2, TC_ACT_ARGMIN = 3 };

struct TcGemm {
    TcSeg seg[2];
    int nseg = 1;
    // W: split-fp16 planes [N, K], K = 64 * (seg[0].num_kb + seg[1].num_kb)
    const __half* W_hi = nullptr;
    const __half* W_lo = nullptr;
    int M = 0, N = 0, K = 0;
    long long ldw = 0;     // row pitch of W in elements (0: K)
    long long w_rows = 0;  // rows of the W tensor (0: N)
    int passes = 3;  // 3: hi*hi + hi*lo + lo*hi; 1: hi*hi only
    // k-block width in elements shared by all segments and W: 64 (128-byte swizzle rows), or 32 / 16 for narrow
    // operands (64- / 32-byte rows): TMA then moves only bytes that exist. num_kb / kb_per_tap count kw-wide blocks.
    int kw = 64;
    // L2 prefetch of the A operand a few k-blocks ahead of the TMA ring (long-K GEMMs whose A streams from HBM)
    int prefetch = 0;
    // cap on the persistent grid (0: one CTA per SM): lets the kernel share the GPU with a resident cooperative
    // kernel (the LSTM wavefront) instead of queueing CTAs behind it
    int max_ctas = 0;
    // batched mode (attention): `batch` independent problems of M x N x K; problem z reads A rows shifted by
    // z*a_brows, W rows shifted by z*w_brows and writes rows shifted by z*o_brows
    int batch = 1;
    long long a_brows = 0, w_brows = 0, o_brows = 0;
    // epilogue: v = acc + bias; act; v *= gamma; v += res[m]; stores of v and/or ELU(v)
    const float* bias = nullptr;
    const float* gamma = nullptr;
    const float* res = nullptr;  // indexed by the GEMM row m (no remap)
    int ldres = 0;
    int act = TC_ACT_NONE;
    RowMap map;
    float* out_f32 = nullptr;  // v
    int ldo = 0;
    __half* out_hi = nullptr;  // split planes of v
    __half* out_lo = nullptr;
    int ldh = 0;
    __half* elu_hi = nullptr;  // split planes of ELU(v)
    __half* elu_lo = nullptr;
    int ldh2 = 0;
    // TC_ACT_LSTM (see gemm_tc.cu): columns are [i | f | g | o] blocks of 16 units per 64-wide tile
    float* cell = nullptr;     // [M, H] cell state, updated in place
    int hidden = 0;
    // split planes receive v - plane_shift[n] (the VQ reads frames centred on the codebook mean)
    const float* plane_shift = nullptr;
    // TC_ACT_ARGMIN: per row argmin_n (bias[n] - 2*acc[m, n]) merged across tiles with a packed atomicMin
    unsigned long long* best = nullptr;
    // optional per-CTA timeline (clock64 stamps, 64 slots per CTA) for performance debugging
    long long* dbg = nullptr;
};

void set_debug_timeline(long long* dev_buf);

// What the last launch_tap_gemm_tc on this thread ran: kernel variant id (BN * 10 + passes, e.g. 2563 =
// tap_gemm_tc_kernel<256, 3>) and its algorithmic FLOPs 2*M*N*K*batch (split-precision passes not counted).
struct LaunchInfo { int kern = 0; double flops = 0; };
LaunchInfo& last_launch_info();

void launch_tap_gemm_tc(const TcGemm& g, cudaStream_t s);
// One LSTM layer, all L steps, in a single cooperative launch (see gemm_tc.cu). Time-major tensors.
// `counters` needs lstm_counter_ints(B, L) ints.
size_t lstm_counter_ints(int B, int L);
// Runs steps [t_begin, t_end) (t_end < 0: L); the counters are zeroed by the launch that starts at step 0, later
// segments rely on h / c / counters of the earlier ones. lstm_ctas(B): CTAs the launch keeps resident.
void launch_lstm_persistent(const float* xin, float* y, __half* h_hi, __half* h_lo, float* cell, int* counters,
                            const __half* w_hi, const __half* w_lo, int B, int L, int D, cudaStream_t s, int t_begin = 0,
                            int t_end = -1);
int lstm_ctas(int B, int D);
void launch_split_f16(const float* x, __half* hi, __half* lo, long long rows, int cols, long long ld_in,
                      long long ld_out, cudaStream_t s);

// Cached tensor map of a 2-D fp16 tensor (`rows` rows of `inner` elements, row r starting r*stride elements after the
// base; stride < inner: overlapping rows) with a [box_rows, kw] box under the swizzle that matches kw (64 / 32 / 16).
const CUtensorMap& tc_make_map(const __half* base, long long rows, long long inner, long long stride, int box_rows, int kw);
int tc_num_sms();
long long* tc_debug_timeline();

// Levels 0 -> 1 of the SEANet encoder in ONE tcgen05 kernel (enc_fused.cu): strided conv (32 -> 64, k4 s2 or k8 s4) of the ELU(y0)
// planes, ELU, k3 conv (64 -> 32), ELU, 1x1 conv (32 -> 64) + composed shortcut, ELU -> planes in the padded layout of
// the next strided conv. x1, ELU(x1) and ELU(h1) never leave the SM (they live in tensor memory).
struct EncL1Weights {
    const __half* w1 = nullptr;  // [256, k0]: rows [Wc_hi | Wd_hi | Wc_lo | Wd_lo] (composed shortcut, strided conv)
    const __half* w2 = nullptr;  // [192, 64]: rows [Wk3_hi | Wk3_lo], row = tap * 32 + cout
    const __half* w3 = nullptr;  // [128, 32]: rows [W1x1_hi | W1x1_lo]
    const float* bias = nullptr; // b_d[64] | b_k3[32] | b_tail[64]
    int k0 = 128;                // window width 2 * stride * 32: 128 (stride 2) or 256 (stride 4) = row length of w1
};
struct EncL1Args {
    const __half* y0_hi = nullptr;  // ELU(y0) planes of level 0: clip pitch stride * (T1 + 2) rows of 32 channels
    const __half* y0_lo = nullptr;
    long long y0_elems = 0;
    int Bc = 0, T1 = 0;             // clips, level-1 length; row space m = b * (T1 + 2) + t
    RowMap map;                     // destination layout of ELU(y1) (Pin = T1 + 2, Tvalid = T1)
    __half* ye_hi = nullptr;        // ELU(y1) planes, 64 channels per row
    __half* ye_lo = nullptr;
    float* y_f32 = nullptr;         // optional fp32 copy of y1 at the same rows (debug tap)
};
bool enc_l1_fused_supported(int cin, int stride);
void launch_enc_l1_fused(const EncL1Weights& w, const EncL1Args& a, cudaStream_t s);

// Level 0 of the SEANet encoder (conv0 + ResBlock 0) with the k3 and 1x1 products on the tensor cores (enc_l0_tc.cu);
// conv0 and the composed shortcut stay on the CUDA cores with their weights in the kernel-parameter constant bank.
struct EncL0Weights {
    const __half* wk3 = nullptr;   // [96, 32]: rows [Wk3_hi | Wk3_lo], row = tap * 16 + cout
    const __half* w1x1 = nullptr;  // [64, 16]: rows [W1x1_hi | W1x1_lo]
    const float* consts = nullptr; // HOST: w0[7][32] | b0[32] | wsc[7][32] | b1[16] | b2[32]
};
bool enc_l0_tc_supported();
void launch_enc_l0_tc(const EncL0Weights& w, const float* wav, __half* ye_hi, __half* ye_lo, float* y_f32, int B, int T,
                      const RowMap& map, cudaStream_t s);

// Convenience: a dense / tap-mode segment over rows [rows, Cin] with pitch lda.
inline TcSeg tc_taps(const __half* hi, const __half* lo, long long rows, int Cin, int lda, int taps, int center) {
    TcSeg s;
    s.hi = hi; s.lo = lo; s.rows = rows; s.inner = Cin; s.stride = lda;
    s.kb_per_tap = Cin / 64; s.num_kb = taps * s.kb_per_tap; s.shift0 = -center;
    return s;
}
// Window-mode segment: row r = elements [r*stride, r*stride + inner) of a flat buffer holding total_elems.
inline TcSeg tc_window(const __half* hi, const __half* lo, long long total_elems, long long inner, long long stride,
                       int shift0 = 0, int kw = 64) {
    TcSeg s;
    s.hi = hi; s.lo = lo; s.inner = inner; s.stride = stride;
    s.rows = total_elems >= inner ? (total_elems - inner) / stride + 1 : 0;
    s.num_kb = (int)((inner + kw - 1) / kw); s.kb_per_tap = s.num_kb; s.shift0 = shift0;
    return s;
}

}  // namespace wt
