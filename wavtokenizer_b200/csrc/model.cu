// Model handle, weight preparation, workspace arena and the C ABI (include/wavtok_b200.h).
// Host-side orchestration of the hot path: encode_infer -> VQ -> codes_to_features -> decode
// (reference decoder/pretrained.py:186-239).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/wavtok_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"

namespace wt {

namespace {

thread_local std::string g_last_error;

}  // namespace

void set_last_error(const std::string& msg) { g_last_error = msg; }

namespace {

// Clips per encoder-front pass (early SEANet tensors: 9.2 MB per clip each). Device-resident calls use 128 (fewer, longer
// launches: -0.3..-0.45 ms per 256-clip step against 64); the host-buffer entry point uses 64: its H2D pieces and the
// per-piece level-0 launches pipeline at the finer grain (same e2e time either way, half the workspace).
constexpr int ENC_CHUNK_DEFAULT = 128, ENC_CHUNK_HOST = 64;
thread_local bool g_host_entry = false;  // set by wt_encode_decode_host for the duration of the call
inline int enc_chunk() {
    static int env = [] {
        const char* e = std::getenv("WT_ENC_CHUNK");
        int n = e ? std::atoi(e) : 0;
        return n >= 1 && n <= 256 ? n : 0;
    }();
    return env ? env : (g_host_entry ? ENC_CHUNK_HOST : ENC_CHUNK_DEFAULT);
}
#define ENC_CHUNK enc_chunk()
// clips per sub-chunk of encoder levels 0-1 (encoder_front_tc); 0 = off. Measured (profiles/r02_enc_subchunk_sweep.md):
// sub-chunks of 2 / 4 / 8 / 16 clips are SLOWER than whole 64-clip chunks (36.4 / 34.2 / 33.2 / 32.3 vs 32.3 ms per
// step): the planes of even 4 clips (37 MB read + 74 MB written by the level-0 strided conv) do not survive in L2
// between producer and consumer, and each extra launch costs ~6 us of prologue + tail. Kept as a tunable, off.
constexpr int ENC_SUB_DEFAULT = 0;
inline int enc_sub() {
    static int v = [] {
        const char* e = std::getenv("WT_ENC_SUB");
        int n = e ? std::atoi(e) : ENC_SUB_DEFAULT;
        return n >= 0 && n <= 256 ? n : ENC_SUB_DEFAULT;
    }();
    return v;
}
inline int tc_prefetch() {
    static const int v = [] { const char* e = std::getenv("WT_TC_PREFETCH"); return e ? std::atoi(e) : 0; }();
    return v;
}
constexpr int ENC_GROUP = 1024; // clips per LSTM / final-conv / VQ pass (the recurrent GEMM's M)
constexpr int COPY_PIECE = 16;  // clips per H2D / D2H piece of the host-buffer entry point (4.6 MB of audio)
// Clips per decoder pass. Device-resident calls: 256 (29 k -> 58 k rows: the 768-wide GEMMs fill 9.2 -> 18.5 waves of CTA
// pairs, -1.0 ms per 256-clip step); host-buffer entry: 128, so that the audio of the first chunk leaves while the
// second one is computed (same e2e time as 256, measured).
constexpr int DEC_CHUNK_DEFAULT = 256, DEC_CHUNK_HOST = 128;
inline int dec_chunk() {
    static int env = [] {
        const char* e = std::getenv("WT_DEC_CHUNK");
        int n = e ? std::atoi(e) : 0;
        return n >= 1 && n <= 1024 ? n : 0;
    }();
    return env ? env : (g_host_entry ? DEC_CHUNK_HOST : DEC_CHUNK_DEFAULT);
}
#define DEC_CHUNK dec_chunk()

struct ConvW {
    float* w = nullptr;  // [cout, k*cin], K index = tap*cin + c
    float* b = nullptr;  // [cout]
    __half* w_hi = nullptr;  // split-fp16 planes of w for the tcgen05 path
    __half* w_lo = nullptr;
    int cout = 0, cin = 0, k = 1, stride = 1;
    std::vector<float> hw, hb;  // host copies (relayouted weight, bias) for composing fused weights
};

struct HalfW {  // split-fp16 planes of a plain [N, K] weight
    __half* hi = nullptr;
    __half* lo = nullptr;
};

struct TapReq {
    float* buf = nullptr;
    int64_t cap = 0;
    int B = 0, T = 0, C = 0;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

}  // namespace wt

using namespace wt;

struct wt_handle {
    wt_config cfg{};
    int device = 0;
    int plan = 2;  // default: tcgen05 everywhere, single-pass fp16 for the ConvNeXt GEMMs (wt_set_plan)
    int64_t launches = 0;
    std::vector<void*> owned;  // weight allocations

    // encoder
    float *conv0_w = nullptr, *conv0_b = nullptr;
    struct ResBlk { ConvW c1, c2, sc; } rb[4];
    ConvW down[4];
    struct Lstm { float *w_ih = nullptr, *w_hh = nullptr, *bias = nullptr; } lstm[4];
    ConvW enc_last;
    // tcgen05 encoder weights: k3 conv padded to K % 64 == 0; conv1x1 and shortcut fused along K
    struct RbTc { HalfW w1; int Kp1 = 0, kw1 = 64; HalfW w2; int kb0 = 0, kb1 = 0, kw2 = 64; float* bias2 = nullptr; } rb_tc[4];
    // LSTM weights with gate rows permuted so that each 64-wide tile holds [i | f | g | o] x 16 hidden units
    struct LstmTc { HalfW w_ih, w_hh; float* bias = nullptr; } lstm_tc[4];
    float* rb0_pack = nullptr;   // level-0 fused kernel weights (encoder_ops.cu resblock0_fused_kernel)
    EncL1Weights l1_fused;       // level 0 -> 1 fused tcgen05 kernel weights (enc_fused.cu); w1 == nullptr: not available
    // SEANet decoder (feature_extractor.encodec.decoder): optional, fp32 CUDA-core plan only
    struct SeanetDec {
        bool present = false;
        ConvW first, last;
        Lstm lstm[4];
        struct Up { float* w = nullptr; float* b = nullptr; int cin = 0, cout = 0, stride = 1; } up[4];
        ResBlk rb[4];
    } sdec;
    EncL0Weights l0_tc;          // level 0 with tensor-core k3 / 1x1 products (enc_l0_tc.cu); wk3 == nullptr: not available
    std::vector<float> l0_consts;
    float* zero_rows = nullptr;  // zero planes standing in for h_{-1}
    // vq
    float* codebooks = nullptr;  // [num_quantizers * bins, D]
    float* cnorm = nullptr;      // [bins] of codebook 0
    // tcgen05 VQ: codebook 0 centred on its mean (distances unchanged, far better conditioned)
    float* cb_mean = nullptr;    // [D]
    HalfW cb_c;                  // split planes of (C - mean) [bins, D]
    HalfW cb_h;                  // split planes of codebook 0 itself (host-buffer entry: codes -> decoder row planes)
    const long long* dec_codes = nullptr;  // set by wt_encode_decode_host around do_decode: codes [B, L] of `features`
    float* cnorm_c = nullptr;    // ||C - mean||^2 [bins]
    // decoder
    ConvW embed;
    struct Resnet { float *n1w, *n1b, *n2w, *n2b; ConvW c1, c2; } pos[4];
    struct Attn { float *nw, *nb, *wqkv, *bqkv; ConvW proj; HalfW wqkv_h; } attn{};
    float *gn5w = nullptr, *gn5b = nullptr;
    float *norm_scale = nullptr, *norm_shift = nullptr;  // [E, dim]
    struct Cnx { float *dw, *db, *scale, *shift, *w1, *b1, *w2, *b2, *gamma; HalfW w1_h, w2_h; };
    std::vector<Cnx> cnx;
    float *fln_w = nullptr, *fln_b = nullptr;
    float *head_w = nullptr, *head_b = nullptr;
    float *basis = nullptr, *wsq = nullptr;
    HalfW head_h, basis_h;
    int Kp = 0;   // padded K of the inverse-DFT GEMM (multiple of 64)
    int ldz = 0;  // row pitch of the head output (n_fft + 2 rounded up to 4 floats)

    // workspace arena (bump allocator, reset per call)
    char* arena = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    // persistent staging for the host-buffer entry point
    char* stage = nullptr;
    size_t stage_cap = 0;
    // host-buffer entry point: copy stream + events so that H2D of encoder chunk i+1 overlaps the encoder of chunk
    // i and D2H of decoder chunk i overlaps the decoder of chunk i+1; hooks are set only inside that entry point
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // second layer of the LSTM wavefront (encoder_back_tc)
    std::vector<cudaEvent_t> aux_evs;
    std::vector<cudaStream_t> lane_streams;  // conv fronts of a ragged batch run on several streams (do_encode_ragged)
    std::vector<cudaEvent_t> lane_evs;
    std::vector<cudaEvent_t> copy_evs;
    // wav_ready(first clip, clips): called right before the first kernel that reads those clips' audio;
    // audio_done(first clip, clips): called right after the kernel that wrote those clips' audio. Both work on PIECES of
    // COPY_PIECE clips (finer than the compute chunks), so that only one piece of H2D stands in front of the first kernel
    // and only one piece of D2H behind the last one.
    std::function<void(int /*first clip*/, int /*clips*/)> wav_ready, audio_done;
    // Sticky device-side error flag (code out of range in a gather). It is NOT read back synchronously: the launch is
    // followed by an async copy into pinned memory plus an event, and the flag is examined when that event has
    // completed -- at the next call on the handle, at wt_check_errors(), or at the end of wt_encode_decode_host.
    int* err_flag = nullptr;  // device
    int* err_host = nullptr;  // pinned
    cudaEvent_t err_ev = nullptr;
    bool err_armed = false;
    std::string err_what;

    std::map<std::string, TapReq> taps;

    // optional per-category kernel timing (CUDA events on the launching stream; bench.py's roofline)
    bool timing = false;
    struct Ev { cudaEvent_t a, b; int cat; int kern = 0; double flops = 0; double bytes = 0; };
    std::vector<Ev> evs;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t get_event() {
        if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
        cudaEvent_t e;
        WT_CUDA(cudaEventCreate(&e));
        return e;
    }

    ~wt_handle() {  // the caller (wt_destroy) has made `device` current
        for (auto& e : evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        for (auto e : ev_pool) cudaEventDestroy(e);
        for (auto e : copy_evs) cudaEventDestroy(e);
        for (auto e : aux_evs) cudaEventDestroy(e);
        if (err_ev) cudaEventDestroy(err_ev);
        for (auto e : lane_evs) cudaEventDestroy(e);
        for (auto st : lane_streams) cudaStreamDestroy(st);
        if (aux_stream) cudaStreamDestroy(aux_stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        for (void* p : owned) cudaFree(p);
        if (arena) cudaFree(arena);
        if (stage) cudaFree(stage);
        if (err_flag) cudaFree(err_flag);
        if (err_host) cudaFreeHost(err_host);
    }

    float* upload(const std::vector<float>& v) {
        float* d = nullptr;
        WT_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(float)));
        owned.push_back(d);
        if (!v.empty()) WT_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
        return d;
    }

    __half* upload_halves(const std::vector<__half>& v) {
        __half* d = nullptr;
        WT_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 8) * sizeof(__half)));
        owned.push_back(d);
        if (!v.empty()) WT_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(__half), cudaMemcpyHostToDevice));
        return d;
    }

    HalfW upload_split(const std::vector<float>& v) {
        std::vector<__half> hi(v.size()), lo(v.size());
        for (size_t i = 0; i < v.size(); ++i) {
            hi[i] = __float2half_rn(v[i]);
            lo[i] = __float2half_rn(v[i] - __half2float(hi[i]));
        }
        HalfW w;  // one allocation, lo right after hi: a single 3-D TMA box fetches both planes of a tile
        const size_t n8 = align_up(std::max<size_t>(v.size(), 1), 8);
        WT_CUDA(cudaMalloc(&w.hi, 2 * n8 * sizeof(__half)));
        owned.push_back(w.hi);
        w.lo = w.hi + n8;
        WT_CUDA(cudaMemcpy(w.hi, hi.data(), v.size() * sizeof(__half), cudaMemcpyHostToDevice));
        WT_CUDA(cudaMemcpy(w.lo, lo.data(), v.size() * sizeof(__half), cudaMemcpyHostToDevice));
        return w;
    }

    void ensure_arena(size_t bytes) {
        if (bytes <= arena_cap) return;
        WT_CUDA(cudaDeviceSynchronize());
        if (arena) WT_CUDA(cudaFree(arena));
        arena = nullptr;
        arena_cap = 0;
        WT_CUDA(cudaMalloc(&arena, bytes));
        arena_cap = bytes;
    }
    float* alloc(size_t n_floats) {
        size_t bytes = align_up(n_floats * sizeof(float), 256);
        if (arena_off + bytes > arena_cap) throw Error(4, "workspace arena overflow (internal sizing bug)");
        float* p = reinterpret_cast<float*>(arena + arena_off);
        arena_off += bytes;
        return p;
    }

    // src rows: clip b at row b*Tp, T valid rows of C floats with row pitch ld
    void tap(const char* name, const float* src, int B, int T, int C, int b0, cudaStream_t s, int Tp = 0, int ld = 0) {
        if (taps.empty() || b0 != 0) return;
        auto it = taps.find(name);
        if (it == taps.end()) return;
        if (!Tp) Tp = T;
        if (!ld) ld = C;
        it->second.B = B; it->second.T = T; it->second.C = C;
        int64_t n = (int64_t)B * T * C;
        if (!it->second.buf || n > it->second.cap) return;
        if (Tp == T && ld == C) {
            WT_CUDA(cudaMemcpyAsync(it->second.buf, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
        } else {
            for (int b = 0; b < B; ++b)
                WT_CUDA(cudaMemcpy2DAsync(it->second.buf + (size_t)b * T * C, (size_t)C * sizeof(float),
                                          src + (size_t)b * Tp * ld, (size_t)ld * sizeof(float),
                                          (size_t)C * sizeof(float), T, cudaMemcpyDeviceToDevice, s));
        }
    }
};

namespace wt {
namespace {

// ---------------------------------------------------------------------------------------
// weight preparation (host)
// ---------------------------------------------------------------------------------------
struct Table {
    std::unordered_map<std::string, std::pair<const float*, int64_t>> m;
    bool has(const std::string& name) const { return m.find(name) != m.end(); }
    const float* get(const std::string& name, int64_t numel) const {
        auto it = m.find(name);
        if (it == m.end()) throw Error(WT_ERR_VALUE, "missing checkpoint tensor: " + name);
        if (it->second.second != numel)
            throw Error(WT_ERR_VALUE, "size mismatch for " + name + ": got " + std::to_string(it->second.second) +
                                          ", expected " + std::to_string(numel));
        return it->second.first;
    }
};

// [cout, cin, k] -> [cout, k*cin] with K index tap*cin + c
std::vector<float> relayout_conv(const float* w, int cout, int cin, int k) {
    std::vector<float> o((size_t)cout * cin * k);
    for (int n = 0; n < cout; ++n)
        for (int c = 0; c < cin; ++c)
            for (int j = 0; j < k; ++j) o[((size_t)n * k + j) * cin + c] = w[((size_t)n * cin + c) * k + j];
    return o;
}

// old-style weight_norm fold, w = g * v / ||v|| over (cin, k) (reference encoder/modules/conv.py:25-34, 115)
std::vector<float> fold_weight_norm(const float* g, const float* v, int cout, int cin, int k) {
    std::vector<float> w((size_t)cout * cin * k);
    for (int n = 0; n < cout; ++n) {
        const float* vn = v + (size_t)n * cin * k;
        double s = 0;
        for (int i = 0; i < cin * k; ++i) s += (double)vn[i] * vn[i];
        float norm = (float)std::sqrt(s);
        for (int i = 0; i < cin * k; ++i) w[(size_t)n * cin * k + i] = g[n] * vn[i] / norm;
    }
    return w;
}

ConvW load_wn_conv(wt_handle* h, const Table& t, const std::string& p, int cout, int cin, int k, int stride) {
    const float* g = t.get(p + "conv.conv.weight_g", cout);
    const float* v = t.get(p + "conv.conv.weight_v", (int64_t)cout * cin * k);
    const float* b = t.get(p + "conv.conv.bias", cout);
    auto w = fold_weight_norm(g, v, cout, cin, k);
    ConvW c;
    c.cout = cout; c.cin = cin; c.k = k; c.stride = stride;
    auto rl = relayout_conv(w.data(), cout, cin, k);
    c.w = h->upload(rl);
    if ((k * cin) % 64 == 0) { HalfW hw = h->upload_split(rl); c.w_hi = hw.hi; c.w_lo = hw.lo; }
    c.b = h->upload(std::vector<float>(b, b + cout));
    c.hw = rl;
    c.hb.assign(b, b + cout);
    return c;
}

ConvW load_conv(wt_handle* h, const Table& t, const std::string& p, int cout, int cin, int k) {
    const float* w = t.get(p + "weight", (int64_t)cout * cin * k);
    const float* b = t.get(p + "bias", cout);
    ConvW c;
    c.cout = cout; c.cin = cin; c.k = k; c.stride = 1;
    auto rl = relayout_conv(w, cout, cin, k);
    c.w = h->upload(rl);
    if (cin % 64 == 0) { HalfW hw = h->upload_split(rl); c.w_hi = hw.hi; c.w_lo = hw.lo; }
    c.b = h->upload(std::vector<float>(b, b + cout));
    return c;
}

float* load_vec(wt_handle* h, const Table& t, const std::string& name, int64_t n) {
    const float* p = t.get(name, n);
    return h->upload(std::vector<float>(p, p + n));
}

void prepare(wt_handle* h, const Table& t) {
    const wt_config& c = h->cfg;
    std::vector<float> h_conv0_w, h_conv0_b;  // host copies: the level-0 shortcut is composed with conv0
    const std::string E = "feature_extractor.encodec.encoder.model.";
    // ---- encoder (reference encoder/modules/seanet.py:105-141) ----
    {
        const int C = c.n_filters;
        const float* g = t.get(E + "0.conv.conv.weight_g", C);
        const float* v = t.get(E + "0.conv.conv.weight_v", (int64_t)C * 7);
        const float* b = t.get(E + "0.conv.conv.bias", C);
        h_conv0_w = fold_weight_norm(g, v, C, 1, 7);
        h_conv0_b.assign(b, b + C);
        h->conv0_w = h->upload(h_conv0_w);
        h->conv0_b = h->upload(h_conv0_b);
    }
    int ch = c.n_filters, idx = 1;
    for (int i = 0; i < 4; ++i) {
        std::string p = E + std::to_string(idx) + ".";
        h->rb[i].c1 = load_wn_conv(h, t, p + "block.1.", ch / 2, ch, 3, 1);
        h->rb[i].c2 = load_wn_conv(h, t, p + "block.3.", ch, ch / 2, 1, 1);
        h->rb[i].sc = load_wn_conv(h, t, p + "shortcut.", ch, ch, 1, 1);
        int s = c.strides[i];
        h->down[i] = load_wn_conv(h, t, E + std::to_string(idx + 2) + ".", 2 * ch, ch, 2 * s, s);
        ch *= 2;
        idx += 3;
    }
    if (ch != c.dimension) throw Error(WT_ERR_VALUE, "n_filters * 16 must equal dimension");
    for (int l = 0; l < c.lstm_layers; ++l) {
        std::string p = E + std::to_string(idx) + ".lstm.";
        std::string sfx = "_l" + std::to_string(l);
        const int64_t D = ch;
        h->lstm[l].w_ih = load_vec(h, t, p + "weight_ih" + sfx, 4 * D * D);
        h->lstm[l].w_hh = load_vec(h, t, p + "weight_hh" + sfx, 4 * D * D);
        const float* bi = t.get(p + "bias_ih" + sfx, 4 * D);
        const float* bh = t.get(p + "bias_hh" + sfx, 4 * D);
        std::vector<float> bias(4 * D);
        for (int64_t i = 0; i < 4 * D; ++i) bias[i] = bi[i] + bh[i];
        h->lstm[l].bias = h->upload(bias);
    }
    h->enc_last = load_wn_conv(h, t, E + std::to_string(idx + 2) + ".", c.dimension, ch, 7, 1);
    {   // tcgen05 encoder weight forms
        int C = c.n_filters;
        for (int i = 0; i < 4; ++i) {
            auto& rt = h->rb_tc[i];
            const ConvW& c1 = h->rb[i].c1;
            const ConvW& c2 = h->rb[i].c2;
            const ConvW& sc = h->rb[i].sc;
            const int K1 = 3 * C;
            // k-block widths that fit the operands exactly (gemm_tc.cuh `kw`): the k3 window of 3C elements and the
            // [C/2 | C] (level 0: [16 | 8 audio taps]) halves of the fused 1x1 convs
            // (measured: 32-byte rows and three 64-byte k-blocks per tile are slower than zero-filled 128-byte rows at
            //  level 0, profiles/r01_enc_chunk_s3_summary.md, so only the [32 | 64] halves of level 1 go narrow)
            rt.kw1 = 64;
            rt.Kp1 = (int)align_up(K1, rt.kw1);
            std::vector<float> w1((size_t)(C / 2) * rt.Kp1, 0.f);
            for (int n = 0; n < C / 2; ++n)
                for (int k = 0; k < K1; ++k) w1[(size_t)n * rt.Kp1 + k] = c1.hw[(size_t)n * K1 + k];
            rt.w1 = h->upload_split(w1);
            rt.kw2 = C / 2 == 32 ? 32 : 64;
            // shortcut operand: the raw-audio window at level 0; at level 1 the window of ELU(y0) the level-0 strided conv
            // reads (composed shortcut, below); x itself deeper down
            const int Cx = i == 0 ? 8 : (i == 1 ? 2 * c.strides[0] * (C / 2) : C);
            rt.kb0 = (C / 2 + rt.kw2 - 1) / rt.kw2;
            rt.kb1 = (Cx + rt.kw2 - 1) / rt.kw2;
            const int K2 = rt.kw2 * (rt.kb0 + rt.kb1);
            const int off1 = rt.kw2 * rt.kb0;  // first K column of the shortcut half
            std::vector<float> w2((size_t)C * K2, 0.f), b2(C);
            for (int n = 0; n < C; ++n) {
                for (int k = 0; k < C / 2; ++k) w2[(size_t)n * K2 + k] = c2.hw[(size_t)n * (C / 2) + k];
                if (i == 0) {
                    // level 0: shortcut(conv0(wav)) = (Wsc W0) * wav + Wsc b0 + b_sc, a k7 conv of the raw audio
                    // (7 taps, padded to 8), composed in fp64; its A operand is the audio window conv0 writes
                    double bacc = 0;
                    for (int j = 0; j < 7; ++j) {
                        double acc = 0;
                        for (int k = 0; k < C; ++k) acc += (double)sc.hw[(size_t)n * C + k] * h_conv0_w[(size_t)k * 7 + j];
                        w2[(size_t)n * K2 + off1 + j] = (float)acc;
                    }
                    for (int k = 0; k < C; ++k) bacc += (double)sc.hw[(size_t)n * C + k] * h_conv0_b[k];
                    b2[n] = (float)((double)c2.hb[n] + (double)sc.hb[n] + bacc);
                } else if (i == 1) {
                    // level 1: x1 = W_d * a + b_d with a the window of ELU(y0) planes the strided conv reads, so
                    // shortcut(x1) = (W_sc W_d) * a + W_sc b_d + b_sc reads those planes too (composed in fp64) and the
                    // strided conv no longer writes the raw-x planes (0.6 GB per 64 clips written and read back)
                    const ConvW& dn = h->down[0];
                    const int Kd = dn.k * dn.cin;
                    for (int j = 0; j < Kd; ++j) {
                        double acc = 0;
                        for (int k = 0; k < C; ++k) acc += (double)sc.hw[(size_t)n * C + k] * dn.hw[(size_t)k * Kd + j];
                        w2[(size_t)n * K2 + off1 + j] = (float)acc;
                    }
                    double bacc = 0;
                    for (int k = 0; k < C; ++k) bacc += (double)sc.hw[(size_t)n * C + k] * dn.hb[k];
                    b2[n] = (float)((double)c2.hb[n] + (double)sc.hb[n] + bacc);
                } else {
                    for (int k = 0; k < C; ++k) w2[(size_t)n * K2 + off1 + k] = sc.hw[(size_t)n * C + k];
                    b2[n] = c2.hb[n] + sc.hb[n];
                }
            }
            rt.w2 = h->upload_split(w2);
            rt.bias2 = h->upload(b2);
            if (i == 1 && C == 64 && (h->down[0].k * h->down[0].cin == 128 || h->down[0].k * h->down[0].cin == 256) &&
                enc_l1_fused_supported(C / 2, c.strides[0])) {
                // fused level 0 -> 1 kernel (enc_fused.cu): the three weight tiles as single fp16 tensors whose rows
                // stack the hi / lo planes in the order the MMAs address them
                const ConvW& dn = h->down[0];
                auto hi16 = [](float v) { return __float2half_rn(v); };
                auto lo16 = [](float v) { const __half hh = __float2half_rn(v); return __float2half_rn(v - __half2float(hh)); };
                const int K0 = dn.k * dn.cin;  // 128 (stride 2) or 256 (stride 4)
                std::vector<__half> p1((size_t)256 * K0), p2((size_t)192 * 64), p3((size_t)128 * 32);
                for (int n = 0; n < 64; ++n)
                    for (int j = 0; j < K0; ++j) {
                        const float wc = w2[(size_t)n * K2 + off1 + j], wd = dn.hw[(size_t)n * K0 + j];
                        p1[(size_t)(0 + n) * K0 + j] = hi16(wc);
                        p1[(size_t)(64 + n) * K0 + j] = hi16(wd);
                        p1[(size_t)(128 + n) * K0 + j] = lo16(wc);
                        p1[(size_t)(192 + n) * K0 + j] = lo16(wd);
                    }
                for (int tap = 0; tap < 3; ++tap)
                    for (int n = 0; n < 32; ++n)
                        for (int k = 0; k < 64; ++k) {
                            const float w = c1.hw[(size_t)n * 192 + tap * 64 + k];
                            p2[(size_t)(tap * 32 + n) * 64 + k] = hi16(w);
                            p2[(size_t)(96 + tap * 32 + n) * 64 + k] = lo16(w);
                        }
                for (int n = 0; n < 64; ++n)
                    for (int k = 0; k < 32; ++k) {
                        const float w = c2.hw[(size_t)n * 32 + k];
                        p3[(size_t)n * 32 + k] = hi16(w);
                        p3[(size_t)(64 + n) * 32 + k] = lo16(w);
                    }
                std::vector<float> fb(160);
                for (int n = 0; n < 64; ++n) fb[n] = dn.hb[n];
                for (int n = 0; n < 32; ++n) fb[64 + n] = c1.hb[n];
                for (int n = 0; n < 64; ++n) fb[96 + n] = b2[n];
                h->l1_fused.w1 = h->upload_halves(p1);
                h->l1_fused.w2 = h->upload_halves(p2);
                h->l1_fused.w3 = h->upload_halves(p3);
                h->l1_fused.bias = h->upload(fb);
                h->l1_fused.k0 = K0;
            }
            if (i == 0 && C == 32) {
                // fused level-0 kernel: w0t[7][32] b0[32] w1t[96][16] b1[16] w2t[16][32] wsct[8][32] b2[32]
                std::vector<float> pk((size_t)resblock0_pack_floats(), 0.f);
                float* q = pk.data();
                for (int j = 0; j < 7; ++j)
                    for (int n = 0; n < 32; ++n) q[j * 32 + n] = h_conv0_w[(size_t)n * 7 + j];
                q += 224;
                for (int n = 0; n < 32; ++n) q[n] = h_conv0_b[n];
                q += 32;
                for (int k = 0; k < 96; ++k)
                    for (int n = 0; n < 16; ++n) q[k * 16 + n] = c1.hw[(size_t)n * 96 + k];
                q += 1536;
                for (int n = 0; n < 16; ++n) q[n] = c1.hb[n];
                q += 16;
                for (int k = 0; k < 16; ++k)
                    for (int n = 0; n < 32; ++n) q[k * 32 + n] = c2.hw[(size_t)n * 16 + k];
                q += 512;
                for (int j = 0; j < 7; ++j)
                    for (int n = 0; n < 32; ++n) q[j * 32 + n] = w2[(size_t)n * K2 + off1 + j];
                q += 256;
                for (int n = 0; n < 32; ++n) q[n] = b2[n];
                h->rb0_pack = h->upload(pk);
                {   // tensor-core level 0 (enc_l0_tc.cu): MMA weight tiles as stacked hi / lo rows, the rest as constants
                    auto hi16 = [](float v) { return __float2half_rn(v); };
                    auto lo16 = [](float v) { const __half hh = __float2half_rn(v); return __float2half_rn(v - __half2float(hh)); };
                    std::vector<__half> pk3((size_t)96 * 32), p11((size_t)64 * 16);
                    for (int tap = 0; tap < 3; ++tap)
                        for (int n = 0; n < 16; ++n)
                            for (int k = 0; k < 32; ++k) {
                                const float w = c1.hw[(size_t)n * 96 + tap * 32 + k];
                                pk3[(size_t)(tap * 16 + n) * 32 + k] = hi16(w);
                                pk3[(size_t)(48 + tap * 16 + n) * 32 + k] = lo16(w);
                            }
                    for (int n = 0; n < 32; ++n)
                        for (int k = 0; k < 16; ++k) {
                            const float w = c2.hw[(size_t)n * 16 + k];
                            p11[(size_t)n * 16 + k] = hi16(w);
                            p11[(size_t)(32 + n) * 16 + k] = lo16(w);
                        }
                    h->l0_consts.assign(528, 0.f);
                    float* kc = h->l0_consts.data();
                    for (int j = 0; j < 7; ++j)
                        for (int n = 0; n < 32; ++n) {
                            kc[j * 32 + n] = h_conv0_w[(size_t)n * 7 + j];
                            kc[256 + j * 32 + n] = w2[(size_t)n * K2 + off1 + j];
                        }
                    for (int n = 0; n < 32; ++n) { kc[224 + n] = h_conv0_b[n]; kc[496 + n] = b2[n]; }
                    for (int n = 0; n < 16; ++n) kc[480 + n] = c1.hb[n];
                    h->l0_tc.wk3 = h->upload_halves(pk3);
                    h->l0_tc.w1x1 = h->upload_halves(p11);
                    h->l0_tc.consts = h->l0_consts.data();
                }
            }
            C *= 2;
        }
        const int D = c.dimension;
        if (D % 16 == 0) {
            for (int l = 0; l < c.lstm_layers; ++l) {
                std::string p = E + std::to_string(idx) + ".lstm.";
                std::string sfx = "_l" + std::to_string(l);
                const float* wi = t.get(p + "weight_ih" + sfx, (int64_t)4 * D * D);
                const float* wh = t.get(p + "weight_hh" + sfx, (int64_t)4 * D * D);
                const float* bi = t.get(p + "bias_ih" + sfx, 4 * D);
                const float* bh = t.get(p + "bias_hh" + sfx, 4 * D);
                std::vector<float> pwi((size_t)4 * D * D), pwh((size_t)4 * D * D), pb((size_t)4 * D);
                for (int np = 0; np < 4 * D; ++np) {
                    const int nt = np / 64, gate = (np % 64) / 16, j = np % 16;
                    const int src = gate * D + nt * 16 + j;
                    std::memcpy(&pwi[(size_t)np * D], wi + (size_t)src * D, D * sizeof(float));
                    std::memcpy(&pwh[(size_t)np * D], wh + (size_t)src * D, D * sizeof(float));
                    pb[np] = bi[src] + bh[src];
                }
                h->lstm_tc[l].w_ih = h->upload_split(pwi);
                h->lstm_tc[l].w_hh = h->upload_split(pwh);
                h->lstm_tc[l].bias = h->upload(pb);
            }
        }
    }

    // ---- SEANet decoder, when the checkpoint carries it (reference encoder/modules/seanet.py:189-238) ----
    {
        const std::string Dp = "feature_extractor.encodec.decoder.model.";
        if (t.has(Dp + "0.conv.conv.weight_v")) {
            auto& sd = h->sdec;
            int chd = c.n_filters * 16;
            sd.first = load_wn_conv(h, t, Dp + "0.", chd, c.dimension, 7, 1);
            for (int l = 0; l < c.lstm_layers; ++l) {
                const std::string p = Dp + "1.lstm.";
                const std::string sfx = "_l" + std::to_string(l);
                const int64_t D = chd;
                sd.lstm[l].w_ih = load_vec(h, t, p + "weight_ih" + sfx, 4 * D * D);
                sd.lstm[l].w_hh = load_vec(h, t, p + "weight_hh" + sfx, 4 * D * D);
                const float* bi = t.get(p + "bias_ih" + sfx, 4 * D);
                const float* bh = t.get(p + "bias_hh" + sfx, 4 * D);
                std::vector<float> bias(4 * D);
                for (int64_t i = 0; i < 4 * D; ++i) bias[i] = bi[i] + bh[i];
                sd.lstm[l].bias = h->upload(bias);
            }
            int didx = 2;
            for (int i = 0; i < 4; ++i) {
                const int s_ = c.strides[3 - i], k = 2 * s_, cin = chd, cout = chd / 2;
                const std::string p = Dp + std::to_string(didx + 1) + ".convtr.convtr.";
                const float* g = t.get(p + "weight_g", cin);
                const float* v = t.get(p + "weight_v", (int64_t)cin * cout * k);
                const float* b = t.get(p + "bias", cout);
                // weight_norm over dim 0 (input channels), then GEMM form W2[j*cout + co][ci] = w[ci][co][j]
                std::vector<float> w2((size_t)k * cout * cin);
                for (int ci = 0; ci < cin; ++ci) {
                    const float* vi = v + (size_t)ci * cout * k;
                    double ss = 0;
                    for (int e = 0; e < cout * k; ++e) ss += (double)vi[e] * vi[e];
                    const float sc = g[ci] / (float)std::sqrt(ss);
                    for (int co = 0; co < cout; ++co)
                        for (int j = 0; j < k; ++j) w2[((size_t)j * cout + co) * cin + ci] = sc * vi[(size_t)co * k + j];
                }
                sd.up[i].w = h->upload(w2);
                sd.up[i].b = h->upload(std::vector<float>(b, b + cout));
                sd.up[i].cin = cin; sd.up[i].cout = cout; sd.up[i].stride = s_;
                const std::string rp = Dp + std::to_string(didx + 2) + ".";
                sd.rb[i].c1 = load_wn_conv(h, t, rp + "block.1.", cout / 2, cout, 3, 1);
                sd.rb[i].c2 = load_wn_conv(h, t, rp + "block.3.", cout, cout / 2, 1, 1);
                sd.rb[i].sc = load_wn_conv(h, t, rp + "shortcut.", cout, cout, 1, 1);
                chd = cout;
                didx += 3;
            }
            sd.last = load_wn_conv(h, t, Dp + std::to_string(didx + 1) + ".", 1, chd, 7, 1);
            sd.present = true;
        }
    }

    // ---- codebooks (reference encoder/quantization/core_vq.py:126-133) ----
    {
        const int64_t n = (int64_t)c.vq_bins * c.dimension;
        std::vector<float> all((size_t)c.num_quantizers * n);
        for (int q = 0; q < c.num_quantizers; ++q) {
            const float* e = t.get("feature_extractor.encodec.quantizer.vq.layers." + std::to_string(q) +
                                       "._codebook.embed", n);
            std::memcpy(all.data() + (size_t)q * n, e, n * sizeof(float));
        }
        h->codebooks = h->upload(all);
        std::vector<float> cn(c.vq_bins);
        for (int j = 0; j < c.vq_bins; ++j) {
            double s = 0;
            for (int d = 0; d < c.dimension; ++d) { double v = all[(size_t)j * c.dimension + d]; s += v * v; }
            cn[j] = (float)s;
        }
        h->cnorm = h->upload(cn);
        const int D0 = c.dimension;
        std::vector<double> mu(D0, 0.0);
        for (int j = 0; j < c.vq_bins; ++j)
            for (int d = 0; d < D0; ++d) mu[d] += all[(size_t)j * D0 + d];
        std::vector<float> muf(D0), cc((size_t)c.vq_bins * D0), cnc(c.vq_bins);
        for (int d = 0; d < D0; ++d) muf[d] = (float)(mu[d] / c.vq_bins);
        for (int j = 0; j < c.vq_bins; ++j) {
            double s2 = 0;
            for (int d = 0; d < D0; ++d) {
                float v = all[(size_t)j * D0 + d] - muf[d];
                cc[(size_t)j * D0 + d] = v;
                s2 += (double)v * v;
            }
            cnc[j] = (float)s2;
        }
        h->cb_mean = h->upload(muf);
        h->cb_c = h->upload_split(cc);
        h->cb_h = h->upload_split(std::vector<float>(all.begin(), all.begin() + n));
        h->cnorm_c = h->upload(cnc);
    }

    // ---- backbone (reference decoder/models.py:166-216) ----
    const int D = c.dim, H = c.intermediate_dim, NE = c.adanorm_num_embeddings;
    h->embed = load_conv(h, t, "backbone.embed.", D, c.dimension, 7);
    int slot = 0;
    for (int i : {0, 1, 3, 4}) {
        std::string p = "backbone.pos_net." + std::to_string(i) + ".";
        auto& r = h->pos[slot++];
        r.n1w = load_vec(h, t, p + "norm1.weight", D); r.n1b = load_vec(h, t, p + "norm1.bias", D);
        r.n2w = load_vec(h, t, p + "norm2.weight", D); r.n2b = load_vec(h, t, p + "norm2.bias", D);
        r.c1 = load_conv(h, t, p + "conv1.", D, D, 3);
        r.c2 = load_conv(h, t, p + "conv2.", D, D, 3);
    }
    {
        std::string p = "backbone.pos_net.2.";
        h->attn.nw = load_vec(h, t, p + "norm.weight", D);
        h->attn.nb = load_vec(h, t, p + "norm.bias", D);
        std::vector<float> w((size_t)3 * D * D), b((size_t)3 * D);
        int o = 0;
        for (const char* nm : {"q", "k", "v"}) {
            const float* wp = t.get(p + nm + ".weight", (int64_t)D * D);
            const float* bp = t.get(p + nm + ".bias", D);
            std::memcpy(w.data() + (size_t)o * D * D, wp, (size_t)D * D * sizeof(float));
            std::memcpy(b.data() + (size_t)o * D, bp, (size_t)D * sizeof(float));
            ++o;
        }
        h->attn.wqkv = h->upload(w);
        h->attn.wqkv_h = h->upload_split(w);
        h->attn.bqkv = h->upload(b);
        h->attn.proj = load_conv(h, t, p + "proj_out.", D, D, 1);
    }
    h->gn5w = load_vec(h, t, "backbone.pos_net.5.weight", D);
    h->gn5b = load_vec(h, t, "backbone.pos_net.5.bias", D);
    h->norm_scale = load_vec(h, t, "backbone.norm.scale.weight", (int64_t)NE * D);
    h->norm_shift = load_vec(h, t, "backbone.norm.shift.weight", (int64_t)NE * D);
    h->cnx.resize(c.num_layers);
    for (int i = 0; i < c.num_layers; ++i) {
        std::string p = "backbone.convnext." + std::to_string(i) + ".";
        auto& x = h->cnx[i];
        {   // depthwise taps transposed to [tap][channel] so that lanes read consecutive channels
            const float* wp = t.get(p + "dwconv.weight", (int64_t)D * 7);
            std::vector<float> wt((size_t)7 * D);
            for (int cc = 0; cc < D; ++cc)
                for (int j = 0; j < 7; ++j) wt[(size_t)j * D + cc] = wp[(size_t)cc * 7 + j];
            x.dw = h->upload(wt);
        }
        x.db = load_vec(h, t, p + "dwconv.bias", D);
        x.scale = load_vec(h, t, p + "norm.scale.weight", (int64_t)NE * D);
        x.shift = load_vec(h, t, p + "norm.shift.weight", (int64_t)NE * D);
        x.w1 = load_vec(h, t, p + "pwconv1.weight", (int64_t)H * D);
        { const float* wp = t.get(p + "pwconv1.weight", (int64_t)H * D); x.w1_h = h->upload_split(std::vector<float>(wp, wp + (size_t)H * D)); }
        { const float* wp = t.get(p + "pwconv2.weight", (int64_t)D * H); x.w2_h = h->upload_split(std::vector<float>(wp, wp + (size_t)D * H)); }
        x.b1 = load_vec(h, t, p + "pwconv1.bias", H);
        x.w2 = load_vec(h, t, p + "pwconv2.weight", (int64_t)D * H);
        x.b2 = load_vec(h, t, p + "pwconv2.bias", D);
        x.gamma = load_vec(h, t, p + "gamma", D);
    }
    h->fln_w = load_vec(h, t, "backbone.final_layer_norm.weight", D);
    h->fln_b = load_vec(h, t, "backbone.final_layer_norm.bias", D);

    // ---- head + windowed inverse real DFT basis (reference decoder/heads.py:36-40,
    //      decoder/spectral_ops.py:56-57: irfft(norm="backward") * window) ----
    const int N = c.n_fft, half = N / 2 + 1;
    h->head_w = load_vec(h, t, "head.out.weight", (int64_t)(N + 2) * D);
    { const float* wp = t.get("head.out.weight", (int64_t)(N + 2) * D); h->head_h = h->upload_split(std::vector<float>(wp, wp + (size_t)(N + 2) * D)); }
    h->ldz = (int)align_up(N + 2, 4);
    h->head_b = load_vec(h, t, "head.out.bias", N + 2);
    const float* win = t.get("head.istft.window", N);
    h->Kp = (int)align_up(2 * half, 64);
    std::vector<float> basis((size_t)N * h->Kp, 0.f), wsq(N);
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < N; ++n) {
        wsq[n] = win[n] * win[n];
        for (int k = 0; k < half; ++k) {
            double wk = (k == 0 || k == N / 2) ? 1.0 : 2.0;
            long long r = ((long long)k * n) % N;
            double ang = two_pi * (double)r / (double)N;
            basis[(size_t)n * h->Kp + k] = (float)(wk * std::cos(ang) / N * (double)win[n]);
            basis[(size_t)n * h->Kp + half + k] = (float)(-wk * std::sin(ang) / N * (double)win[n]);
        }
    }
    h->basis = h->upload(basis);
    h->basis_h = h->upload_split(basis);
    h->wsq = h->upload(wsq);
    WT_CUDA(cudaMalloc(&h->err_flag, sizeof(int)));
    WT_CUDA(cudaMemset(h->err_flag, 0, sizeof(int)));
    {
        void* z = nullptr;
        const size_t bytes = 2 * (size_t)ENC_GROUP * c.dimension * sizeof(__half);  // hi and lo zero planes
        WT_CUDA(cudaMalloc(&z, bytes));
        WT_CUDA(cudaMemset(z, 0, bytes));
        h->owned.push_back(z);
        h->zero_rows = reinterpret_cast<float*>(z);
    }
    WT_CUDA(cudaMallocHost(&h->err_host, sizeof(int)));
}

// ---------------------------------------------------------------------------------------
// shape helpers
// ---------------------------------------------------------------------------------------
int frames_for(const wt_config& c, int T) {
    long long n = T;
    for (int i = 0; i < 4; ++i) n = (n + c.strides[i] - 1) / c.strides[i];
    return (int)n;
}

size_t enc_front_floats(const wt_config& c, int Bc, int T) {
    size_t tot = 0;
    size_t Tc = T, C = c.n_filters;
    auto a = [&](size_t n) { tot += align_up(n * sizeof(float), 256) / sizeof(float); };
    a((size_t)Bc * Tc * C);
    for (int i = 0; i < 4; ++i) {
        a((size_t)Bc * Tc * (C / 2));
        a((size_t)Bc * Tc * C);
        a((size_t)Bc * Tc * C);
        size_t Tn = (Tc + c.strides[i] - 1) / c.strides[i];
        a((size_t)Bc * Tn * 2 * C);
        Tc = Tn; C *= 2;
    }
    return tot;
}

// tcgen05 encoder front: planes of x / ELU(x) (pitch T+2), ELU(h1) (pitch T+2), ELU(y) (pitch (T'+1)*s)
size_t enc_front_tc_floats(const wt_config& c, int Bc, int T) {
    size_t tot = 0;
    size_t Tc = T, C = c.n_filters;
    auto a = [&](size_t halves) { tot += align_up(((halves + 1) / 2) * sizeof(float), 256) / sizeof(float); };
    for (int i = 0; i < 4; ++i) {
        const size_t s_ = c.strides[i], Tn = (Tc + s_ - 1) / s_;
        for (int k = 0; k < 4; ++k) a((size_t)Bc * (Tc + 2) * C);
        for (int k = 0; k < 2; ++k) a((size_t)Bc * (Tc + 2) * (C / 2));
        for (int k = 0; k < 2; ++k) a((size_t)Bc * (Tn + 2) * s_ * C);  // (level 0 uses Tn + 2 window slots per clip)
        tot += align_up((size_t)Bc * (Tc + 2) * C * sizeof(float), 256) / sizeof(float);  // optional fp32 tap copies
        tot += align_up((size_t)Bc * (Tn + 2) * 2 * C * sizeof(float), 256) / sizeof(float);
        Tc = Tn; C *= 2;
    }
    return tot;
}

size_t enc_back_floats(const wt_config& c, int Bg, int L) {
    size_t tot = 0;
    auto a = [&](size_t n) { tot += align_up(n * sizeof(float), 256) / sizeof(float); };
    const size_t M = (size_t)Bg * L, D = c.dimension;
    a(M * D);                                          // pre-LSTM rows of the group
    a(M * 4 * D);                                      // xin (layer 1)
    a(M * 4 * D);                                      // xin (layer 2: the two layers overlap as a wavefront)
    a((size_t)Bg * D);                                 // c (layer 2)
    a(lstm_counter_ints(Bg, L) + 64);                  // counters (layer 2)
    for (int l = 0; l < c.lstm_layers; ++l) a(M * D);  // y_l
    a((size_t)Bg * 4 * D);                             // gates
    a((size_t)Bg * D);                                 // c
    a(M * D);                                          // lstm + skip
    a(M * D);                                          // z
    a(M * D);                                          // split planes of the pre-LSTM rows (tcgen05 plan)
    for (int l = 0; l < c.lstm_layers; ++l) a(M * D);  // split planes of y_l
    a((size_t)Bg * (L + 6) * D);                       // ELU(lstm + skip) planes, reflect-padded
    a(M * D);                                          // centred z planes for the tcgen05 VQ
    a(M * 2);                                          // packed (distance, index) keys
    a(lstm_counter_ints(Bg, L) + 64);                  // per-(batch tile, step, k-block) arrival counters of the LSTM
    return tot;
}

size_t dec_chunk_floats(const wt_config& c, int Bc, int L, int Kp) {
    // covers both plans (the tcgen05 plan adds 3 halo rows per clip and the split-fp16 operand planes)
    size_t tot = 0;
    auto a = [&](size_t n) { tot += align_up(n * sizeof(float), 256) / sizeof(float); };
    const size_t R = (size_t)Bc * (L + 3);
    const size_t ldz = align_up(c.n_fft + 2, 4);
    size_t big = std::max<size_t>(std::max<size_t>(3 * c.dim, c.intermediate_dim), std::max<size_t>(ldz, Kp));
    a(R * c.dimension);                          // xin (fp32 or both fp16 planes)
    a(R * c.dim); a(R * c.dim); a(R * c.dim);    // x, t1/t2, a planes
    a(R * big); a(R * big);                      // qkv / z / frames ; S or GELU planes
    a(R * Kp);                                   // S planes (tcgen05 plan)
    const size_t Lpad = align_up(L, 64);
    a(R * Lpad); a(R * Lpad);                    // attention scores (fp32) and probability planes
    a((size_t)Bc * c.dim * Lpad);                // V^T planes
    return tot;
}

size_t workspace_bytes(const wt_handle* h, int B, int T) {
    const wt_config& c = h->cfg;
    int L = frames_for(c, T);
    size_t e = enc_back_floats(c, std::min(B, ENC_GROUP), L) +
               std::max(enc_front_floats(c, std::min(B, ENC_CHUNK), T),
                        enc_front_tc_floats(c, std::min(B, ENC_CHUNK), T));
    size_t d = dec_chunk_floats(c, std::min(B, DEC_CHUNK), L, h->Kp);
    return std::max(e, d) * sizeof(float) + 4096;
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
enum Cat : int { CAT_ENC_CONV = 0, CAT_LSTM, CAT_VQ, CAT_DEC_CONV, CAT_PWCONV, CAT_HEAD, CAT_ATTN, CAT_MEM, CAT_COUNT };

// Kernel ids of the per-kernel timing record (wt_timing_read_kernel). tcgen05 GEMM variants report BN * 10 + passes
// through last_launch_info(); the other kernels of the step are named here.
enum Kern : int {
    KERN_LSTM = 1, KERN_RB0 = 2, KERN_GROUPNORM = 3, KERN_DWCONV_LN = 4, KERN_LAYERNORM = 5, KERN_SPECTRAL = 6,
    KERN_OLA = 7, KERN_SOFTMAX = 8, KERN_VT = 9, KERN_ROWS = 10, KERN_GATHER = 11, KERN_LSTM_SKIP = 12, KERN_VQ_MISC = 13,
    KERN_ENC_L1F = 14, KERN_ENC_L0TC = 15
};

// Counts one kernel launch and, when timing is on, brackets it with CUDA events on its stream. `kern` / `flops` /
// `bytes` name the kernel and its ALGORITHMIC work (compulsory bytes for the memory-bound kernels).
struct Scope {
    wt_handle* h;
    cudaStream_t s;
    int idx = -1;
    int kern;
    double flops, bytes;
    Scope(wt_handle* h_, int cat, cudaStream_t s_, int kern_ = 0, double flops_ = 0, double bytes_ = 0)
        : h(h_), s(s_), kern(kern_), flops(flops_), bytes(bytes_) {
        ++h->launches;
        if (h->timing) {
            wt_handle::Ev e{h->get_event(), h->get_event(), cat};
            last_launch_info() = LaunchInfo{};
            cudaEventRecord(e.a, s);
            h->evs.push_back(e);
            idx = (int)h->evs.size() - 1;
        }
    }
    ~Scope() {
        if (idx >= 0) {
            cudaEventRecord(h->evs[idx].b, s);
            // the tcgen05 GEMM launcher reports its variant and FLOPs; other kernels are named by the caller
            h->evs[idx].kern = kern ? kern : last_launch_info().kern;
            h->evs[idx].flops = kern ? flops : last_launch_info().flops;
            h->evs[idx].bytes = bytes;
        }
    }
};

struct Runner {
    wt_handle* h;
    cudaStream_t s;
    int cat = CAT_MEM;

    void gemm(const TapGemm& g) {
        Scope sc(h, cat, s);
        launch_tap_gemm_simt(g, s);
    }

    // Conv1d as tap-GEMM on channels-last rows. `reflect`: SConv1d non-causal padding rule
    // (reference encoder/modules/conv.py:195-211); otherwise symmetric zero padding (k-1)/2.
    void conv(const ConvW& w, const float* in, float* out, int Bc, int Tin, bool reflect, int pro, int act = ACT_NONE,
              const float* res = nullptr, const float* gamma = nullptr) {
        TapGemm g;
        const int s_ = w.stride, k = w.k;
        int left, right_total;
        int Tout;
        if (reflect) {
            int pt = k - s_;
            int right = pt / 2;
            left = pt - right;
            Tout = (Tin + s_ - 1) / s_;
            int extra = Tout * s_ - Tin;
            right_total = right + extra;
        } else {
            left = (k - 1) / 2;
            right_total = left;
            Tout = Tin;
        }
        g.A = in; g.W = w.w; g.bias = w.b; g.gamma = gamma; g.res = res; g.out = out;
        g.M = Bc * Tout; g.N = w.cout; g.K = k * w.cin;
        g.Cin = w.cin; g.taps = k; g.stride = s_; g.pad_left = left;
        g.Tin = Tin; g.Tout = Tout;
        int max_pad = std::max(left, right_total);
        g.Trefl = std::max(Tin, max_pad + 1);
        g.pad_mode = reflect ? PAD_REFLECT : PAD_ZERO;
        g.pro = pro; g.act = act;
        g.lda = w.cin; g.ldo = w.cout; g.ldres = w.cout;
        gemm(g);
    }

    void linear(const float* A, const float* W, const float* bias, float* out, long long M, int N, int K, int act,
                const float* gamma, const float* res, int lda = 0, int ldo = 0, int ldres = 0) {
        TapGemm g;
        g.A = A; g.W = W; g.bias = bias; g.gamma = gamma; g.res = res; g.out = out;
        g.M = (int)M; g.N = N; g.K = K; g.Cin = K; g.taps = 1; g.stride = 1; g.pad_left = 0;
        g.Tin = 1; g.Tout = 1; g.Trefl = 1; g.pad_mode = PAD_ZERO; g.pro = PRO_NONE; g.act = act;
        g.lda = lda ? lda : K; g.ldo = ldo ? ldo : N; g.ldres = ldres ? ldres : N;
        gemm(g);
    }
};

// ---------------------------------------------------------------------------------------
// encoder: SEANetEncoder.forward on a chunk (reference encoder/modules/seanet.py:143-144)
// returns z rows [Bc*L, D] (channels-last)
// ---------------------------------------------------------------------------------------
// front: conv0 + 4 x (ResBlock, ELU, strided conv) on a chunk of clips; the last strided conv
// writes its [Bc*L, D] rows into `pre` (a slice of the group-wide pre-LSTM buffer).
void encoder_front(wt_handle* h, const float* wav, int Bc, int T, int b0, float* pre, cudaStream_t s) {
    const wt_config& c = h->cfg;
    Runner r{h, s};
    int Tc = T, C = c.n_filters;
    float* cur = h->alloc((size_t)Bc * Tc * C);
    r.cat = CAT_ENC_CONV;
    { Scope sc(h, CAT_ENC_CONV, s); launch_conv0(wav, h->conv0_w, h->conv0_b, cur, Bc, T, C, s); }
    h->tap("enc0", cur, Bc, Tc, C, b0, s);
    int idx = 1;
    for (int i = 0; i < 4; ++i) {
        // SEANetResnetBlock (seanet.py:45-63): shortcut(x) + conv1x1(ELU(conv_k3(ELU(x))))
        float* h1 = h->alloc((size_t)Bc * Tc * (C / 2));
        float* sc = h->alloc((size_t)Bc * Tc * C);
        float* y = h->alloc((size_t)Bc * Tc * C);
        r.conv(h->rb[i].c1, cur, h1, Bc, Tc, true, PRO_ELU);
        r.conv(h->rb[i].sc, cur, sc, Bc, Tc, true, PRO_NONE);
        r.conv(h->rb[i].c2, h1, y, Bc, Tc, true, PRO_ELU, ACT_NONE, sc);
        h->tap(("enc" + std::to_string(idx)).c_str(), y, Bc, Tc, C, b0, s);
        // ELU + strided SConv1d (seanet.py:121-131)
        int Tn = (Tc + c.strides[i] - 1) / c.strides[i];
        float* z = (i == 3) ? pre : h->alloc((size_t)Bc * Tn * 2 * C);
        r.conv(h->down[i], y, z, Bc, Tc, true, PRO_ELU);
        h->tap(("enc" + std::to_string(idx + 2)).c_str(), z, Bc, Tn, 2 * C, b0, s);
        cur = z; Tc = Tn; C *= 2; idx += 3;
    }
}

// back: SLSTM (reference encoder/modules/lstm.py:31-39) + ELU + final k7 conv on a whole group of
// clips at once, so that each recurrent step is one [Bg, D] x [D, 4D] contraction.
float* encoder_back(wt_handle* h, const float* pre, int Bg, int L, int b0, cudaStream_t s) {
    const wt_config& c = h->cfg;
    Runner r{h, s};
    const int D = c.dimension;
    const long long M = (long long)Bg * L;
    float* xin = h->alloc((size_t)M * 4 * D);
    float* ybuf[4];
    for (int l = 0; l < c.lstm_layers; ++l) ybuf[l] = h->alloc((size_t)M * D);
    float* gates = h->alloc((size_t)Bg * 4 * D);
    float* cst = h->alloc((size_t)Bg * D);
    const float* lin = pre;
    r.cat = CAT_LSTM;
    for (int l = 0; l < c.lstm_layers; ++l) {
        r.linear(lin, h->lstm[l].w_ih, h->lstm[l].bias, xin, M, 4 * D, D, ACT_NONE, nullptr, nullptr);
        WT_CUDA(cudaMemsetAsync(cst, 0, (size_t)Bg * D * sizeof(float), s));
        for (int t = 0; t < L; ++t) {
            const float* g_in;
            long long ldg;
            if (t == 0) {
                g_in = xin; ldg = (long long)L * 4 * D;  // h_{-1} = 0
            } else {
                r.linear(ybuf[l] + (size_t)(t - 1) * D, h->lstm[l].w_hh, nullptr, gates, Bg, 4 * D, D, ACT_NONE, nullptr,
                         xin + (size_t)t * 4 * D, L * D, 4 * D, L * 4 * D);
                g_in = gates; ldg = 4 * D;
            }
            { Scope sc(h, CAT_LSTM, s); launch_lstm_pointwise(g_in, cst, ybuf[l] + (size_t)t * D, Bg, D, ldg, (long long)L * D, s); }
        }
        lin = ybuf[l];
    }
    float* lo = h->alloc((size_t)M * D);
    { Scope sc(h, CAT_LSTM, s); launch_add(lin, pre, lo, M * D, s); }
    r.cat = CAT_ENC_CONV;
    h->tap("enc13", lo, Bg, L, D, b0, s);
    float* z = h->alloc((size_t)M * D);
    r.conv(h->enc_last, lo, z, Bg, L, true, PRO_ELU);
    h->tap("enc15", z, Bg, L, D, b0, s);
    return z;
}

// SEANetDecoder.forward (reference encoder/modules/seanet.py:189-238) on z [B, dimension, L] -> audio [B, L * hop]
// (fp32 CUDA-core kernels: this entry is next to the hot path, not on it).
size_t seanet_decoder_floats(const wt_config& c, int B, int L) {
    size_t per_clip = 0;
    long long T = L;
    int ch = c.n_filters * 16;
    per_clip += (size_t)L * (c.dimension + 5 * ch + 4 * ch) + (size_t)(4 + 1) * ch;  // rows, conv, LSTM y (x2) / skip / xin, gates, cell
    for (int i = 0; i < 4; ++i) {
        const int s_ = c.strides[3 - i];
        per_clip += (size_t)T * 2 * s_ * (ch / 2);                 // per-tap products
        T *= s_;
        ch /= 2;
        per_clip += (size_t)T * ch * 4 + (size_t)T * (ch / 2);     // y, shortcut, block output, next x; hidden
    }
    per_clip += (size_t)T;
    return per_clip * (size_t)B + 64 * 64;
}

void seanet_decoder(wt_handle* h, const float* z_bcl, int B, int L, float* audio, cudaStream_t s) {
    const wt_config& c = h->cfg;
    const auto& sd = h->sdec;
    if (!sd.present)
        throw Error(WT_ERR_RUNTIME, "seanet_decoder: the checkpoint holds no feature_extractor.encodec.decoder weights");
    if (B <= 0 || L <= 0) throw Error(WT_ERR_VALUE, "seanet_decoder: expected z [B, C, L] with B, L > 0");
    h->ensure_arena(seanet_decoder_floats(c, B, L) * sizeof(float) + (1 << 20));
    h->arena_off = 0;
    Runner r{h, s};
    r.cat = CAT_ENC_CONV;
    int ch = c.n_filters * 16;
    const long long M = (long long)B * L;
    float* zr = h->alloc((size_t)M * c.dimension);
    { Scope sc(h, CAT_MEM, s); launch_transpose_bcl_to_blc(z_bcl, zr, B, c.dimension, L, s); }
    float* x = h->alloc((size_t)M * ch);
    r.conv(sd.first, zr, x, B, L, true, PRO_NONE);
    {   // SLSTM (reference encoder/modules/lstm.py:31-39): rows are clip-major [b*L + t]
        const int D = ch;
        float* xin = h->alloc((size_t)M * 4 * D);
        float* ybuf[4];
        for (int l = 0; l < c.lstm_layers; ++l) ybuf[l] = h->alloc((size_t)M * D);
        float* gates = h->alloc((size_t)B * 4 * D);
        float* cst = h->alloc((size_t)B * D);
        const float* lin = x;
        r.cat = CAT_LSTM;
        for (int l = 0; l < c.lstm_layers; ++l) {
            r.linear(lin, sd.lstm[l].w_ih, sd.lstm[l].bias, xin, M, 4 * D, D, ACT_NONE, nullptr, nullptr);
            WT_CUDA(cudaMemsetAsync(cst, 0, (size_t)B * D * sizeof(float), s));
            for (int t = 0; t < L; ++t) {
                const float* g_in;
                long long ldg;
                if (t == 0) {
                    g_in = xin; ldg = (long long)L * 4 * D;  // h_{-1} = 0
                } else {
                    r.linear(ybuf[l] + (size_t)(t - 1) * D, sd.lstm[l].w_hh, nullptr, gates, B, 4 * D, D, ACT_NONE, nullptr,
                             xin + (size_t)t * 4 * D, L * D, 4 * D, L * 4 * D);
                    g_in = gates; ldg = 4 * D;
                }
                { Scope sc(h, CAT_LSTM, s); launch_lstm_pointwise(g_in, cst, ybuf[l] + (size_t)t * D, B, D, ldg, (long long)L * D, s); }
            }
            lin = ybuf[l];
        }
        float* lo = h->alloc((size_t)M * D);
        { Scope sc(h, CAT_LSTM, s); launch_add(lin, x, lo, M * D, s); }
        x = lo;
        r.cat = CAT_ENC_CONV;
    }
    int T = L;
    for (int i = 0; i < 4; ++i) {
        const auto& up = sd.up[i];
        const int s_ = up.stride, k = 2 * s_;
        // ELU + SConvTranspose1d: per-tap products as one GEMM (ELU in the loader), then the overlap gather
        float* g2 = h->alloc((size_t)B * T * k * up.cout);
        {
            TapGemm g;
            g.A = x; g.W = up.w; g.bias = nullptr; g.out = g2;
            g.M = B * T; g.N = k * up.cout; g.K = up.cin; g.Cin = up.cin; g.taps = 1; g.stride = 1; g.pad_left = 0;
            g.Tin = 1; g.Tout = 1; g.Trefl = 1; g.pad_mode = PAD_ZERO; g.pro = PRO_ELU; g.act = ACT_NONE;
            g.lda = up.cin; g.ldo = k * up.cout; g.ldres = k * up.cout;
            r.gemm(g);
        }
        const int total = k - s_, right = total / 2, left = total - right;
        const int Tn = T * s_;
        float* y = h->alloc((size_t)B * Tn * up.cout);
        { Scope sc(h, CAT_MEM, s); launch_convtr_gather(g2, up.b, y, B, T, up.cout, s_, left, s); }
        // SEANetResnetBlock (seanet.py:45-63)
        const int C = up.cout;
        float* h1 = h->alloc((size_t)B * Tn * (C / 2));
        float* scb = h->alloc((size_t)B * Tn * C);
        float* yo = h->alloc((size_t)B * Tn * C);
        r.conv(sd.rb[i].c1, y, h1, B, Tn, true, PRO_ELU);
        r.conv(sd.rb[i].sc, y, scb, B, Tn, true, PRO_NONE);
        r.conv(sd.rb[i].c2, h1, yo, B, Tn, true, PRO_ELU, ACT_NONE, scb);
        x = yo; T = Tn; ch = C;
    }
    r.conv(sd.last, x, audio, B, T, true, PRO_ELU);  // [B*T, 1] rows = audio [B, T]
}

// Nearest-code search on the tensor cores: (x - mu).(c - mu) with 3-pass split-fp16 operands, argmin of
// ||c - mu||^2 - 2 x.c fused into the GEMM epilogue (reference encoder/quantization/core_vq.py:175-183).
// zc planes [M, D] hold the centred frames; `keys` is an [M] u64 scratch.
void vq_tc(wt_handle* h, const __half* zc_hi, const __half* zc_lo, long long M, unsigned long long* keys,
           long long* codes, cudaStream_t s) {
    const wt_config& c = h->cfg;
    WT_CUDA(cudaMemsetAsync(keys, 0xFF, (size_t)M * sizeof(unsigned long long), s));
    TcGemm g;
    g.seg[0] = tc_taps(zc_hi, zc_lo, M, c.dimension, c.dimension, 1, 0);
    g.W_hi = h->cb_c.hi; g.W_lo = h->cb_c.lo; g.M = (int)M; g.N = c.vq_bins; g.K = c.dimension; g.passes = 3;
    g.act = TC_ACT_ARGMIN; g.bias = h->cnorm_c; g.best = keys;
    { Scope sc(h, CAT_VQ, s); launch_tap_gemm_tc(g, s); }
    { Scope sc(h, CAT_VQ, s); launch_best_to_codes(keys, codes, M, s); }
}

// ---------------------------------------------------------------------------------------
// tcgen05 encoder. Every conv is a GEMM over a TMA tensor map of the channels-last activation:
//   * stride-1 k-tap conv and strided conv (k = 2s): "window" map with OVERLAPPING rows
//     (row stride = s*C elements, row length = k*C) over the reflect-padded clip slots;
//   * ResBlock tail: conv1x1(ELU(h1)) + shortcut1x1(x) as two A segments into one accumulator.
// Producers (conv0 kernel / GEMM epilogues) write the split-fp16 planes the consumer reads, already
// ELU-activated and re-mapped into the consumer's padded layout including its reflect halo rows.
// Layouts (rows per clip): RB input  P = T+2, data at 1, halo 1/1;  strided-conv input P = (T'+1)*s,
// data at left = s - s/2, halo left / (s/2 + extra)  (reference encoder/modules/conv.py:54-61,195-211).
// ---------------------------------------------------------------------------------------
bool encoder_tc_supported(const wt_config& c, int T) {
    long long Tc = T;
    for (int i = 0; i < 4; ++i) {
        const int s_ = c.strides[i];
        const long long Tn = (Tc + s_ - 1) / s_;
        const long long hr = s_ / 2 + (Tn * s_ - Tc);
        if (Tc < 4 || Tc <= hr + 1 || Tc <= s_) return false;  // reflect halo must stay inside the clip
        Tc = Tn;
    }
    return Tc >= 4 && c.dimension == 512 && c.n_filters == 32;
}

// Planes of one encoder level's input: x (raw) and ELU(x), clip pitch T + 2 rows, data at row 1, reflect halo 1 / 1.
struct EncPlanes { __half *xr_hi = nullptr, *xr_lo = nullptr, *xe_hi = nullptr, *xe_lo = nullptr; };

// Levels [i0, i1) of the SEANet front on Bc clips. Level i0 reads the audio (i0 == 0) or `in` (length Tin per clip);
// level i1 - 1 writes the pre-LSTM rows (i1 == 4; `pre*` point at the group's time-major rows [t*Bg + b] already
// offset to the first clip) or the planes `out` of level i1 (caller-allocated, already offset to the first clip).
void encoder_levels_tc(wt_handle* h, const float* wav, int Bc, int Tin, int i0, int i1, int b0, int Bg, EncPlanes in,
                       EncPlanes out, float* pre, __half* pre_hi, __half* pre_lo, cudaStream_t s) {
    const wt_config& c = h->cfg;
    auto halves = [&](size_t n) { return reinterpret_cast<__half*>(h->alloc((n + 1) / 2)); };
    auto run = [&](TcGemm& g) {
        g.prefetch = tc_prefetch();
        Scope sc(h, CAT_ENC_CONV, s);
        launch_tap_gemm_tc(g, s);
    };
    auto want = [&](const std::string& name) { return b0 == 0 && h->taps.count(name) > 0; };
    int Tc = Tin, C = c.n_filters << i0;
    // level 0 operands come from the conv0 kernel: ELU(x0) planes and the 8-wide raw-audio windows that stand in
    // for x0 in the shortcut (composed weights); deeper levels get x and ELU(x) planes from the strided conv
    // Level 0 (conv0 + ResBlock 0) runs as ONE fp32 CUDA-core kernel that keeps ELU(x0) / ELU(h1) in shared memory
    // (K = 7 / 96 / 24 is too small for the MMA pipeline to pay); WT_ENC_L0_TC=1 selects the tcgen05 formulation.
    static const bool l0_tc = std::getenv("WT_ENC_L0_TC") != nullptr;
    const bool fused0 = !l0_tc && h->rb0_pack != nullptr;
    __half *xr_hi = in.xr_hi, *xr_lo = in.xr_lo, *xe_hi = in.xe_hi, *xe_lo = in.xe_lo;
    // Composed shortcut of level 1 (weights: prepare()): its ResBlock tail reads the ELU(y0) planes of level 0 instead of
    // raw-x planes of its own. Both row spaces then need ONE clip pitch: level 1 works on T1 + 2 rows per clip, so the
    // strided layout of level 0 gets T1 + 2 window slots per clip instead of T1 + 1.
    const bool composed1 = true;
    const __half *y0_hi = nullptr, *y0_lo = nullptr;  // ELU(y0) planes of level 0 and their size
    long long nY0 = 0;
    // Level 0's strided conv and the whole ResBlock of level 1 in ONE kernel (enc_fused.cu): x1, ELU(x1) and ELU(h1) stay
    // on the SM. The x1 tap (enc3) does not exist then, so a request for it selects the unfused launches.
    const bool l1f = fused0 && i0 == 0 && i1 >= 2 && h->l1_fused.w1 != nullptr && C == 32 && !want("enc3") &&
                     enc_l1_fused_supported(C, c.strides[0]);
    if (i0 == 0 && !fused0) {
        const size_t nX = (size_t)Bc * (Tc + 2) * C;
        const size_t nWin = (size_t)Bc * (Tc + 2) * 8;
        xr_hi = halves(nWin); xr_lo = halves(nWin); xe_hi = halves(nX); xe_lo = halves(nX);
        if (h->wav_ready) h->wav_ready(b0, Bc);
        Scope sc(h, CAT_ENC_CONV, s);
        launch_conv0_planes(wav, h->conv0_w, h->conv0_b, xr_hi, xr_lo, xe_hi, xe_lo, Bc, Tc, C, s);
    }
    int idx = 1 + 3 * i0;
    for (int i = i0; i < i1; ++i) {
        const int s_ = c.strides[i];
        const int Tn = (Tc + s_ - 1) / s_;
        const int P = Tc + 2;
        const long long rowsX = (long long)Bc * P;
        const auto& rt = h->rb_tc[i];
        const bool fused = fused0 && i == 0;
        // ---- G1: h1 = conv_k3(ELU(x)) -> ELU(h1) planes, same row space ----
        const size_t nH = (size_t)rowsX * (C / 2);
        const bool l1f_here = l1f && i == 1;
        __half *he_hi = (fused || l1f_here) ? nullptr : halves(nH), *he_lo = (fused || l1f_here) ? nullptr : halves(nH);
        if (!fused && !l1f_here) {
            TcGemm g;
            g.kw = rt.kw1;
            g.seg[0] = tc_window(xe_hi, xe_lo, rowsX * C, 3 * C, C, 0, rt.kw1);
            g.W_hi = rt.w1.hi; g.W_lo = rt.w1.lo; g.M = (int)rowsX; g.N = C / 2; g.K = rt.Kp1; g.passes = 3;
            g.bias = h->rb[i].c1.b;
            g.elu_hi = he_hi; g.elu_lo = he_lo; g.ldh2 = C / 2;
            run(g);
        }
        // ---- G2: y = conv1x1(ELU(h1)) + shortcut(x) -> ELU(y) planes in the strided conv's padded layout ----
        const int right = s_ / 2, left = s_ - right, extra = Tn * s_ - Tc;
        const int Q = Tn + ((composed1 && i == 0) ? 2 : 1);  // window slots per clip in the strided conv's row space
        const int Py = Q * s_;
        const size_t nY = (size_t)Bc * Py * C;
        __half *ye_hi = halves(nY), *ye_lo = halves(nY);
        float* y_tap = want("enc" + std::to_string(idx)) ? h->alloc(nY) : nullptr;
        if (fused) {
            // algorithmic work per sample: conv0 224 + k3 1536 + 1x1 512 + shortcut 1024 MACs; 4 B in, 128 B of planes out.
            // Host-buffer entry: one launch per copy piece, each behind ITS piece of the H2D stream only.
            const int piece = h->wav_ready ? COPY_PIECE : Bc;
            for (int p0 = 0; p0 < Bc; p0 += piece) {
                const int np = std::min(piece, Bc - p0);
                if (h->wav_ready) h->wav_ready(b0 + p0, np);
                const bool l0tc = h->l0_tc.wk3 && enc_l0_tc_supported();
                Scope sc(h, CAT_ENC_CONV, s, l0tc ? KERN_ENC_L0TC : KERN_RB0, 6592.0 * np * Tc, 132.0 * np * Tc);
                if (l0tc) {
                    RowMap mp;
                    mp.Pin = Tc + 2; mp.Tvalid = Tc; mp.Pout = Py; mp.off = left; mp.hl = left; mp.hr = right + extra;
                    launch_enc_l0_tc(h->l0_tc, wav + (size_t)p0 * Tc, ye_hi + (size_t)p0 * Py * C, ye_lo + (size_t)p0 * Py * C,
                                     y_tap ? y_tap + (size_t)p0 * Py * C : nullptr, np, Tc, mp, s);
                } else
                launch_resblock0_fused(wav + (size_t)p0 * Tc, h->rb0_pack, ye_hi + (size_t)p0 * Py * C, ye_lo + (size_t)p0 * Py * C,
                                       y_tap ? y_tap + (size_t)p0 * Py * C : nullptr, np, Tc, Py, left, right + extra, s);
            }
        } else if (l1f_here) {
            EncL1Args fa;
            fa.y0_hi = y0_hi; fa.y0_lo = y0_lo; fa.y0_elems = nY0; fa.Bc = Bc; fa.T1 = Tc;
            fa.map.Pin = P; fa.map.Tvalid = Tc; fa.map.Pout = Py; fa.map.off = left; fa.map.hl = left; fa.map.hr = right + extra;
            fa.ye_hi = ye_hi; fa.ye_lo = ye_lo; fa.y_f32 = y_tap;
            // algorithmic work per level-1 position: strided conv 64x128 + k3 32x192 + 1x1 64x32 + shortcut 64x64 MACs;
            // two new level-0 rows (32 channels, split planes) in, one row of 64 channels out
            const double k0 = (double)h->l1_fused.k0;
            Scope sc(h, CAT_ENC_CONV, s, KERN_ENC_L1F, (2.0 * 64 * k0 + 24576.0) * Bc * Tc, (2.0 * k0 + 256.0) * Bc * Tc);
            launch_enc_l1_fused(h->l1_fused, fa, s);
        } else {
            TcGemm g;
            g.nseg = 2;
            g.kw = rt.kw2;
            g.seg[0] = tc_window(he_hi, he_lo, rowsX * (C / 2), C / 2, C / 2, 0, rt.kw2);
            if (i == 0) g.seg[1] = tc_window(xr_hi, xr_lo, rowsX * 8, 8, 8, /*shift0=*/1, rt.kw2);  // audio windows
            else if (i == 1 && composed1)  // row m = b*(T1+2) + t of this GEMM <-> window slot m of the ELU(y0) planes
                g.seg[1] = tc_window(y0_hi, y0_lo, nY0, 2LL * c.strides[0] * (C / 2), (long long)c.strides[0] * (C / 2), 0, rt.kw2);
            else g.seg[1] = tc_window(xr_hi, xr_lo, rowsX * C, C, C, /*shift0=*/1, rt.kw2);
            g.W_hi = rt.w2.hi; g.W_lo = rt.w2.lo; g.M = (int)rowsX; g.N = C; g.K = rt.kw2 * (rt.kb0 + rt.kb1); g.passes = 3;
            g.bias = rt.bias2;
            g.map.Pin = P; g.map.Tvalid = Tc; g.map.Pout = Py; g.map.off = left; g.map.hl = left; g.map.hr = right + extra;
            g.elu_hi = ye_hi; g.elu_lo = ye_lo; g.ldh2 = C;
            g.out_f32 = y_tap; g.ldo = C;
            run(g);
        }
        if (y_tap) h->tap(("enc" + std::to_string(idx)).c_str(), y_tap + (size_t)left * C, Bc, Tc, C, b0, s, Py);
        // ---- G3: z = strided conv(ELU(y)) -> next level's planes (x and ELU(x)), or the pre-LSTM rows ----
        if (i == 0) { y0_hi = ye_hi; y0_lo = ye_lo; nY0 = (long long)nY; }
        const int C2 = 2 * C;
        if (l1f && i == 0) {  // the fused kernel of the next iteration reads the ELU(y0) planes directly
            xr_hi = xr_lo = xe_hi = xe_lo = nullptr;
            Tc = Tn; C = C2; idx += 3;
            continue;
        }
        TcGemm g;
        g.seg[0] = tc_window(ye_hi, ye_lo, (long long)nY, 2 * s_ * C, (long long)s_ * C);
        g.W_hi = h->down[i].w_hi; g.W_lo = h->down[i].w_lo; g.M = Bc * Q; g.N = C2; g.K = 2 * s_ * C; g.passes = 3;
        g.bias = h->down[i].b;
        g.map.Pin = Q; g.map.Tvalid = Tn;
        float* z_tap = nullptr;
        if (i < 3) {
            const size_t nX2 = (size_t)Bc * (Tn + 2) * C2;
            const bool need_x = !(composed1 && i == 0 && i1 > 1);  // level 1's shortcut reads ELU(y0) instead
            if (i == i1 - 1) { xr_hi = out.xr_hi; xr_lo = out.xr_lo; xe_hi = out.xe_hi; xe_lo = out.xe_lo; }
            else {
                xr_hi = need_x ? halves(nX2) : nullptr; xr_lo = need_x ? halves(nX2) : nullptr;
                xe_hi = halves(nX2); xe_lo = halves(nX2);
            }
            g.map.Pout = Tn + 2; g.map.off = 1; g.map.hl = 1; g.map.hr = 1;
            if (need_x) { g.out_hi = xr_hi; g.out_lo = xr_lo; g.ldh = C2; }
            g.elu_hi = xe_hi; g.elu_lo = xe_lo; g.ldh2 = C2;
            if (want("enc" + std::to_string(idx + 2))) { z_tap = h->alloc(nX2); g.out_f32 = z_tap; g.ldo = C2; }
        } else {
            g.map.Pout = Tn; g.map.off = 0; g.map.hl = 0; g.map.hr = 0;
            g.map.sb = 1; g.map.st = Bg;  // time-major rows for the LSTM
            g.out_f32 = pre; g.ldo = C2;
            g.out_hi = pre_hi; g.out_lo = pre_lo; g.ldh = C2;
        }
        run(g);
        if (i < 3) {
            if (z_tap) h->tap(("enc" + std::to_string(idx + 2)).c_str(), z_tap + (size_t)C2, Bc, Tn, C2, b0, s, Tn + 2);
        } else if (want("enc" + std::to_string(idx + 2))) {
            auto& tr = h->taps["enc" + std::to_string(idx + 2)];
            tr.B = Bc; tr.T = Tn; tr.C = C2;
            if (tr.buf && (int64_t)Bc * Tn * C2 <= tr.cap)
                for (int b = 0; b < Bc; ++b)  // gather clip b out of the time-major rows
                    WT_CUDA(cudaMemcpy2DAsync(tr.buf + (size_t)b * Tn * C2, (size_t)C2 * sizeof(float),
                                              pre + (size_t)b * C2, (size_t)Bg * C2 * sizeof(float),
                                              (size_t)C2 * sizeof(float), Tn, cudaMemcpyDeviceToDevice, s));
        }
        Tc = Tn; C = C2; idx += 3;
    }
}

// SEANet front on a chunk of Bc clips. Levels 0-1 move 0.3 GB per clip through split-fp16 planes (32 / 64 channels at
// 72000 / 36000 rows) and are HBM-bound (profiles/r01_ncu_table_encoder_chunk64_s4.txt: 4.5-4.8 TB/s). WT_ENC_SUB=n
// runs them in sub-chunks of n clips (same arena addresses every sub-chunk), each writing its slice of the level-2
// input planes of the whole chunk, with levels 2-3 once per chunk; measured slower at every n (see ENC_SUB_DEFAULT).
void encoder_front_tc(wt_handle* h, const float* wav, int Bc, int T, int b0, int Bg, float* pre, __half* pre_hi,
                      __half* pre_lo, cudaStream_t s) {
    const wt_config& c = h->cfg;
    const int sub = enc_sub();
    if (sub <= 0 || sub >= Bc || !h->taps.empty()) {
        encoder_levels_tc(h, wav, Bc, T, 0, 4, b0, Bg, EncPlanes{}, EncPlanes{}, pre, pre_hi, pre_lo, s);
        return;
    }
    auto halves = [&](size_t n) { return reinterpret_cast<__half*>(h->alloc((n + 1) / 2)); };
    const int T1 = (T + c.strides[0] - 1) / c.strides[0], T2 = (T1 + c.strides[1] - 1) / c.strides[1];
    const int C2 = 4 * c.n_filters;
    const size_t per_clip = (size_t)(T2 + 2) * C2, nX2 = (size_t)Bc * per_clip;
    EncPlanes l2;
    l2.xr_hi = halves(nX2); l2.xr_lo = halves(nX2); l2.xe_hi = halves(nX2); l2.xe_lo = halves(nX2);
    const size_t mark = h->arena_off;
    for (int bs = 0; bs < Bc; bs += sub) {
        const int Bs = std::min(sub, Bc - bs);
        h->arena_off = mark;
        EncPlanes o;
        o.xr_hi = l2.xr_hi + bs * per_clip; o.xr_lo = l2.xr_lo + bs * per_clip;
        o.xe_hi = l2.xe_hi + bs * per_clip; o.xe_lo = l2.xe_lo + bs * per_clip;
        encoder_levels_tc(h, wav + (size_t)bs * T, Bs, T, 0, 2, b0 + bs, Bg, EncPlanes{}, o, nullptr, nullptr, nullptr, s);
    }
    h->arena_off = mark;
    encoder_levels_tc(h, nullptr, Bc, T2, 2, 4, b0, Bg, l2, EncPlanes{}, pre, pre_hi, pre_lo, s);
}

// SLSTM + ELU + final k7 conv on the tensor cores: hoisted input projections, one GEMM launch per time
// step whose epilogue is the LSTM cell (gates never leave the SM), final conv over the reflect-padded rows.
// `len_tab` (device, [Bg], optional): a ragged group. Row b of every time step belongs to clip b of len_tab[b] <= L frames;
// the recurrence runs all L steps for every row (steps past a clip's end compute on filler and are never read), the
// reflect padding in front of the last conv is taken at each clip's own end, and the frames t >= len_tab[b] of z are
// filler the caller drops.
float* encoder_back_tc(wt_handle* h, const float* pre, const __half* pre_hi, const __half* pre_lo, int Bg, int L,
                       int b0, __half** zc_hi_out, __half** zc_lo_out, cudaStream_t s, const int* len_tab = nullptr) {
    const wt_config& c = h->cfg;
    const int D = c.dimension;
    const long long M = (long long)Bg * L;
    auto halves = [&](size_t n) { return reinterpret_cast<__half*>(h->alloc((n + 1) / 2)); };
    float* xin = h->alloc((size_t)M * 4 * D);
    float* cst = h->alloc((size_t)Bg * D);
    float* ybuf[4];
    __half *yh_hi[4], *yh_lo[4];
    for (int l = 0; l < c.lstm_layers; ++l) {
        ybuf[l] = h->alloc((size_t)M * D);
        yh_hi[l] = halves((size_t)M * D);
        yh_lo[l] = halves((size_t)M * D);
    }
    const __half *lin_hi = pre_hi, *lin_lo = pre_lo;
    const __half* zero = reinterpret_cast<const __half*>(h->zero_rows);
    int* counters = reinterpret_cast<int*>(h->alloc(lstm_counter_ints(Bg, L) + 32));
    // all LSTM tensors are time-major: row t*Bg + b. Step t reads rows [(t-1)*Bg, t*Bg) of the hidden-state
    // planes through ONE tensor map (row shift (t-1)*Bg), so no descriptor is built inside the time loop.
    static const bool stepwise = std::getenv("WT_LSTM_STEPWISE") != nullptr;
    static const bool no_wave = std::getenv("WT_LSTM_NO_WAVEFRONT") != nullptr;
    auto input_projection = [&](int l, const __half* ahi, const __half* alo, float* out, long long row0, long long rows,
                                int max_ctas, cudaStream_t st) {
        // xin = W_ih x + b_ih + b_hh for `rows` time-major rows (gate rows permuted like the recurrent tile)
        const auto& w = h->lstm_tc[l];
        TcGemm g;
        g.seg[0] = tc_taps(ahi + row0 * D, alo + row0 * D, rows, D, D, 1, 0);
        g.W_hi = w.w_ih.hi; g.W_lo = w.w_ih.lo; g.M = (int)rows; g.N = 4 * D; g.K = D; g.passes = 3;
        g.bias = w.bias; g.out_f32 = out + row0 * 4 * D; g.ldo = 4 * D;
        g.max_ctas = max_ctas;
        Scope sc(h, CAT_LSTM, st);
        launch_tap_gemm_tc(g, st);
    };
    const int lstm_sms = lstm_ctas(Bg, D);
    int sms = 0;
    WT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const bool wavefront = !stepwise && !no_wave && c.lstm_layers == 2 && L >= 16 && 2 * lstm_sms <= sms;
    if (wavefront) {
        // Layer wavefront: the recurrence of either layer keeps only lstm_sms (64 of 148) SMs busy and
        // is a pure latency chain, so the time axis is cut into NSEG segments and layer 2 of segment k runs on a second
        // stream WHILE layer 1 runs segment k+1. Per segment on the main stream: layer-1 steps, then the layer-2 input
        // projection of those steps on the SMs the resident layer-2 kernel leaves free (grid cap). 5 segment-times
        // instead of 8 for NSEG = 4.
        static const int NSEG = [] {
            const char* e = std::getenv("WT_LSTM_NSEG");
            const int v = e ? std::atoi(e) : 4;
            return v < 2 ? 2 : (v > 16 ? 16 : v);
        }();
        if (!h->aux_stream) WT_CUDA(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
        while (h->aux_evs.size() < (size_t)NSEG + 2) {
            cudaEvent_t e;
            WT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            h->aux_evs.push_back(e);
        }
        cudaStream_t s2 = h->aux_stream;
        float* xin2 = h->alloc((size_t)M * 4 * D);
        float* cst2 = h->alloc((size_t)Bg * D);
        int* counters2 = reinterpret_cast<int*>(h->alloc(lstm_counter_ints(Bg, L) + 32));
        // the second stream starts after everything queued so far (its buffers may still be read by earlier work)
        WT_CUDA(cudaEventRecord(h->aux_evs[NSEG], s));
        WT_CUDA(cudaStreamWaitEvent(s2, h->aux_evs[NSEG], 0));
        input_projection(0, pre_hi, pre_lo, xin, 0, M, 0, s);
        for (int k = 0; k < NSEG; ++k) {
            const int t0 = (int)((long long)L * k / NSEG), t1 = (int)((long long)L * (k + 1) / NSEG);
            {
                Scope sc(h, CAT_LSTM, s, KERN_LSTM, 16.0 * D * D * Bg * (t1 - t0), 0);
                launch_lstm_persistent(xin, ybuf[0], yh_hi[0], yh_lo[0], cst, counters, h->lstm_tc[0].w_hh.hi,
                                       h->lstm_tc[0].w_hh.lo, Bg, L, D, s, t0, t1);
            }
            // layer-2 input projection of steps [t0, t1): leaves room for the layer-2 recurrence of segment k-1
            input_projection(1, yh_hi[0], yh_lo[0], xin2, (long long)t0 * Bg, (long long)(t1 - t0) * Bg,
                             k == 0 ? 0 : sms - lstm_sms, s);
            WT_CUDA(cudaEventRecord(h->aux_evs[k], s));
            WT_CUDA(cudaStreamWaitEvent(s2, h->aux_evs[k], 0));
            {
                Scope sc(h, CAT_LSTM, s2, KERN_LSTM, 16.0 * D * D * Bg * (t1 - t0), 0);
                launch_lstm_persistent(xin2, ybuf[1], yh_hi[1], yh_lo[1], cst2, counters2, h->lstm_tc[1].w_hh.hi,
                                       h->lstm_tc[1].w_hh.lo, Bg, L, D, s2, t0, t1);
            }
        }
        WT_CUDA(cudaEventRecord(h->aux_evs[NSEG + 1], s2));
        WT_CUDA(cudaStreamWaitEvent(s, h->aux_evs[NSEG + 1], 0));
        lin_hi = yh_hi[1]; lin_lo = yh_lo[1];
    } else
    for (int l = 0; l < c.lstm_layers; ++l) {
        const auto& w = h->lstm_tc[l];
        input_projection(l, lin_hi, lin_lo, xin, 0, M, 0, s);
        if (!stepwise) {
            Scope sc(h, CAT_LSTM, s, KERN_LSTM, 16.0 * D * D * Bg * L, 0);
            launch_lstm_persistent(xin, ybuf[l], yh_hi[l], yh_lo[l], cst, counters, w.w_hh.hi, w.w_hh.lo, Bg, L, D, s);
            lin_hi = yh_hi[l]; lin_lo = yh_lo[l];
            continue;
        }
        WT_CUDA(cudaMemsetAsync(cst, 0, (size_t)Bg * D * sizeof(float), s));
        for (int t = 0; t < L; ++t) {
            TcGemm g;
            if (t == 0) {
                g.seg[0] = tc_taps(zero, zero + (size_t)ENC_GROUP * D, Bg, D, D, 1, 0);  // h_{-1} = 0
            } else {
                g.seg[0] = tc_taps(yh_hi[l], yh_lo[l], M, D, D, 1, 0);
                g.seg[0].shift0 = (t - 1) * Bg;
            }
            g.W_hi = w.w_hh.hi; g.W_lo = w.w_hh.lo; g.M = Bg; g.N = 4 * D; g.K = D; g.passes = 3;
            g.act = TC_ACT_LSTM; g.res = xin + (size_t)t * Bg * 4 * D; g.ldres = 4 * D; g.cell = cst; g.hidden = D;
            g.out_f32 = ybuf[l] + (size_t)t * Bg * D; g.ldo = D;
            g.out_hi = yh_hi[l] + (size_t)t * Bg * D; g.out_lo = yh_lo[l] + (size_t)t * Bg * D; g.ldh = D;
            Scope sc(h, CAT_LSTM, s);
            launch_tap_gemm_tc(g, s);
        }
        lin_hi = yh_hi[l]; lin_lo = yh_lo[l];
    }
    const float* ylast = c.lstm_layers ? ybuf[c.lstm_layers - 1] : pre;
    float* lo = h->alloc((size_t)M * D);
    const size_t nE = (size_t)Bg * (L + 6) * D;
    __half *e_hi = halves(nE), *e_lo = halves(nE);
    {
        Scope sc(h, CAT_LSTM, s, KERN_LSTM_SKIP, 0, (double)M * D * 12 + (double)nE * 4);
        if (c.lstm_layers) launch_lstm_skip_elu_pad(ylast, pre, lo, e_hi, e_lo, Bg, L, D, s, len_tab);
        else launch_lstm_skip_elu_pad(pre, h->zero_rows /*unused*/, lo, e_hi, e_lo, Bg, L, D, s, len_tab);
    }
    h->tap("enc13", lo, Bg, L, D, b0, s);
    float* z = h->alloc((size_t)M * D);
    __half *zc_hi = halves((size_t)M * D), *zc_lo = halves((size_t)M * D);
    *zc_hi_out = zc_hi; *zc_lo_out = zc_lo;
    {
        TcGemm g;
        g.seg[0] = tc_window(e_hi, e_lo, (long long)nE, 7 * D, D);
        g.W_hi = h->enc_last.w_hi; g.W_lo = h->enc_last.w_lo; g.M = Bg * (L + 6); g.N = D; g.K = 7 * D; g.passes = 3;
        g.bias = h->enc_last.b;
        g.map.Pin = L + 6; g.map.Tvalid = L; g.map.Pout = L; g.map.off = 0;
        g.out_f32 = z; g.ldo = D;
        g.out_hi = zc_hi; g.out_lo = zc_lo; g.ldh = D; g.plane_shift = h->cb_mean;  // centred frames for the VQ
        Scope sc(h, CAT_ENC_CONV, s);
        launch_tap_gemm_tc(g, s);
    }
    h->tap("enc15", z, Bg, L, D, b0, s);
    return z;
}

// ---------------------------------------------------------------------------------------
// decoder: VocosBackbone + ISTFTHead on a chunk (reference decoder/models.py:223-235,
// decoder/heads.py:42-67)
// ---------------------------------------------------------------------------------------
void decoder_chunk_simt(wt_handle* h, const float* features /*[Bc, Din, L]*/, int Bc, int L, int bw, float* audio,
                        int b0, cudaStream_t s) {
    const wt_config& c = h->cfg;
    Runner r{h, s};
    const int D = c.dim, Hd = c.intermediate_dim, Din = c.dimension;
    const long long M = (long long)Bc * L;
    const size_t big = std::max<size_t>(std::max<size_t>(3 * D, Hd), std::max<size_t>(h->ldz, h->Kp));
    float* xin = h->alloc((size_t)M * Din);
    float* x = h->alloc((size_t)M * D);
    float* t1 = h->alloc((size_t)M * D);
    float* t2 = h->alloc((size_t)M * D);
    float* bigA = h->alloc((size_t)M * big);
    float* bigB = h->alloc((size_t)M * big);
    const float eps = 1e-6f;

    { Scope sc(h, CAT_MEM, s); launch_transpose_bcl_to_blc(features, xin, Bc, Din, L, s); }
    r.cat = CAT_DEC_CONV;
    r.conv(h->embed, xin, x, Bc, L, false, PRO_NONE);
    h->tap("dec_embed", x, Bc, L, D, b0, s);

    auto resnet = [&](const wt_handle::Resnet& p) {
        { Scope sc(h, CAT_MEM, s); launch_groupnorm(x, p.n1w, p.n1b, out_f32(t1), Bc, L, L, D, 32, eps, 1, s); }
        r.conv(p.c1, t1, t2, Bc, L, false, PRO_NONE);
        { Scope sc(h, CAT_MEM, s); launch_groupnorm(t2, p.n2w, p.n2b, out_f32(t1), Bc, L, L, D, 32, eps, 1, s); }
        r.conv(p.c2, t1, x, Bc, L, false, PRO_NONE, ACT_NONE, x);
    };
    resnet(h->pos[0]); h->tap("dec_pos0", x, Bc, L, D, b0, s);
    resnet(h->pos[1]); h->tap("dec_pos1", x, Bc, L, D, b0, s);
    {   // AttnBlock (reference decoder/models.py:107-127)
        { Scope sc(h, CAT_MEM, s); launch_groupnorm(x, h->attn.nw, h->attn.nb, out_f32(t1), Bc, L, L, D, 32, eps, 0, s); }
        r.linear(t1, h->attn.wqkv, h->attn.bqkv, bigA, M, 3 * D, D, ACT_NONE, nullptr, nullptr);
        { Scope sc(h, CAT_ATTN, s); launch_attention(bigA, out_f32(t2), Bc, L, L, D, s); }
        r.linear(t2, h->attn.proj.w, h->attn.proj.b, x, M, D, D, ACT_NONE, nullptr, x);
        h->tap("dec_pos2", x, Bc, L, D, b0, s);
    }
    resnet(h->pos[2]); h->tap("dec_pos3", x, Bc, L, D, b0, s);
    resnet(h->pos[3]); h->tap("dec_pos4", x, Bc, L, D, b0, s);
    { Scope sc(h, CAT_MEM, s); launch_groupnorm(x, h->gn5w, h->gn5b, out_f32(t1), Bc, L, L, D, 32, eps, 0, s); }
    h->tap("dec_pos5", t1, Bc, L, D, b0, s);
    // AdaLayerNorm keyed by bandwidth_id (reference decoder/modules.py:81-86)
    { Scope sc(h, CAT_MEM, s); launch_layernorm(t1, h->norm_scale + (size_t)bw * D, h->norm_shift + (size_t)bw * D, out_f32(x), M, D, eps, s); }
    h->tap("dec_norm", x, Bc, L, D, b0, s);
    for (int i = 0; i < c.num_layers; ++i) {
        // ConvNeXtBlock (reference decoder/modules.py:43-60)
        const auto& p = h->cnx[i];
        r.cat = CAT_PWCONV;
        { Scope sc(h, CAT_MEM, s); launch_dwconv_ln(x, p.dw, p.db, p.scale + (size_t)bw * D, p.shift + (size_t)bw * D, out_f32(t1), Bc, L, L, D, eps, s); }
        r.linear(t1, p.w1, p.b1, bigA, M, Hd, D, ACT_GELU, nullptr, nullptr);
        r.linear(bigA, p.w2, p.b2, x, M, D, Hd, ACT_NONE, p.gamma, x);
        h->tap(("dec_cnx" + std::to_string(i)).c_str(), x, Bc, L, D, b0, s);
    }
    { Scope sc(h, CAT_MEM, s); launch_layernorm(x, h->fln_w, h->fln_b, out_f32(t1), M, D, eps, s); }
    h->tap("dec_final", t1, Bc, L, D, b0, s);
    const int N = c.n_fft, half = N / 2 + 1;
    r.cat = CAT_HEAD;
    r.linear(t1, h->head_w, h->head_b, bigA, M, N + 2, D, ACT_NONE, nullptr, nullptr, 0, h->ldz);
    h->tap("dec_headlin", bigA, Bc, L, N + 2, b0, s, L, h->ldz);
    { Scope sc(h, CAT_MEM, s); launch_spectral(bigA, h->ldz, out_f32(bigB), M, half, h->Kp, s); }
    r.linear(bigB, h->basis, nullptr, bigA, M, N, h->Kp, ACT_NONE, nullptr, nullptr);
    { Scope sc(h, CAT_MEM, s); launch_overlap_add(bigA, h->wsq, audio, Bc, L, L, N, c.hop_length, s); }
}

// tcgen05 plan: same dataflow in the padded row space (3 zero rows after each clip), every contraction on
// the tensor cores with split-fp16 operands, every A operand written as hi/lo planes by its producer.
// `rg` (optional): a ragged chunk. L is then the longest clip of the chunk (the common pitch is L + 3), rg.len[b] the frames
// of clip b, rg.off[b] its place in the packed `features` / `audio` arrays. Row-wise work (GEMMs, LayerNorm, spectral) runs
// over the padded row space; the kernels that look across rows of a clip read the clip's own length.
void decoder_chunk_tc(wt_handle* h, const float* features /*[Bc, Din, L]*/, int Bc, int L, int bw, float* audio,
                      int b0, cudaStream_t s, Ragged rg = Ragged{}) {
    const wt_config& c = h->cfg;
    Runner r{h, s};
    const int D = c.dim, Hd = c.intermediate_dim, Din = c.dimension;
    const int Lp = L + 3;
    const long long R = (long long)Bc * Lp;
    const int pw_passes = h->plan >= 2 ? 1 : 3;  // plan 2: single-pass fp16 for the 24 ConvNeXt GEMMs
    const size_t big = std::max<size_t>(std::max<size_t>(3 * D, Hd), std::max<size_t>(h->ldz, h->Kp));
    auto halves = [&](size_t n) { return reinterpret_cast<__half*>(h->alloc((n + 1) / 2)); };
    __half* xin_hi = halves((size_t)R * Din);
    __half* xin_lo = halves((size_t)R * Din);
    float* x = h->alloc((size_t)R * D);
    float* t2 = h->alloc((size_t)R * D);
    __half* a_hi = halves((size_t)R * D);
    __half* a_lo = halves((size_t)R * D);
    float* bigA = h->alloc((size_t)R * big);
    const size_t gw = std::max(Hd, 3 * D);  // GELU planes, also the q|k|v planes of the attention block
    __half* g_hi = halves((size_t)R * gw);
    __half* g_lo = halves((size_t)R * gw);
    __half* S_hi = halves((size_t)R * h->Kp);
    __half* S_lo = halves((size_t)R * h->Kp);
    const float eps = 1e-6f;

    auto gemm = [&](const __half* ahi, const __half* alo, int Cin, int taps, const __half* whi, const __half* wlo, int N,
                    int passes, const float* bias, int act, const float* gamma, const float* res, float* of32, int ldo,
                    __half* ohi, __half* olo, int ldh) {
        TcGemm g;
        g.seg[0] = tc_taps(ahi, alo, R, Cin, Cin, taps, (taps - 1) / 2);
        g.W_hi = whi; g.W_lo = wlo; g.M = (int)R; g.N = N; g.K = taps * Cin; g.passes = passes;
        g.bias = bias; g.act = act; g.gamma = gamma; g.res = res; g.ldres = ldo;
        g.out_f32 = of32; g.ldo = ldo; g.out_hi = ohi; g.out_lo = olo; g.ldh = ldh;
        g.prefetch = tc_prefetch();
        Scope sc(h, r.cat, s);
        launch_tap_gemm_tc(g, s);
    };

    if (h->dec_codes && !rg.len && h->cb_h.hi) {
        // host-buffer entry: the features ARE codebook rows of these codes (one codebook): gather the planes directly
        Scope sc(h, CAT_MEM, s, KERN_ROWS, 0, (double)Bc * L * (Din * 4 + 8));
        launch_codes_to_row_planes(h->cb_h.hi, h->cb_h.lo, h->dec_codes + (size_t)b0 * L, xin_hi, xin_lo, Bc, L, Lp, Din,
                                   c.vq_bins, s);
    } else {
        Scope sc(h, CAT_MEM, s, KERN_ROWS, 0, (double)Bc * L * Din * 8);
        launch_features_to_rows(features, out_split(xin_hi, xin_lo), Bc, Din, L, Lp, s, rg);
    }
    r.cat = CAT_DEC_CONV;
    gemm(xin_hi, xin_lo, Din, 7, h->embed.w_hi, h->embed.w_lo, D, 3, h->embed.b, ACT_NONE, nullptr, nullptr, x, D,
         nullptr, nullptr, 0);
    h->tap("dec_embed", x, Bc, L, D, b0, s, Lp);

    auto resnet = [&](const wt_handle::Resnet& p) {
        { Scope sc(h, CAT_MEM, s, KERN_GROUPNORM, 0, (double)Bc * L * D * 8); launch_groupnorm(x, p.n1w, p.n1b, out_split(a_hi, a_lo), Bc, L, Lp, D, 32, eps, 1, s, rg); }
        gemm(a_hi, a_lo, D, 3, p.c1.w_hi, p.c1.w_lo, D, 3, p.c1.b, ACT_NONE, nullptr, nullptr, t2, D, nullptr, nullptr, 0);
        { Scope sc(h, CAT_MEM, s, KERN_GROUPNORM, 0, (double)Bc * L * D * 8); launch_groupnorm(t2, p.n2w, p.n2b, out_split(a_hi, a_lo), Bc, L, Lp, D, 32, eps, 1, s, rg); }
        gemm(a_hi, a_lo, D, 3, p.c2.w_hi, p.c2.w_lo, D, 3, p.c2.b, ACT_NONE, nullptr, x, x, D, nullptr, nullptr, 0);
    };
    resnet(h->pos[0]); h->tap("dec_pos0", x, Bc, L, D, b0, s, Lp);
    resnet(h->pos[1]); h->tap("dec_pos1", x, Bc, L, D, b0, s, Lp);
    {   // AttnBlock on the tensor cores: QKV GEMM -> batched q.k^T -> softmax -> batched P.V -> proj (+x)
        const int Lpad = (int)align_up(L, 64);
        float* S = h->alloc((size_t)R * Lpad);
        __half *p_hi = halves((size_t)R * Lpad), *p_lo = halves((size_t)R * Lpad);
        __half *vt_hi = halves((size_t)Bc * D * Lpad), *vt_lo = halves((size_t)Bc * D * Lpad);
        __half *qkv_hi = g_hi, *qkv_lo = g_lo;  // [R, 3D] planes (the GELU buffer is free here)
        { Scope sc(h, CAT_MEM, s, KERN_GROUPNORM, 0, (double)Bc * L * D * 8); launch_groupnorm(x, h->attn.nw, h->attn.nb, out_split(a_hi, a_lo), Bc, L, Lp, D, 32, eps, 0, s, rg); }
        gemm(a_hi, a_lo, D, 1, h->attn.wqkv_h.hi, h->attn.wqkv_h.lo, 3 * D, 3, h->attn.bqkv, ACT_NONE, nullptr, nullptr,
             nullptr, 0, qkv_hi, qkv_lo, 3 * D);
        r.cat = CAT_ATTN;
        {   // scores[b, i, j] = q_i . k_j for the L frames of clip b
            TcGemm g;
            g.seg[0] = tc_taps(qkv_hi, qkv_lo, R, D, 3 * D, 1, 0);
            g.W_hi = qkv_hi + D; g.W_lo = qkv_lo + D; g.ldw = 3 * D; g.w_rows = R;
            g.M = L; g.N = Lpad; g.K = D; g.passes = 3;
            g.batch = Bc; g.a_brows = Lp; g.w_brows = Lp; g.o_brows = Lp;
            g.out_f32 = S; g.ldo = Lpad;
            Scope sc(h, CAT_ATTN, s);
            launch_tap_gemm_tc(g, s);
        }
        { Scope sc(h, CAT_ATTN, s, KERN_SOFTMAX, 0, (double)Bc * L * L * 8); launch_softmax_planes(S, Lpad, p_hi, p_lo, Lpad, Bc, L, Lp, 1.0f / sqrtf((float)D), s, rg); }
        { Scope sc(h, CAT_ATTN, s, KERN_VT, 0, (double)Bc * L * D * 8); launch_vt_planes(qkv_hi, qkv_lo, vt_hi, vt_lo, Bc, L, Lp, D, Lpad, s, rg); }
        {   // out[b, i, :] = sum_j P[b, i, j] v_j
            TcGemm g;
            g.seg[0] = tc_taps(p_hi, p_lo, R, Lpad, Lpad, 1, 0);
            g.W_hi = vt_hi; g.W_lo = vt_lo; g.ldw = Lpad; g.w_rows = (long long)Bc * D;
            g.M = L; g.N = D; g.K = Lpad; g.passes = 3;
            g.batch = Bc; g.a_brows = Lp; g.w_brows = D; g.o_brows = Lp;
            g.out_hi = a_hi; g.out_lo = a_lo; g.ldh = D;
            Scope sc(h, CAT_ATTN, s);
            launch_tap_gemm_tc(g, s);
        }
        r.cat = CAT_DEC_CONV;
        gemm(a_hi, a_lo, D, 1, h->attn.proj.w_hi, h->attn.proj.w_lo, D, 3, h->attn.proj.b, ACT_NONE, nullptr, x, x, D,
             nullptr, nullptr, 0);
        h->tap("dec_pos2", x, Bc, L, D, b0, s, Lp);
    }
    resnet(h->pos[2]); h->tap("dec_pos3", x, Bc, L, D, b0, s, Lp);
    resnet(h->pos[3]); h->tap("dec_pos4", x, Bc, L, D, b0, s, Lp);
    { Scope sc(h, CAT_MEM, s, KERN_GROUPNORM, 0, (double)Bc * L * D * 8); launch_groupnorm(x, h->gn5w, h->gn5b, out_f32(t2), Bc, L, Lp, D, 32, eps, 0, s, rg); }
    h->tap("dec_pos5", t2, Bc, L, D, b0, s, Lp);
    { Scope sc(h, CAT_MEM, s, KERN_LAYERNORM, 0, (double)Bc * L * D * 8); launch_layernorm(t2, h->norm_scale + (size_t)bw * D, h->norm_shift + (size_t)bw * D, out_f32(x), R, D, eps, s); }
    h->tap("dec_norm", x, Bc, L, D, b0, s, Lp);
    r.cat = CAT_PWCONV;
    for (int i = 0; i < c.num_layers; ++i) {
        const auto& p = h->cnx[i];
        { Scope sc(h, CAT_MEM, s, KERN_DWCONV_LN, 0, (double)Bc * L * D * (pw_passes == 3 ? 8 : 6)); launch_dwconv_ln(x, p.dw, p.db, p.scale + (size_t)bw * D, p.shift + (size_t)bw * D, out_split(a_hi, pw_passes == 3 ? a_lo : nullptr), Bc, L, Lp, D, eps, s, rg); }
        // pointwise MLP 768 -> 2304 -> 768. WT_PW_SPLIT = n runs GEMM-1 -> GEMM-2 over n row slices in turn, so that the
        // 2304-wide fp16 plane of a slice (133 MB for a whole 128-clip chunk, more than the L2) is still L2-resident when
        // GEMM-2 reads it (slices start at multiples of 256 rows = one CTA-pair tile).
        static const int pw_split = [] { const char* e = std::getenv("WT_PW_SPLIT"); const int v = e ? std::atoi(e) : 1; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
        const long long slice = pw_split > 1 ? (long long)align_up((size_t)((R + pw_split - 1) / pw_split), 256) : R;
        for (long long r0 = 0; r0 < R; r0 += slice) {
            const long long nr = std::min(slice, R - r0);
            for (int which = 0; which < 2; ++which) {
                TcGemm g;
                if (which == 0) {
                    g.seg[0] = tc_taps(a_hi + r0 * D, a_lo + r0 * D, nr, D, D, 1, 0);
                    g.W_hi = p.w1_h.hi; g.W_lo = p.w1_h.lo; g.N = Hd; g.K = D; g.bias = p.b1; g.act = ACT_GELU;
                    g.out_hi = g_hi + r0 * Hd; g.out_lo = pw_passes == 3 ? g_lo + r0 * Hd : nullptr; g.ldh = Hd;
                } else {
                    g.seg[0] = tc_taps(g_hi + r0 * Hd, g_lo + r0 * Hd, nr, Hd, Hd, 1, 0);
                    g.W_hi = p.w2_h.hi; g.W_lo = p.w2_h.lo; g.N = D; g.K = Hd; g.bias = p.b2; g.gamma = p.gamma;
                    g.res = x + r0 * D; g.ldres = D; g.out_f32 = x + r0 * D; g.ldo = D;
                }
                g.M = (int)nr; g.passes = pw_passes; g.prefetch = tc_prefetch();
                Scope sc(h, r.cat, s);
                launch_tap_gemm_tc(g, s);
            }
        }
        h->tap(("dec_cnx" + std::to_string(i)).c_str(), x, Bc, L, D, b0, s, Lp);
    }
    // final LayerNorm -> head GEMM -> exp / sincos -> inverse DFT GEMM -> overlap-add. In the host-buffer entry this tail
    // runs over SLICES of TAIL_SLICE clips in turn: the audio of a slice goes to the host (D2H on the copy stream) while
    // the next slice is computed, so only the last slice's D2H stays exposed behind the last kernel of the call.
    const int N = c.n_fft, half = N / 2 + 1;
    constexpr int TAIL_SLICE = 2 * COPY_PIECE;
    const bool sliced = h->audio_done && !rg.len && Bc > TAIL_SLICE;
    const int step_clips = sliced ? TAIL_SLICE : Bc;
    for (int c0 = 0; c0 < Bc; c0 += step_clips) {
        const int nc = std::min(step_clips, Bc - c0);
        const long long r0 = (long long)c0 * Lp, nr = (long long)nc * Lp;
        auto gemm_rows = [&](const __half* ahi, const __half* alo, int Cin, const __half* whi, const __half* wlo, int Nn,
                             const float* bias, float* of32, int ldo) {
            TcGemm g;
            g.seg[0] = tc_taps(ahi, alo, nr, Cin, Cin, 1, 0);
            g.W_hi = whi; g.W_lo = wlo; g.M = (int)nr; g.N = Nn; g.K = Cin; g.passes = 3;
            g.bias = bias; g.out_f32 = of32; g.ldo = ldo;
            g.prefetch = tc_prefetch();
            Scope sc(h, r.cat, s);
            launch_tap_gemm_tc(g, s);
        };
        { Scope sc(h, CAT_MEM, s, KERN_LAYERNORM, 0, (double)nc * L * D * 8);
          launch_layernorm(x + r0 * D, h->fln_w, h->fln_b, out_split(a_hi + r0 * D, a_lo + r0 * D, t2 + r0 * D), nr, D, eps, s); }
        if (!sliced) h->tap("dec_final", t2, Bc, L, D, b0, s, Lp);
        r.cat = CAT_HEAD;
        float* zrows = bigA + r0 * h->ldz;   // z of this slice (pitch ldz >= N: it never reaches into an earlier slice's frames)
        float* frames = bigA + r0 * N;       // windowed frames of this slice (pitch N), written when its z is dead
        gemm_rows(a_hi + r0 * D, a_lo + r0 * D, D, h->head_h.hi, h->head_h.lo, N + 2, h->head_b, zrows, h->ldz);
        if (!sliced) h->tap("dec_headlin", bigA, Bc, L, N + 2, b0, s, Lp, h->ldz);
        { Scope sc(h, CAT_MEM, s, KERN_SPECTRAL, 0, (double)nc * L * (N + 2) * 8);
          launch_spectral(zrows, h->ldz, out_split(S_hi + r0 * h->Kp, S_lo + r0 * h->Kp), nr, half, h->Kp, s); }
        gemm_rows(S_hi + r0 * h->Kp, S_lo + r0 * h->Kp, h->Kp, h->basis_h.hi, h->basis_h.lo, N, nullptr, frames, N);
        if (h->audio_done && !rg.len) {  // overlap-add per copy piece, each followed by ITS D2H
            for (int p0 = c0; p0 < c0 + nc; p0 += COPY_PIECE) {
                const int np = std::min(COPY_PIECE, c0 + nc - p0);
                { Scope sc(h, CAT_MEM, s, KERN_OLA, 0, (double)np * L * (N + c.hop_length) * 4);
                  launch_overlap_add(bigA + (size_t)p0 * Lp * N, h->wsq, audio + (size_t)p0 * L * c.hop_length, np, L, Lp, N, c.hop_length, s); }
                h->audio_done(b0 + p0, np);
            }
        } else {
            Scope sc(h, CAT_MEM, s, KERN_OLA, 0, (double)Bc * L * (N + c.hop_length) * 4);
            launch_overlap_add(bigA, h->wsq, audio, Bc, L, Lp, N, c.hop_length, s, rg);
        }
    }
}

void decoder_chunk(wt_handle* h, const float* features, int Bc, int L, int bw, float* audio, int b0, cudaStream_t s) {
    if (h->plan >= 1) decoder_chunk_tc(h, features, Bc, L, bw, audio, b0, s);
    else decoder_chunk_simt(h, features, Bc, L, bw, audio, b0, s);
}

// Deferred read-back of the sticky error flag (see wt_handle::err_flag). arm: queue the copy + event behind the kernel
// that may have raised it. poll: if the event has completed (block = wait for it), examine the flag and throw the
// IndexError the reference raises for an out-of-range code (torch embedding: "index out of range in self").
void arm_err_flag(wt_handle* h, cudaStream_t s, const char* what) {
    if (!h->err_ev) WT_CUDA(cudaEventCreateWithFlags(&h->err_ev, cudaEventDisableTiming));
    WT_CUDA(cudaMemcpyAsync(h->err_host, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    WT_CUDA(cudaEventRecord(h->err_ev, s));
    h->err_armed = true;
    h->err_what = what;
}

void poll_err_flag(wt_handle* h, bool block) {
    if (!h->err_armed) return;
    if (block) {
        WT_CUDA(cudaEventSynchronize(h->err_ev));
    } else {
        cudaError_t q = cudaEventQuery(h->err_ev);
        if (q == cudaErrorNotReady) return;
        WT_CUDA(q);
    }
    h->err_armed = false;
    if (*h->err_host) {
        *h->err_host = 0;
        WT_CUDA(cudaMemset(h->err_flag, 0, sizeof(int)));
        throw Error(WT_ERR_INDEX, h->err_what + ": index out of range in self");
    }
}

void do_encode(wt_handle* h, const float* wav, int B, int T, float* features_out, int64_t* codes_out, float* z_out,
               cudaStream_t s) {
    const wt_config& c = h->cfg;
    if (B < 0 || T <= 0) throw Error(WT_ERR_VALUE, "encode: expected wav [B, T] with T > 0");
    h->ensure_arena(workspace_bytes(h, B, T));
    const int L = frames_for(c, T), D = c.dimension;
    for (int g0 = 0; g0 < B; g0 += ENC_GROUP) {
        const int Bg = std::min(ENC_GROUP, B - g0);
        h->arena_off = 0;
        const bool tc = h->plan >= 1 && encoder_tc_supported(c, T) && c.lstm_layers >= 1;
        float* pre = h->alloc((size_t)Bg * L * D);
        __half* pre_hi = reinterpret_cast<__half*>(h->alloc((size_t)Bg * L * D / 2));
        __half* pre_lo = reinterpret_cast<__half*>(h->alloc((size_t)Bg * L * D / 2));
        const size_t mark = h->arena_off;
        for (int b0 = 0; b0 < Bg; b0 += ENC_CHUNK) {
            const int Bc = std::min(ENC_CHUNK, Bg - b0);
            h->arena_off = mark;
            const size_t ro = (size_t)b0 * L * D;
            const size_t to = (size_t)b0 * D;  // time-major: clip b0 starts at row b0 of every time step
            if (!tc && h->wav_ready) h->wav_ready(g0 + b0, Bc);  // (the tensor-core front calls it per piece itself)
            if (tc) encoder_front_tc(h, wav + (size_t)(g0 + b0) * T, Bc, T, g0 + b0, Bg, pre + to, pre_hi + to, pre_lo + to, s);
            else encoder_front(h, wav + (size_t)(g0 + b0) * T, Bc, T, g0 + b0, pre + ro, s);
        }
        h->arena_off = mark;
        __half *zc_hi = nullptr, *zc_lo = nullptr;
        float* z = tc ? encoder_back_tc(h, pre, pre_hi, pre_lo, Bg, L, g0, &zc_hi, &zc_lo, s)
                      : encoder_back(h, pre, Bg, L, g0, s);
        const long long M = (long long)Bg * L;
        if (z_out) {
            Scope sc(h, CAT_MEM, s);
            launch_transpose_blc_to_bcl(z, z_out + (size_t)g0 * D * L, Bg, L, D, s);
        }
        if (codes_out) {
            long long* codes = reinterpret_cast<long long*>(codes_out) + (size_t)g0 * L;
            if (tc) {
                unsigned long long* keys = reinterpret_cast<unsigned long long*>(h->alloc((size_t)M * 2));
                vq_tc(h, zc_hi, zc_lo, M, keys, codes, s);
            } else {
                Scope sc(h, CAT_VQ, s);
                launch_vq_simt(z, h->codebooks, h->cnorm, M, D, c.vq_bins, codes, s);
            }
            if (features_out) {
                Scope sc(h, CAT_MEM, s, KERN_GATHER, 0, (double)Bg * L * (D * 4 + 8));
                launch_codes_to_features(h->codebooks, codes, features_out + (size_t)g0 * D * L, 1, Bg, L, D, c.vq_bins,
                                         nullptr, s);
            }
        }
    }
}

// Ragged encode: B clips of DIFFERENT lengths in one call (the reference feeds one file at a time,
// infer.py:44-54, so clip b's result is by definition that of a batch-of-one call). The conv front runs per run of
// equal-length clips (reflect padding, strides and chunk shapes depend on the clip's own length); what makes a
// batch-of-one slow is the LSTM, a latency chain of L steps per layer whatever the batch, so ALL clips share one
// recurrence: their pre-LSTM rows sit side by side in the time-major group buffer (row t*B + b), shorter clips padded
// with zero rows that only their own (dropped) steps read. Last conv and VQ run once over the padded frames.
// `lengths` is host memory; wav holds the clips back to back; outputs are packed the same way:
// features [sum_b D * L_b] (clip b as [D, L_b]), codes [sum_b L_b].
void do_encode_ragged(wt_handle* h, const float* wav, const int32_t* lengths, int B, float* features_out,
                      int64_t* codes_out, cudaStream_t s) {
    const wt_config& c = h->cfg;
    if (B <= 0) return;
    if (h->plan < 1 || c.lstm_layers < 1) throw Error(WT_ERR_VALUE, "encode_ragged: needs the tcgen05 plan and an LSTM");
    const int D = c.dimension;
    std::vector<long long> woff(B + 1, 0), foff(B + 1, 0);
    std::vector<int> Ls(B);
    int Tmax = 0, Lmax = 0;
    for (int b = 0; b < B; ++b) {
        const int T = lengths[b];
        if (T <= 0) throw Error(WT_ERR_VALUE, "encode_ragged: every clip needs T > 0");
        if (!encoder_tc_supported(c, T)) throw Error(WT_ERR_VALUE, "encode_ragged: clip too short for the batched path");
        Ls[b] = frames_for(c, T);
        woff[b + 1] = woff[b] + T;
        foff[b + 1] = foff[b] + Ls[b];
        Tmax = std::max(Tmax, T);
        Lmax = std::max(Lmax, Ls[b]);
    }
    for (int g0 = 0; g0 < B; g0 += ENC_GROUP) {
        const int Bg = std::min(ENC_GROUP, B - g0);
        int Lg = 0, Tg = 0, run_max = 1;
        for (int b = g0, run = 0; b < g0 + Bg; ++b) {
            Lg = std::max(Lg, Ls[b]);
            Tg = std::max(Tg, (int)lengths[b]);
            run = (b > g0 && lengths[b] == lengths[b - 1]) ? run + 1 : 1;
            run_max = std::max(run_max, std::min(run, ENC_CHUNK));
        }
        // The fronts of single clips are chains of small launches (a 1 s clip has 1 - 30 tiles per GEMM at the deeper
        // levels): they run on NL streams, each with its own scratch region, so that they fill the GPU together.
        static const int NL = [] {
            const char* e = std::getenv("WT_RAGGED_LANES");
            const int v = e ? std::atoi(e) : 8;  // 64 clips of 64 lengths: 22.1 / 15.2 / 12.3 / 11.1 ms with 1 / 2 / 4 / 8 lanes
            return v < 1 ? 1 : (v > 8 ? 8 : v);
        }();
        const size_t front_floats = enc_front_tc_floats(c, run_max, Tg) + 1024;
        const size_t need = (enc_back_floats(c, Bg, Lg) + NL * front_floats +
                             3 * align_up((size_t)Bg * Lg * D, 64) + (size_t)Bg * Lg * 2 + Bg + 4096) * sizeof(float);
        h->ensure_arena(need);
        h->arena_off = 0;
        const size_t nPre = (size_t)Bg * Lg * D;
        float* pre = h->alloc(nPre);
        __half* pre_hi = reinterpret_cast<__half*>(h->alloc(nPre / 2));
        __half* pre_lo = reinterpret_cast<__half*>(h->alloc(nPre / 2));
        int* len_dev = reinterpret_cast<int*>(h->alloc(Bg));
        long long* codes_tmp = reinterpret_cast<long long*>(h->alloc((size_t)Bg * Lg * 2));
        WT_CUDA(cudaMemsetAsync(pre, 0, nPre * sizeof(float), s));
        WT_CUDA(cudaMemsetAsync(pre_hi, 0, nPre * sizeof(__half), s));
        WT_CUDA(cudaMemsetAsync(pre_lo, 0, nPre * sizeof(__half), s));
        WT_CUDA(cudaMemcpyAsync(len_dev, Ls.data() + g0, Bg * sizeof(int), cudaMemcpyHostToDevice, s));
        const size_t mark = h->arena_off;
        while ((int)h->lane_streams.size() < NL - 1) {
            cudaStream_t st;
            WT_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            h->lane_streams.push_back(st);
        }
        while ((int)h->lane_evs.size() < NL) {
            cudaEvent_t e;
            WT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            h->lane_evs.push_back(e);
        }
        WT_CUDA(cudaEventRecord(h->lane_evs[0], s));  // fork: the lanes start once the group buffers are cleared
        for (int l = 1; l < NL; ++l) WT_CUDA(cudaStreamWaitEvent(h->lane_streams[l - 1], h->lane_evs[0], 0));
        int run_idx = 0;
        for (int b = g0; b < g0 + Bg; ++run_idx) {  // one front pass per run of equal-length clips
            int e = b + 1;
            while (e < g0 + Bg && lengths[e] == lengths[b] && e - b < ENC_CHUNK) ++e;
            const int lane = run_idx % NL;
            cudaStream_t ls = lane == 0 ? s : h->lane_streams[lane - 1];
            h->arena_off = mark + (size_t)lane * front_floats * sizeof(float);  // a lane reuses its scratch in stream order
            const size_t to = (size_t)(b - g0) * D;  // time-major: clip b - g0 starts at row b - g0 of every time step
            encoder_front_tc(h, wav + woff[b], e - b, lengths[b], /*b0 (no taps)=*/1, Bg, pre + to, pre_hi + to, pre_lo + to, ls);
            b = e;
        }
        for (int l = 1; l < NL; ++l) {  // join
            WT_CUDA(cudaEventRecord(h->lane_evs[l], h->lane_streams[l - 1]));
            WT_CUDA(cudaStreamWaitEvent(s, h->lane_evs[l], 0));
        }
        h->arena_off = mark;
        __half *zc_hi = nullptr, *zc_lo = nullptr;
        encoder_back_tc(h, pre, pre_hi, pre_lo, Bg, Lg, /*b0=*/1, &zc_hi, &zc_lo, s, len_dev);
        const long long M = (long long)Bg * Lg;
        unsigned long long* keys = reinterpret_cast<unsigned long long*>(h->alloc((size_t)M * 2));
        vq_tc(h, zc_hi, zc_lo, M, keys, codes_tmp, s);
        for (int b = g0; b < g0 + Bg; ++b) {  // drop the filler frames: clip b keeps its first L_b codes
            long long* dst = reinterpret_cast<long long*>(codes_out) + foff[b];
            WT_CUDA(cudaMemcpyAsync(dst, codes_tmp + (size_t)(b - g0) * Lg, (size_t)Ls[b] * sizeof(long long),
                                    cudaMemcpyDeviceToDevice, s));
            if (features_out) {
                Scope sc(h, CAT_MEM, s, KERN_GATHER, 0, (double)Ls[b] * (D * 4 + 8));
                launch_codes_to_features(h->codebooks, dst, features_out + foff[b] * D, 1, 1, Ls[b], D, c.vq_bins, nullptr, s);
            }
        }
    }
}

void do_decode(wt_handle* h, const float* features, int B, int L, int bw, float* audio, cudaStream_t s) {
    const wt_config& c = h->cfg;
    if (B < 0 || L <= 0) throw Error(WT_ERR_VALUE, "decode: expected features [B, C, L] with L > 0");
    if (bw < 0 || bw >= c.adanorm_num_embeddings) throw Error(WT_ERR_INDEX, "index out of range in self (bandwidth_id)");
    const int chunk = L <= 256 ? DEC_CHUNK : std::min(DEC_CHUNK, DEC_CHUNK_HOST);  // long streams: rows per pass stay bounded
    size_t need = dec_chunk_floats(c, std::min(B, chunk), L, h->Kp) * sizeof(float) + 4096;
    h->ensure_arena(need);
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int Bc = std::min(chunk, B - b0);
        h->arena_off = 0;
        decoder_chunk(h, features + (size_t)b0 * c.dimension * L, Bc, L, bw,
                      audio + (size_t)b0 * L * c.hop_length, b0, s);
        if (h->audio_done && h->plan < 1) h->audio_done(b0, Bc);  // (the tensor-core chunk calls it per copy piece itself)
    }
}

// Ragged decode: B feature maps of DIFFERENT lengths in one call. Clip b's audio is by definition
// what a batch-of-one decode returns (GroupNorm / attention over ITS frames, zero conv padding and iSTFT "same" trimming at
// ITS ends; reference decoder/models.py:10-16,107-127, spectral_ops.py:58-73). `lengths` (host) = frames per clip;
// features: clip b as [dimension, L_b], packed back to back; audio: clip b as L_b * hop samples, packed the same way.
void do_decode_ragged(wt_handle* h, const float* features, const int32_t* lengths, int B, int bw, float* audio,
                      cudaStream_t s) {
    const wt_config& c = h->cfg;
    if (B <= 0) return;
    if (h->plan < 1) throw Error(WT_ERR_VALUE, "decode_ragged: needs the tcgen05 plan");
    if (bw < 0 || bw >= c.adanorm_num_embeddings) throw Error(WT_ERR_INDEX, "index out of range in self (bandwidth_id)");
    std::vector<long long> off(B + 1, 0);
    for (int b = 0; b < B; ++b) {
        if (lengths[b] <= 0) throw Error(WT_ERR_VALUE, "decode_ragged: every clip needs L > 0");
        off[b + 1] = off[b] + lengths[b];
    }
    // chunks of clips whose padded rows (Bc * (Lmax + 3)) stay near one decoder chunk of the uniform path
    const long long row_budget = (long long)DEC_CHUNK * 228;
    for (int b0 = 0; b0 < B;) {
        int e = b0, Lm = 0;
        while (e < B) {
            const int Ln = std::max(Lm, (int)lengths[e]);
            if (e > b0 && (long long)(e - b0 + 1) * (Ln + 3) > row_budget) break;
            Lm = Ln; ++e;
        }
        const int Bc = e - b0;
        h->ensure_arena(dec_chunk_floats(c, Bc, Lm, h->Kp) * sizeof(float) + (size_t)Bc * 16 + 8192);
        h->arena_off = 0;
        int* len_dev = reinterpret_cast<int*>(h->alloc(Bc));
        long long* off_dev = reinterpret_cast<long long*>(h->alloc((size_t)Bc * 2));
        WT_CUDA(cudaMemcpyAsync(len_dev, lengths + b0, Bc * sizeof(int), cudaMemcpyHostToDevice, s));
        WT_CUDA(cudaMemcpyAsync(off_dev, off.data() + b0, Bc * sizeof(long long), cudaMemcpyHostToDevice, s));
        Ragged rg;
        rg.len = len_dev; rg.off = off_dev;
        decoder_chunk_tc(h, features, Bc, Lm, bw, audio, /*b0 (no taps)=*/1, s, rg);
        b0 = e;
    }
}

template <typename F>
int guarded(wt_handle* h, F&& f) {
    try {
        if (!h) throw Error(WT_ERR_VALUE, "null handle");
        DeviceGuard dg(h->device);  // the caller's current device is restored on return
        poll_err_flag(h, false);    // an out-of-range code of an earlier call surfaces here once its flag has landed
        f();
        return WT_OK;
    } catch (const Error& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return WT_ERR_RUNTIME;
    }
}

}  // namespace
}  // namespace wt

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

const char* wt_last_error(void) { return g_last_error.c_str(); }
const char* wt_version(void) { return "wavtok_b200 0.1 (sm_100a)"; }

int wt_create(const wt_config* cfg, const wt_tensor* tensors, int32_t n_tensors, int32_t device, wt_handle** out) {
    if (out) *out = nullptr;
    try {
        if (!cfg || !tensors || !out) throw Error(WT_ERR_VALUE, "wt_create: null argument");
        if (cfg->n_filters != 32 || cfg->dimension != 512)
            throw Error(WT_ERR_VALUE, "unsupported encoder geometry (n_filters must be 32, dimension 512)");
        if (cfg->dim != 768) throw Error(WT_ERR_VALUE, "unsupported backbone width (dim must be 768)");
        if (cfg->dim % 32 || cfg->intermediate_dim % 16 || cfg->n_fft % 2 || cfg->n_fft % cfg->hop_length)
            throw Error(WT_ERR_VALUE, "unsupported backbone / head geometry");
        if (cfg->lstm_layers < 0 || cfg->lstm_layers > 4) throw Error(WT_ERR_VALUE, "lstm_layers must be in [0, 4]");
        if (cfg->vq_bins % 128 || cfg->num_quantizers < 1) throw Error(WT_ERR_VALUE, "vq_bins must be a multiple of 128");
        for (int i = 0; i < 4; ++i)
            if (cfg->strides[i] < 1) throw Error(WT_ERR_VALUE, "strides must be positive");
        int ndev = 0;
        WT_CUDA(cudaGetDeviceCount(&ndev));
        if (device < 0 || device >= ndev) throw Error(WT_ERR_RUNTIME, "invalid CUDA device ordinal");
        DeviceGuard dg(device);
        cudaDeviceProp prop;
        WT_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw Error(WT_ERR_RUNTIME, std::string("wavtok_b200 is built for sm_100a only; device is ") + prop.name);
        Table t;
        for (int i = 0; i < n_tensors; ++i)
            if (tensors[i].name && tensors[i].data) t.m[tensors[i].name] = {tensors[i].data, tensors[i].numel};
        std::unique_ptr<wt_handle> h(new wt_handle());
        h->cfg = *cfg;
        h->device = device;
        prepare(h.get(), t);
        WT_CUDA(cudaDeviceSynchronize());
        *out = h.release();
        return WT_OK;
    } catch (const Error& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return WT_ERR_RUNTIME;
    }
}

int wt_destroy(wt_handle* h) {
    if (!h) return WT_OK;
    try {
        DeviceGuard dg(h->device);
        cudaDeviceSynchronize();
        delete h;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return WT_ERR_RUNTIME;
    }
    return WT_OK;
}

int32_t wt_frames_for(const wt_handle* h, int32_t T) { return h ? frames_for(h->cfg, T) : -1; }

int64_t wt_workspace_bytes(const wt_handle* h, int32_t B, int32_t T) {
    return h ? (int64_t)workspace_bytes(h, B, T) : -1;
}

int wt_reserve(wt_handle* h, int32_t B, int32_t T) {
    return guarded(h, [&] { h->ensure_arena(workspace_bytes(h, B, T)); });
}

int wt_encode(wt_handle* h, const float* wav, int32_t B, int32_t T, float* features_out, int64_t* codes_out,
              void* stream) {
    return guarded(h, [&] {
        if (!wav || !codes_out) throw Error(WT_ERR_VALUE, "wt_encode: null buffer");
        do_encode(h, wav, B, T, features_out, codes_out, nullptr, (cudaStream_t)stream);
    });
}

int wt_encode_ragged(wt_handle* h, const float* wav, const int32_t* lengths, int32_t B, float* features_out,
                     int64_t* codes_out, void* stream) {
    return guarded(h, [&] {
        if (!wav || !lengths || !codes_out) throw Error(WT_ERR_VALUE, "wt_encode_ragged: null buffer");
        do_encode_ragged(h, wav, lengths, B, features_out, codes_out, (cudaStream_t)stream);
    });
}

int wt_encoder_forward(wt_handle* h, const float* wav, int32_t B, int32_t T, float* z_out, void* stream) {
    return guarded(h, [&] {
        if (!wav || !z_out) throw Error(WT_ERR_VALUE, "wt_encoder_forward: null buffer");
        do_encode(h, wav, B, T, nullptr, nullptr, z_out, (cudaStream_t)stream);
    });
}

int wt_seanet_decoder(wt_handle* h, const float* z, int32_t B, int32_t L, float* audio_out, void* stream) {
    return guarded(h, [&] {
        if (!z || !audio_out) throw Error(WT_ERR_VALUE, "wt_seanet_decoder: null buffer");
        seanet_decoder(h, z, B, L, audio_out, (cudaStream_t)stream);
    });
}

int wt_codes_to_features(wt_handle* h, const int64_t* codes, int32_t K, int32_t B, int32_t L, float* features_out,
                         void* stream) {
    return guarded(h, [&] {
        if (!codes || !features_out) throw Error(WT_ERR_VALUE, "wt_codes_to_features: null buffer");
        if (K < 1 || K > h->cfg.num_quantizers)
            throw Error(WT_ERR_INDEX, "codes_to_features: more code books than the checkpoint holds");
        cudaStream_t s = (cudaStream_t)stream;
        {
            Scope sc(h, CAT_MEM, s, KERN_GATHER, 0, (double)B * L * (h->cfg.dimension * 4 + 8));
            launch_codes_to_features(h->codebooks, reinterpret_cast<const long long*>(codes), features_out, K, B, L,
                                     h->cfg.dimension, h->cfg.vq_bins, h->err_flag, s);
        }
        arm_err_flag(h, s, "codes_to_features");
    });
}

int wt_decode(wt_handle* h, const float* features, int32_t B, int32_t L, int32_t bandwidth_id, float* audio_out,
              void* stream) {
    return guarded(h, [&] {
        if (!features || !audio_out) throw Error(WT_ERR_VALUE, "wt_decode: null buffer");
        do_decode(h, features, B, L, bandwidth_id, audio_out, (cudaStream_t)stream);
    });
}

int wt_decode_ragged(wt_handle* h, const float* features, const int32_t* lengths, int32_t B, int32_t bandwidth_id,
                     float* audio_out, void* stream) {
    return guarded(h, [&] {
        if (!features || !lengths || !audio_out) throw Error(WT_ERR_VALUE, "wt_decode_ragged: null buffer");
        do_decode_ragged(h, features, lengths, B, bandwidth_id, audio_out, (cudaStream_t)stream);
    });
}

int wt_vq(wt_handle* h, const float* x, int64_t N, int64_t* codes_out, float* quantized_out, void* stream) {
    return guarded(h, [&] {
        if (!x || !codes_out) throw Error(WT_ERR_VALUE, "wt_vq: null buffer");
        cudaStream_t s = (cudaStream_t)stream;
        const wt_config& c = h->cfg;
        if (h->plan >= 1) {
            const long long CH = 1 << 18;  // frames per pass: 0.5 GB of centred operand planes
            const int D = c.dimension;
            h->ensure_arena((size_t)std::min<long long>(N, CH) * (D * 4 + 16) + 4096);
            for (long long n0 = 0; n0 < N; n0 += CH) {
                const long long n = std::min(CH, N - n0);
                h->arena_off = 0;
                __half* hi = reinterpret_cast<__half*>(h->alloc((size_t)n * D / 2));
                __half* lo = reinterpret_cast<__half*>(h->alloc((size_t)n * D / 2));
                unsigned long long* keys = reinterpret_cast<unsigned long long*>(h->alloc((size_t)n * 2));
                { Scope sc(h, CAT_VQ, s); launch_center_split(x + (size_t)n0 * D, h->cb_mean, hi, lo, n, D, s); }
                vq_tc(h, hi, lo, n, keys, reinterpret_cast<long long*>(codes_out) + n0, s);
            }
        } else {
            Scope sc(h, CAT_VQ, s);
            launch_vq_simt(x, h->codebooks, h->cnorm, N, c.dimension, c.vq_bins, reinterpret_cast<long long*>(codes_out), s);
        }
        if (quantized_out) {
            Scope sc(h, CAT_MEM, s);
            launch_gather_rows(h->codebooks, reinterpret_cast<const long long*>(codes_out), quantized_out, N,
                               c.dimension, c.vq_bins, nullptr, s);
        }
    });
}

int wt_encode_decode_host(wt_handle* h, const float* wav_host, int32_t B, int32_t T, int32_t bandwidth_id,
                          int64_t* codes_host, float* audio_host, void* stream) {
    return guarded(h, [&] {
        if (!wav_host || !codes_host || !audio_host) throw Error(WT_ERR_VALUE, "wt_encode_decode_host: null buffer");
        struct HostMode { HostMode() { g_host_entry = true; } ~HostMode() { g_host_entry = false; } } host_mode;
        cudaStream_t s = (cudaStream_t)stream;
        const wt_config& c = h->cfg;
        const int L = frames_for(c, T);
        const size_t n_wav = (size_t)B * T, n_feat = (size_t)B * c.dimension * L, n_codes = (size_t)B * L,
                     n_audio = (size_t)B * L * c.hop_length;
        size_t need = align_up(n_wav * 4, 256) + align_up(n_feat * 4, 256) + align_up(n_codes * 8, 256) +
                      align_up(n_audio * 4, 256);
        if (need > h->stage_cap) {
            WT_CUDA(cudaDeviceSynchronize());
            if (h->stage) WT_CUDA(cudaFree(h->stage));
            h->stage = nullptr; h->stage_cap = 0;
            WT_CUDA(cudaMalloc(&h->stage, need));
            h->stage_cap = need;
        }
        char* p = h->stage;
        float* wav = (float*)p; p += align_up(n_wav * 4, 256);
        float* feat = (float*)p; p += align_up(n_feat * 4, 256);
        int64_t* codes = (int64_t*)p; p += align_up(n_codes * 8, 256);
        float* audio = (float*)p;
        // Copies run on a second stream in pieces of COPY_PIECE clips: the level-0 kernel of a piece waits only for ITS
        // clips (0.09 ms of H2D in front of the first kernel instead of a whole 64-clip chunk), and the audio of a piece
        // leaves right after its overlap-add while the next one is computed (pinned host buffers make these true async
        // DMAs); what stays exposed at the end is the D2H of the last piece.
        if (!h->copy_stream) WT_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        cudaStream_t cs = h->copy_stream;
        size_t ev_next = 0;
        auto next_event = [&]() {
            if (ev_next == h->copy_evs.size()) {
                cudaEvent_t e;
                WT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                h->copy_evs.push_back(e);
            }
            return h->copy_evs[ev_next++];
        };
        {   // the copy stream starts after everything already queued on s (the staging buffers may still be in use)
            cudaEvent_t e = next_event();
            WT_CUDA(cudaEventRecord(e, s));
            WT_CUDA(cudaStreamWaitEvent(cs, e, 0));
        }
        // H2D in pieces of COPY_PIECE clips, in clip order: piece i of the wav is on the device when h2d[i] has fired
        std::vector<cudaEvent_t> h2d;
        for (int b0 = 0; b0 < B; b0 += COPY_PIECE) {
            const int np = std::min(COPY_PIECE, B - b0);
            WT_CUDA(cudaMemcpyAsync(wav + (size_t)b0 * T, wav_host + (size_t)b0 * T, (size_t)np * T * 4,
                                    cudaMemcpyHostToDevice, cs));
            cudaEvent_t e = next_event();
            WT_CUDA(cudaEventRecord(e, cs));
            h2d.push_back(e);
        }
        struct Clear {
            wt_handle* h;
            ~Clear() { h->wav_ready = nullptr; h->audio_done = nullptr; }
        } clear{h};
        h->wav_ready = [&](int b0, int n) {  // the compute stream waits for exactly the pieces that hold these clips
            for (int i = b0 / COPY_PIECE; i <= (b0 + n - 1) / COPY_PIECE && i < (int)h2d.size(); ++i)
                WT_CUDA(cudaStreamWaitEvent(s, h2d[i], 0));
        };
        // one codebook and the tensor-core decoder: the decoder gathers its input planes from the codes, the [B, D, L] feature
        // tensor is never materialised (codes_to_features + features_to_rows: 0.23 ms per 256 clips)
        const bool from_codes = c.num_quantizers == 1 && h->plan >= 1 && h->cb_h.hi != nullptr;
        do_encode(h, wav, B, T, from_codes ? nullptr : feat, codes, nullptr, s);
        h->wav_ready = nullptr;
        {
            cudaEvent_t e = next_event();
            WT_CUDA(cudaEventRecord(e, s));
            WT_CUDA(cudaStreamWaitEvent(cs, e, 0));
            WT_CUDA(cudaMemcpyAsync(codes_host, codes, n_codes * 8, cudaMemcpyDeviceToHost, cs));
        }
        const size_t per_clip = (size_t)L * c.hop_length;
        h->audio_done = [&](int b0, int n) {  // D2H of these clips' audio behind the overlap-add that wrote them
            cudaEvent_t e = next_event();
            WT_CUDA(cudaEventRecord(e, s));
            WT_CUDA(cudaStreamWaitEvent(cs, e, 0));
            WT_CUDA(cudaMemcpyAsync(audio_host + (size_t)b0 * per_clip, audio + (size_t)b0 * per_clip,
                                    (size_t)n * per_clip * 4, cudaMemcpyDeviceToHost, cs));
        };
        struct ClearCodes { wt_handle* h; ~ClearCodes() { h->dec_codes = nullptr; } } clear_codes{h};
        if (from_codes) h->dec_codes = reinterpret_cast<const long long*>(codes);
        do_decode(h, feat, B, L, bandwidth_id, audio, s);
        h->dec_codes = nullptr;
        h->audio_done = nullptr;
        WT_CUDA(cudaStreamSynchronize(s));
        WT_CUDA(cudaStreamSynchronize(cs));
        poll_err_flag(h, true);
    });
}

int wt_tap_request(wt_handle* h, const char* stage, float* dev_buf, int64_t capacity) {
    return guarded(h, [&] {
        if (!stage) throw Error(WT_ERR_VALUE, "wt_tap_request: null stage");
        TapReq r;
        r.buf = dev_buf; r.cap = capacity;
        h->taps[stage] = r;
    });
}

int wt_tap_shape(const wt_handle* h, const char* stage, int32_t* B, int32_t* T, int32_t* C) {
    if (!h || !stage) return WT_ERR_VALUE;
    auto it = h->taps.find(stage);
    if (it == h->taps.end()) return WT_ERR_VALUE;
    if (B) *B = it->second.B;
    if (T) *T = it->second.T;
    if (C) *C = it->second.C;
    return WT_OK;
}

int wt_tap_clear(wt_handle* h) {
    return guarded(h, [&] { h->taps.clear(); });
}

int64_t wt_launch_count(const wt_handle* h) { return h ? h->launches : -1; }

int wt_timing_enable(wt_handle* h, int32_t on) {
    return guarded(h, [&] {
        WT_CUDA(cudaDeviceSynchronize());
        for (auto& e : h->evs) { h->ev_pool.push_back(e.a); h->ev_pool.push_back(e.b); }
        h->evs.clear();
        h->timing = on != 0;
    });
}

int wt_timing_read(wt_handle* h, int32_t category, double* total_ms, int64_t* n_launches) {
    return guarded(h, [&] {
        if (category < 0 || category >= CAT_COUNT) throw Error(WT_ERR_VALUE, "wt_timing_read: unknown category");
        WT_CUDA(cudaDeviceSynchronize());
        double ms = 0;
        int64_t n = 0;
        for (auto& e : h->evs) {
            if (e.cat != category) continue;
            float t = 0;
            WT_CUDA(cudaEventElapsedTime(&t, e.a, e.b));
            ms += t;
            ++n;
        }
        if (total_ms) *total_ms = ms;
        if (n_launches) *n_launches = n;
    });
}

int wt_timing_read_kernel(wt_handle* h, int32_t kern, double* total_ms, int64_t* n_launches, double* flops) {
    return guarded(h, [&] {
        WT_CUDA(cudaDeviceSynchronize());
        double ms = 0, fl = 0;
        int64_t n = 0;
        for (auto& e : h->evs) {
            if (e.kern != kern) continue;
            float t = 0;
            WT_CUDA(cudaEventElapsedTime(&t, e.a, e.b));
            ms += t;
            fl += e.flops;
            ++n;
        }
        if (total_ms) *total_ms = ms;
        if (n_launches) *n_launches = n;
        if (flops) *flops = fl;
    });
}

int wt_timing_read_kernel_bytes(wt_handle* h, int32_t kern, double* total_ms, int64_t* n_launches, double* bytes) {
    return guarded(h, [&] {
        WT_CUDA(cudaDeviceSynchronize());
        double ms = 0, by = 0;
        int64_t n = 0;
        for (auto& e : h->evs) {
            if (e.kern != kern) continue;
            float t = 0;
            WT_CUDA(cudaEventElapsedTime(&t, e.a, e.b));
            ms += t;
            by += e.bytes;
            ++n;
        }
        if (total_ms) *total_ms = ms;
        if (n_launches) *n_launches = n;
        if (bytes) *bytes = by;
    });
}

int wt_check_errors(wt_handle* h) {
    return guarded(h, [&] { poll_err_flag(h, true); });
}

int wt_test_tap_gemm(int32_t device, const float* A, int32_t rows, int32_t Cin, int32_t taps, const float* W, int32_t N,
                     const float* bias, const float* gamma, const float* res, int32_t act, int32_t passes,
                     float* out_f32, float* out_split, void* stream) {
    try {
        DeviceGuard dg(device);
        cudaStream_t s = (cudaStream_t)stream;
        const long long K = (long long)taps * Cin;
        __half *a_hi, *a_lo, *w_hi, *w_lo, *o_hi = nullptr, *o_lo = nullptr;
        WT_CUDA(cudaMalloc(&a_hi, (size_t)rows * Cin * 2 * 2));  // hi plane, then lo plane
        a_lo = a_hi + (size_t)rows * Cin;
        WT_CUDA(cudaMalloc(&w_hi, (size_t)N * K * 2 * 2));
        w_lo = w_hi + (size_t)N * K;
        const int ldh = (N + 7) / 8 * 8;
        if (out_split) {
            WT_CUDA(cudaMalloc(&o_hi, (size_t)rows * ldh * 2));
            WT_CUDA(cudaMalloc(&o_lo, (size_t)rows * ldh * 2));
        }
        launch_split_f16(A, a_hi, a_lo, rows, Cin, Cin, Cin, s);
        launch_split_f16(W, w_hi, w_lo, N, (int)K, K, K, s);
        TcGemm g;
        g.seg[0] = tc_taps(a_hi, a_lo, rows, Cin, Cin, taps, (taps - 1) / 2);
        g.W_hi = w_hi; g.W_lo = w_lo; g.M = rows; g.N = N; g.K = (int)K; g.passes = passes;
        g.bias = bias; g.gamma = gamma; g.res = res; g.ldres = N; g.act = act;
        g.out_f32 = out_f32; g.ldo = N; g.out_hi = o_hi; g.out_lo = o_lo; g.ldh = ldh;
        launch_tap_gemm_tc(g, s);
        WT_CUDA(cudaStreamSynchronize(s));
        if (out_split) {
            // reconstruct hi + lo on the host side of the test: copy planes back through a tiny kernel-free path
            std::vector<__half> hh((size_t)rows * ldh), hl((size_t)rows * ldh);
            WT_CUDA(cudaMemcpy(hh.data(), o_hi, hh.size() * 2, cudaMemcpyDeviceToHost));
            WT_CUDA(cudaMemcpy(hl.data(), o_lo, hl.size() * 2, cudaMemcpyDeviceToHost));
            std::vector<float> sum((size_t)rows * N);
            for (long long r = 0; r < rows; ++r)
                for (int n = 0; n < N; ++n)
                    sum[r * N + n] = __half2float(hh[r * ldh + n]) + __half2float(hl[r * ldh + n]);
            WT_CUDA(cudaMemcpy(out_split, sum.data(), sum.size() * 4, cudaMemcpyHostToDevice));
        }
        for (void* p : {(void*)a_hi, (void*)w_hi, (void*)o_hi, (void*)o_lo})
            if (p) cudaFree(p);
        return WT_OK;
    } catch (const Error& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return WT_ERR_RUNTIME;
    }
}

int wt_debug_timeline(long long* dev_buf) {
    set_debug_timeline(dev_buf);
    return WT_OK;
}

int wt_set_plan(wt_handle* h, int32_t plan) {
    return guarded(h, [&] {
        if (plan < 0 || plan > 2) throw Error(WT_ERR_VALUE, "unknown compute plan");
        h->plan = plan;
    });
}

}  // extern "C"
