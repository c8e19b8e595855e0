/*
 * wavtok_b200 — C ABI of the B200-native WavTokenizer inference path.
 *
 * This is the drop-in boundary for the reference's Python inference API
 * (reference decoder/pretrained.py:32-239). Every entry point takes plain pointers and
 * sizes; no torch types cross this boundary. Device pointers are raw CUDA addresses on
 * the handle's device; `stream` is a cudaStream_t passed as void* (NULL = legacy default
 * stream). All compute entry points are asynchronous on `stream`.
 *
 * Status convention: 0 = ok; non-zero = error, message via wt_last_error() (thread-local).
 *   WT_ERR_VALUE    (1)  mirrors a Python ValueError   in the reference (bad shape / size)
 *   WT_ERR_INDEX    (2)  mirrors a Python IndexError   (bandwidth_id or code out of range)
 *   WT_ERR_TYPE     (3)  mirrors a Python TypeError
 *   WT_ERR_RUNTIME  (4)  CUDA / allocation / internal failure
 *
 * Threading: one handle per GPU/rank; a handle may be used by one host thread at a time
 * (the reference is single-threaded Python, not re-entrant by design).
 *
 * Ownership: the caller owns every input/output buffer; the handle owns the prepared
 * weights and a workspace arena that grows to the largest (B, T) seen (or wt_reserve'd).
 */
#ifndef WAVTOK_B200_H
#define WAVTOK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WT_OK 0
#define WT_ERR_VALUE 1
#define WT_ERR_INDEX 2
#define WT_ERR_TYPE 3
#define WT_ERR_RUNTIME 4

typedef struct wt_handle wt_handle;

/* Model hyper-parameters: what the reference resolves from YAML `model.init_args`
 * (decoder/pretrained.py:81-92) plus the SEANet constants hard-wired in
 * decoder/feature_extractors.py:71-74. */
typedef struct wt_config {
    int32_t strides[4];            /* encoder down-sampling strides in execution order (seanet.py:100) */
    int32_t n_filters;             /* 32 */
    int32_t dimension;             /* 512: encoder output / codebook dim */
    int32_t lstm_layers;           /* 2 */
    int32_t vq_bins;               /* 4096 */
    int32_t num_quantizers;        /* codebooks present in the checkpoint (infer uses the first, vq.py:137) */
    int32_t dim;                   /* 768: backbone width */
    int32_t intermediate_dim;      /* 2304 */
    int32_t num_layers;            /* 12 ConvNeXt blocks */
    int32_t adanorm_num_embeddings;/* 4 bandwidth ids */
    int32_t n_fft;
    int32_t hop_length;
} wt_config;

/* One checkpoint tensor, fp32, HOST memory, named by its reference state_dict key
 * (decoder/pretrained.py:101-112). Unknown names are ignored (the reference filters by
 * prefix and the SEANet-decoder keys are dead weight); a missing required name is an error. */
typedef struct wt_tensor {
    const char* name;
    const float* data;
    int64_t numel;
} wt_tensor;

/* Replaces WavTokenizer.from_pretrained0802's module construction + load_state_dict
 * (decoder/pretrained.py:95-114): folds weight-norm, re-lays weights for the kernels,
 * precomputes ||C||^2 and the windowed inverse-DFT basis, uploads to `device`. */
int wt_create(const wt_config* cfg, const wt_tensor* tensors, int32_t n_tensors, int32_t device,
              wt_handle** out);
int wt_destroy(wt_handle* h);

/* Frames produced for T samples: ceil-divide through the four strides (encoder/modules/conv.py:54-61). */
int32_t wt_frames_for(const wt_handle* h, int32_t T);

/* Workspace the handle needs for a (B, T) encode+decode; wt_reserve allocates it up front
 * (otherwise the arena grows on demand, which synchronises the device). */
int64_t wt_workspace_bytes(const wt_handle* h, int32_t B, int32_t T);
int wt_reserve(wt_handle* h, int32_t B, int32_t T);

/* WavTokenizer.encode_infer (decoder/pretrained.py:186-189 -> feature_extractors.py:131-142).
 * wav [B, T] f32 (device) -> features [B, dimension, L] f32, codes [1, B, L] int64. */
int wt_encode(wt_handle* h, const float* wav, int32_t B, int32_t T, float* features_out,
              int64_t* codes_out, void* stream);

/* encode_infer of B clips of DIFFERENT lengths in one call (the caller pattern is the
 * reference's one-file-at-a-time loop, infer.py:44-54, whose per-file results this reproduces exactly: nothing is padded
 * to a common length, every clip keeps its own reflect padding and frame count). `lengths` [B] is HOST memory (samples per
 * clip); wav (device) holds the clips back to back, sum(lengths) samples. Outputs are packed in the same order:
 * features_out clip b as [dimension, L_b] (may be NULL), codes_out [sum_b L_b], L_b = wt_frames_for(lengths[b]).
 * The conv front runs per run of equal-length clips; the LSTM, the last conv and the VQ run once for all clips (the
 * recurrence is a latency chain whose cost does not depend on the batch). Clips too short for the tensor-core encoder
 * layout (a few hundred samples) -> WT_ERR_VALUE: encode those one by one with wt_encode. */
int wt_encode_ragged(wt_handle* h, const float* wav, const int32_t* lengths, int32_t B, float* features_out,
                     int64_t* codes_out, void* stream);

/* feature_extractor.encodec.encoder(wav[B,1,T]) (encoder/modules/seanet.py:143-144):
 * the pre-quantisation latent z [B, dimension, L]. */
int wt_encoder_forward(wt_handle* h, const float* wav, int32_t B, int32_t T, float* z_out, void* stream);

/* SEANet decoder, `model.feature_extractor.encodec.decoder(z)` of the reference (encoder/modules/seanet.py:189-238,
 * built by decoder/feature_extractors.py:76-79): z [B, dimension, L] -> audio [B, L * hop].
 * Its weights (`feature_extractor.encodec.decoder.*`) are optional in wt_create: WT_ERR_RUNTIME when they were not given.
 * Next to the hot path, not on it: fp32 CUDA-core kernels. */
int wt_seanet_decoder(wt_handle* h, const float* z, int32_t B, int32_t L, float* audio_out, void* stream);

/* WavTokenizer.codes_to_features (decoder/pretrained.py:209-239).
 * codes [K, B, L] int64 (device) -> features [B, dimension, L] (sum over the K codebooks). */
int wt_codes_to_features(wt_handle* h, const int64_t* codes, int32_t K, int32_t B, int32_t L,
                         float* features_out, void* stream);

/* An out-of-range code (>= vq_bins; the reference raises IndexError from the embedding lookup, pretrained.py:236) is
 * detected by the gather kernel and recorded in a sticky device flag. wt_codes_to_features does NOT synchronise to
 * read it (the call stays asynchronous, like the reference on a CUDA device where the index assert is asynchronous
 * too): the error is reported, as WT_ERR_INDEX, by the next call on the handle that finds the flag's read-back
 * complete, by wt_encode_decode_host (which synchronises anyway), or by wt_check_errors, which waits for it. */
int wt_check_errors(wt_handle* h);

/* WavTokenizer.decode (decoder/pretrained.py:192-207): features [B, dimension, L] ->
 * audio [B, L*hop_length]. One bandwidth_id for the whole batch (decoder/modules.py:81-86). */
int wt_decode(wt_handle* h, const float* features, int32_t B, int32_t L, int32_t bandwidth_id,
              float* audio_out, void* stream);

/* decode of B feature maps of DIFFERENT lengths in one call (same caller pattern and exactness contract as
 * wt_encode_ragged: clip b's audio equals a batch-of-one wt_decode). lengths [B]: HOST memory, frames per clip;
 * features (device): clip b as [dimension, L_b], packed back to back; audio_out: clip b as L_b * hop_length samples,
 * packed the same way. All clips share one padded row space (pitch = longest clip of the chunk + 3): the GEMMs run over
 * it as usual; GroupNorm, attention, the depthwise conv and the overlap-add read each clip's own length. */
int wt_decode_ragged(wt_handle* h, const float* features, const int32_t* lengths, int32_t B, int32_t bandwidth_id,
                     float* audio_out, void* stream);

/* EuclideanCodebook.quantize + dequantize (encoder/quantization/core_vq.py:175-190) on
 * row-major frames x [N, dimension] -> codes [N] int64, quantized [N, dimension] (may be NULL). */
int wt_vq(wt_handle* h, const float* x, int64_t N, int64_t* codes_out, float* quantized_out, void* stream);

/* Whole hot path with HOST buffers (pinned memory recommended): H2D wav, encode, decode of the
 * quantised features, D2H codes + audio, then synchronises `stream`.
 * wav_host [B, T]; codes_host [B, L] int64; audio_host [B, L*hop]. */
int wt_encode_decode_host(wt_handle* h, const float* wav_host, int32_t B, int32_t T, int32_t bandwidth_id,
                          int64_t* codes_host, float* audio_host, void* stream);

/* Debug taps for per-stage parity tests: after the next encode/decode, the named stage's
 * output is copied (channels-last [B, T, C]) into dev_buf (capacity in floats). Stage names:
 * "enc0".."enc15", "dec_embed", "dec_pos0".."dec_pos5", "dec_norm", "dec_cnx0".."dec_cnxN",
 * "dec_final", "dec_headlin". wt_tap_shape reports the shape seen at the last run. */
int wt_tap_request(wt_handle* h, const char* stage, float* dev_buf, int64_t capacity);
int wt_tap_shape(const wt_handle* h, const char* stage, int32_t* B, int32_t* T, int32_t* C);
int wt_tap_clear(wt_handle* h);

/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t wt_launch_count(const wt_handle* h);

/* Per-category kernel timing for roofline reports: when enabled, every kernel launch is bracketed
 * by CUDA events on its own stream. Categories: 0 encoder convs, 1 LSTM, 2 VQ, 3 decoder convs
 * (embed / ResnetBlock / attention 1x1), 4 ConvNeXt pointwise GEMMs, 5 head + inverse-DFT GEMMs,
 * 6 attention core, 7 memory-bound kernels (norms, depthwise conv, transposes, gathers, overlap-add).
 * wt_timing_enable(h, on) clears the record; wt_timing_read synchronises and sums one category. */
int wt_timing_enable(wt_handle* h, int32_t on);
int wt_timing_read(wt_handle* h, int32_t category, double* total_ms, int64_t* n_launches);
/* Same record, summed per tcgen05 GEMM kernel variant: kern = BN * 10 + passes (2563 = tap_gemm_tc_kernel<256, 3>);
 * flops = algorithmic 2*M*N*K of those launches (split-precision passes not counted). */
int wt_timing_read_kernel(wt_handle* h, int32_t kern, double* total_ms, int64_t* n_launches, double* flops);
/* The other kernels of the step carry ids below 100: 1 lstm_persistent_kernel (flops = 16*D*D per clip and step),
 * 2 resblock0_fused_kernel, 3 groupnorm_kernel, 4 dwconv_ln_kernel, 5 layernorm_kernel, 6 spectral_kernel,
 * 7 overlap_add_kernel, 8 softmax_planes_kernel, 9 vt_planes_kernel, 10 features_to_rows_kernel,
 * 11 codes_to_features_kernel, 12 lstm_skip_elu_pad_kernel. For these `bytes` is the ALGORITHMIC (compulsory) HBM
 * traffic of the launches: the numerator of their bandwidth roofline. */
int wt_timing_read_kernel_bytes(wt_handle* h, int32_t kern, double* total_ms, int64_t* n_launches, double* bytes);

/* convert_audio front-end: replaces encoder/utils.py:79-92, i.e. the channel mix
 * (target_channels == 1: mean over channels; == 2 or mono input: expand) followed by
 * torchaudio.transforms.Resample(sr, target_sr) (polyphase windowed sinc: sinc_interp_hann, lowpass_filter_width 6,
 * rolloff 0.99). Handle-free. wav [B, channels, T] and out [B, target_channels, wt_convert_audio_length(T, sr,
 * target_sr)] are fp32 DEVICE memory on `device`. Errors mirror the reference: channels not in {1, 2} -> WT_ERR_VALUE
 * ("Audio must be mono or stereo."), impossible channel conversion -> WT_ERR_RUNTIME. */
int64_t wt_convert_audio_length(int64_t T, int64_t sr, int64_t target_sr);
int wt_convert_audio(int32_t device, const float* wav, int64_t B, int32_t channels, int64_t T, int64_t sr,
                     int64_t target_sr, int32_t target_channels, float* out, void* stream);

/* save_audio back-end: replaces the limiter of encoder/utils.py:95-103 and the
 * float -> PCM_S 16 conversion inside torchaudio.save(..., encoding='PCM_S', bits_per_sample=16) (utils.py:103,
 * infer.py:70). Handle-free. Each row of wav [B, T] (fp32 DEVICE memory) is one file. mode 0: no limiter
 * (infer.py:70); 1: wav.clamp(-0.99, 0.99) (rescale=False); 2: wav * min(0.99 / max|wav|, 1) (rescale=True).
 * peak [B] fp32 device (required for mode 2, where it receives max|wav| per file; unused otherwise); limited_out [B, T]
 * (optional) the limited float samples handed to the writer; pcm_out [B, T] int16. The 16-bit conversion restates
 * torchaudio 2.0.1 / libsox (round half up at 2^-16 of the 32-bit sample, clip at +32767); see csrc/audio_ops.cu. */
int wt_save_audio_pcm16(int32_t device, const float* wav, int64_t B, int64_t T, int32_t mode, float* peak,
                        float* limited_out, int16_t* pcm_out, void* stream);

/* Kernel-level test hook for the tcgen05 tap-GEMM (handle-free; allocates and frees its own scratch,
 * synchronises). out[m, n] = epi(sum_{j<taps} sum_c A[m + j - (taps-1)/2, c] * W[n, j*Cin + c]) over the
 * rows of A [rows, Cin] (rows outside are zero); W [N, taps*Cin]; all pointers fp32 DEVICE memory; bias /
 * gamma [N] and res [rows, N] optional; act 0 none, 1 GELU; passes 1 or 3 (split-fp16 operand passes).
 * out_f32 [rows, N]; out_split (optional) receives hi + lo of the split-fp16 output planes. */
int wt_test_tap_gemm(int32_t device, const float* A, int32_t rows, int32_t Cin, int32_t taps, const float* W, int32_t N,
                     const float* bias, const float* gamma, const float* res, int32_t act, int32_t passes,
                     float* out_f32, float* out_split, void* stream);

/* Performance debugging: when dev_buf != NULL every subsequent tcgen05 GEMM launch writes per-CTA clock64
 * stamps (64 int64 slots per CTA: 0 prologue done, 1..16 producer issued k-block, 17..32 k-block landed,
 * 40 first accumulator ready, 41 first epilogue done, 42 all roles finished). NULL switches it off. */
int wt_debug_timeline(long long* dev_buf);

/* Compute plan: 0 = fp32 CUDA-core contractions everywhere (bit-conservative);
 * 1 = tcgen05 tensor-core contractions with split-fp16 operands, 3 passes (hi*hi + hi*lo + lo*hi);
 * 2 = as 1, but the 24 ConvNeXt pointwise GEMMs run single-pass fp16. Default: 2. */
int wt_set_plan(wt_handle* h, int32_t plan);

const char* wt_last_error(void);
const char* wt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WAVTOK_B200_H */
