// convert_audio front-end (reference encoder/utils.py:79-92): channel mix
// (mean to mono / expand) fused with torchaudio.transforms.Resample(sr -> target_sr), i.e. the polyphase
// windowed-sinc FIR of torchaudio.functional._get_sinc_resample_kernel / _apply_sinc_resample_kernel
// (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99; torchaudio 2.x, a dependency of the reference that is
// not vendored in it: requirements.txt pins torchaudio==2.0.1, same algorithm).
//
//   out[j*new + p] = sum_k xpad[j*orig + k] * kernel[p][k],  xpad = zero-pad(x, width, width + orig),
//   K = 2*width + orig taps, output cut to ceil(new * T / orig) samples.
//
// One block = 256 consecutive output samples of one (clip, output channel): the input span they touch
// ((256/new + 1)*orig + K samples, channel-mixed on the fly) is staged in shared memory once; the tap table is
// stored transposed [k][phase] so that consecutive threads (consecutive phases) read consecutive addresses.
// Bound: HBM for the common ratios (K = 16..28 taps), fp32 FMA issue for 44.1 kHz -> 24 kHz (171 taps).
#include <cmath>
#include <cstdint>
#include <map>
#include <algorithm>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/wavtok_b200.h"
#include "common.cuh"

namespace wt {

long long gcd_ll(long long a, long long b) { return b ? gcd_ll(b, a % b) : a; }

namespace {

constexpr int RS_THREADS = 256;

__global__ void __launch_bounds__(RS_THREADS)
resample_mix_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ tabT, int C_in,
                    int C_out, int mix_mean, long long T, long long Tout, int orig, int new_, int width, int K) {
    extern __shared__ float xs[];
    const int bc = blockIdx.y;
    const int b = bc / C_out, co = bc - b * C_out;
    const long long i0 = (long long)blockIdx.x * RS_THREADS;
    long long i_last = i0 + RS_THREADS - 1;
    if (i_last > Tout - 1) i_last = Tout - 1;
    const long long j_lo = i0 / new_, j_hi = i_last / new_;
    const int span = (int)((j_hi - j_lo) * orig) + K;
    const long long t_base = j_lo * orig - width;
    const float* xb = in + (long long)b * C_in * T;
    const float inv_c = 1.f / (float)C_in;
    for (int s = threadIdx.x; s < span; s += RS_THREADS) {
        const long long t = t_base + s;
        float v = 0.f;
        if (t >= 0 && t < T) {
            if (mix_mean) {  // wav.mean(-2): sum over channels, then divide (reference encoder/utils.py:84)
                float acc = 0.f;
                for (int c = 0; c < C_in; ++c) acc += xb[(long long)c * T + t];
                v = C_in == 1 ? acc : acc * inv_c;
            } else {
                v = xb[(long long)(C_in == 1 ? 0 : co) * T + t];  // expand
            }
        }
        xs[s] = v;
    }
    __syncthreads();
    const long long i = i0 + threadIdx.x;
    if (i >= Tout) return;
    const long long j = i / new_;
    const int p = (int)(i - j * new_);
    const float* x0 = xs + (int)((j - j_lo) * orig);
    const float* w = tabT + p;
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) acc = fmaf(x0[k], w[(long long)k * new_], acc);
    out[((long long)b * C_out + co) * Tout + i] = acc;
}

// Register-tiled variant for phase counts that are multiples of 4 (44.1 kHz -> 24 kHz: 80 phases x 171 taps): a block owns
// RT_FRAMES output frames x all phases, a thread 4 frames x 4 phases, so every tap costs one 128-bit table load and four
// shared-memory loads for 16 FMAs (the one-output-per-thread kernel above pays two loads per FMA).
constexpr int RT_FRAMES = 32;

__global__ void __launch_bounds__(1024)
resample_mix_tiled_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ tabT, int C_in,
                          int C_out, int mix_mean, long long T, long long Tout, int orig, int new_, int width, int K) {
    extern __shared__ float xs[];
    const int bc = blockIdx.y;
    const int b = bc / C_out, co = bc - b * C_out;
    const long long j_lo = (long long)blockIdx.x * RT_FRAMES;
    const int span = (RT_FRAMES - 1) * orig + K;
    const long long t_base = j_lo * orig - width;
    const float* xb = in + (long long)b * C_in * T;
    const float inv_c = 1.f / (float)C_in;
    for (int s = threadIdx.x; s < span; s += blockDim.x) {
        const long long t = t_base + s;
        float v = 0.f;
        if (t >= 0 && t < T) {
            if (mix_mean) {
                float acc = 0.f;
                for (int c = 0; c < C_in; ++c) acc += xb[(long long)c * T + t];
                v = C_in == 1 ? acc : acc * inv_c;
            } else {
                v = xb[(long long)(C_in == 1 ? 0 : co) * T + t];
            }
        }
        xs[s] = v;
    }
    __syncthreads();
    const int npg = new_ >> 2;
    const int jg = threadIdx.x / npg, pg = threadIdx.x - jg * npg;
    const int p0 = pg * 4;
    const float* x0 = xs + (4 * jg) * orig;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float4 w = *reinterpret_cast<const float4*>(tabT + (long long)k * new_ + p0);
        const float xv[4] = {x0[k], x0[orig + k], x0[2 * orig + k], x0[3 * orig + k]};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            acc[r][0] = fmaf(xv[r], w.x, acc[r][0]);
            acc[r][1] = fmaf(xv[r], w.y, acc[r][1]);
            acc[r][2] = fmaf(xv[r], w.z, acc[r][2]);
            acc[r][3] = fmaf(xv[r], w.w, acc[r][3]);
        }
    }
    float* ob = out + ((long long)b * C_out + co) * Tout;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const long long i = (j_lo + 4 * jg + r) * new_ + p0;
        if (i + 3 < Tout && ((reinterpret_cast<uintptr_t>(ob + i) & 15) == 0)) {
            *reinterpret_cast<float4*>(ob + i) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (i + c < Tout) ob[i + c] = acc[r][c];
        }
    }
}

struct Table {
    int orig = 0, new_ = 0, width = 0, K = 0;
    float* dev = nullptr;  // [K][new_]
};

// torchaudio.functional._get_sinc_resample_kernel, evaluated like the reference does: sample positions in fp64,
// the phase offset -p/new in fp32 (an int64 tensor divided by a Python int), window and sinc in fp64, cast to fp32.
const Table& get_table(int device, long long sr, long long target) {
    static std::map<std::tuple<int, long long, long long>, Table> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(device, sr, target);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const long long g = gcd_ll(sr, target);
    Table t;
    t.orig = (int)(sr / g);
    t.new_ = (int)(target / g);
    if (sr == target) {  // torchaudio.transforms.Resample.forward returns its input when the rates agree: mix only
        t.width = 0; t.K = 1;
        const float one = 1.f;
        WT_CUDA(cudaMalloc(&t.dev, sizeof(float)));
        WT_CUDA(cudaMemcpy(t.dev, &one, sizeof(float), cudaMemcpyHostToDevice));
        return cache.emplace(key, t).first->second;
    }
    const double lowpass = 6.0, rolloff = 0.99;
    const double base = (double)std::min(t.orig, t.new_) * rolloff;
    t.width = (int)std::ceil(lowpass * t.orig / base);
    t.K = 2 * t.width + t.orig;
    std::vector<float> tab((size_t)t.K * t.new_);
    const double pi = 3.14159265358979323846;
    const double scale = base / t.orig;
    for (int p = 0; p < t.new_; ++p) {
        const double ph = (double)((float)(-p) / (float)t.new_);
        for (int k = 0; k < t.K; ++k) {
            double x = (ph + (double)(k - t.width) / (double)t.orig) * base;
            if (x < -lowpass) x = -lowpass;
            if (x > lowpass) x = lowpass;
            const double c = std::cos(x * pi / lowpass / 2.0);
            const double window = c * c;
            x *= pi;
            const double s = x == 0.0 ? 1.0 : std::sin(x) / x;
            tab[(size_t)k * t.new_ + p] = (float)(s * window * scale);
        }
    }
    WT_CUDA(cudaMalloc(&t.dev, tab.size() * sizeof(float)));
    WT_CUDA(cudaMemcpy(t.dev, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    return cache.emplace(key, t).first->second;
}


// ---------------------------------------------------------------------------------------
// save_audio back-end (second half): the limiter of reference encoder/utils.py:95-103
// fused with the float -> 16-bit signed PCM conversion behind torchaudio.save(..., encoding='PCM_S',
// bits_per_sample=16) (also called bare in infer.py:70). Each row of wav [B, T] is one file.
//   limiter (the reference's own arithmetic, pinned by goldens):
//     mode 1  wav.clamp(-0.99, 0.99)
//     mode 2  wav * min(0.99 / max|wav|, 1)   with torch's  scalar / tensor  = reciprocal(tensor) * scalar, in fp32
//     mode 0  none (infer.py:70)
//   quantiser (torchaudio==2.0.1 sox_io backend, NOT vendored and not runnable here: restated, parity unpinned):
//     s32 = trunc(clamp(double(x) * 2^31, INT32_MIN, INT32_MAX));  s16 = s32 > INT32_MAX - 2^15 ? 32767
//     : (uint32(s32 + 2^15) >> 16)      (libsox SOX_SAMPLE_TO_SIGNED_16BIT: round half up, clip at the top)
// Memory-bound: 4 B read + 2 B written per sample (mode 2 reads twice; the second read is L2-resident for
// batches below the 126 MB L2).
constexpr int PCM_THREADS = 256, PCM_PER_THREAD = 8;

__global__ void __launch_bounds__(PCM_THREADS)
absmax_rows_kernel(const float* __restrict__ wav, long long T, int* __restrict__ peak_bits) {
    __shared__ float red[PCM_THREADS / 32];
    const float* row = wav + (long long)blockIdx.y * T;
    const long long n0 = (long long)blockIdx.x * (PCM_THREADS * PCM_PER_THREAD) + threadIdx.x;
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < PCM_PER_THREAD; ++k) {
        const long long n = n0 + (long long)k * PCM_THREADS;
        if (n < T) m = fmaxf(m, fabsf(row[n]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < PCM_THREADS / 32; ++i) m = fmaxf(m, red[i]);
        atomicMax(peak_bits + blockIdx.y, __float_as_int(m));  // non-negative floats order like their bit patterns
    }
}

__device__ __forceinline__ int pcm16_sox(float x) {
    // x * 2^31 is exact in fp32 (power-of-two scaling) and cvt.rzi.s32.f32 truncates toward zero and saturates to
    // [INT32_MIN, INT32_MAX]: the same s32 as clamp(double(x) * 2^31).to(int32), without the fp64 pipe
    const int s32 = __float2int_rz(x * 2147483648.f);
    if (s32 > 2147483647 - 32768) return 32767;
    return (int)(short)(unsigned short)(((unsigned int)s32 + 32768u) >> 16);
}

__device__ __forceinline__ float limit_sample(float v, int mode, float scale) {
    if (mode == 1) return fminf(fmaxf(v, -0.99f), 0.99f);
    if (mode == 2) return __fmul_rn(v, scale);
    return v;
}

// VEC: every row starts 32-byte aligned (T % 8 == 0, aligned bases): a thread owns 8 consecutive samples, two
// 16-byte loads and ONE 16-byte store of eight int16. Otherwise scalar, still coalesced.
template <bool VEC>
__global__ void __launch_bounds__(PCM_THREADS)
limit_pcm16_kernel(const float* __restrict__ wav, long long T, int mode, const float* __restrict__ peak,
                   float* __restrict__ limited, short* __restrict__ pcm) {
    const long long base = (long long)blockIdx.y * T;
    float scale = 1.f;
    if (mode == 2) {
        const float sc = __fmul_rn(__frcp_rn(peak[blockIdx.y]), 0.99f);  // 0.99 / mx as torch evaluates it
        scale = 1.f < sc ? 1.f : sc;                                      // python min(tensor, 1)
    }
    if (VEC) {
        const long long n = ((long long)blockIdx.x * PCM_THREADS + threadIdx.x) * PCM_PER_THREAD;
        if (n >= T) return;
        const float4 a = __ldcs(reinterpret_cast<const float4*>(wav + base + n));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(wav + base + n + 4));
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        unsigned int q[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = limit_sample(v[i], mode, scale);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            q[i] = ((unsigned int)pcm16_sox(v[2 * i]) & 0xFFFFu) | ((unsigned int)pcm16_sox(v[2 * i + 1]) << 16);
        if (limited) {
            *reinterpret_cast<float4*>(limited + base + n) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(limited + base + n + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __stcs(reinterpret_cast<uint4*>(pcm + base + n), make_uint4(q[0], q[1], q[2], q[3]));
    } else {
        const long long n0 = (long long)blockIdx.x * (PCM_THREADS * PCM_PER_THREAD) + threadIdx.x;
#pragma unroll
        for (int k = 0; k < PCM_PER_THREAD; ++k) {
            const long long n = n0 + (long long)k * PCM_THREADS;
            if (n < T) {
                const float v = limit_sample(wav[base + n], mode, scale);
                if (limited) limited[base + n] = v;
                pcm[base + n] = (short)pcm16_sox(v);
            }
        }
    }
}

}  // namespace

}  // namespace wt

extern "C" {

int64_t wt_convert_audio_length(int64_t T, int64_t sr, int64_t target_sr) {
    if (T < 0 || sr <= 0 || target_sr <= 0) return -1;
    if (sr == target_sr) return T;
    const long long g = wt::gcd_ll(sr, target_sr);
    const long long orig = sr / g, nw = target_sr / g;
    return (nw * T + orig - 1) / orig;  // ceil(new * T / orig)
}

int wt_convert_audio(int32_t device, const float* wav, int64_t B, int32_t channels, int64_t T, int64_t sr,
                     int64_t target_sr, int32_t target_channels, float* out, void* stream) {
    using namespace wt;
    try {
        if (!wav || !out) throw Error(WT_ERR_VALUE, "wt_convert_audio: null buffer");
        if (B < 0 || T < 0 || sr <= 0 || target_sr <= 0) throw Error(WT_ERR_VALUE, "wt_convert_audio: bad sizes");
        if (channels != 1 && channels != 2) throw Error(WT_ERR_VALUE, "Audio must be mono or stereo.");
        if (target_channels != 1 && target_channels != 2 && channels != 1)
            throw Error(WT_ERR_RUNTIME, "Impossible to convert from " + std::to_string(channels) + " to " +
                                            std::to_string(target_channels));
        if (target_channels < 1) throw Error(WT_ERR_VALUE, "wt_convert_audio: target_channels must be >= 1");
        WT_CUDA(cudaSetDevice(device));
        cudaStream_t s = (cudaStream_t)stream;
        const long long Tout = wt_convert_audio_length(T, sr, target_sr);
        if (B == 0 || Tout == 0) return WT_OK;
        const int mix = target_channels == 1 ? 1 : 0;
        if (sr == target_sr && !mix && channels == target_channels) {  // Resample(orig == new) returns its input
            WT_CUDA(cudaMemcpyAsync(out, wav, (size_t)B * channels * T * sizeof(float), cudaMemcpyDeviceToDevice, s));
            return WT_OK;
        }
        const Table& t = get_table(device, sr, target_sr);
        const long long by = B * target_channels;
        if (by > 65535) throw Error(WT_ERR_VALUE, "wt_convert_audio: more than 65535 (clip, channel) rows per call");
        static PerDevice<size_t> attr_a_dev, attr_b_dev;  // opted-in dynamic shared memory per device (0: the 48 KB default)
        size_t& attr_a = attr_a_dev.get();
        size_t& attr_b = attr_b_dev.get();
        if (!attr_a) attr_a = 48 * 1024;
        if (!attr_b) attr_b = 48 * 1024;
        const size_t smem_t = ((size_t)(RT_FRAMES - 1) * t.orig + t.K) * sizeof(float);
        const int threads_t = (RT_FRAMES / 4) * (t.new_ / 4);
        if (t.new_ % 4 == 0 && threads_t <= 1024 && threads_t >= 32 && smem_t <= 160 * 1024) {
            if (smem_t > attr_b) {
                WT_CUDA(cudaFuncSetAttribute(resample_mix_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                attr_b = 160 * 1024;
            }
            const long long frames = (Tout + t.new_ - 1) / t.new_;
            dim3 grid((unsigned)((frames + RT_FRAMES - 1) / RT_FRAMES), (unsigned)by);
            resample_mix_tiled_kernel<<<grid, threads_t, smem_t, s>>>(wav, out, t.dev, channels, target_channels, mix, T, Tout,
                                                                      t.orig, t.new_, t.width, t.K);
        } else {
            const int span_max = (RS_THREADS / t.new_ + 2) * t.orig + t.K;
            const size_t smem = (size_t)span_max * sizeof(float);
            if (smem > 200 * 1024) throw Error(WT_ERR_VALUE, "wt_convert_audio: resampling ratio too extreme for the on-chip window");
            if (smem > attr_a) {
                WT_CUDA(cudaFuncSetAttribute(resample_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_a = 200 * 1024;
            }
            dim3 grid((unsigned)((Tout + RS_THREADS - 1) / RS_THREADS), (unsigned)by);
            resample_mix_kernel<<<grid, RS_THREADS, smem, s>>>(wav, out, t.dev, channels, target_channels, mix, T, Tout, t.orig,
                                                               t.new_, t.width, t.K);
        }
        WT_CUDA(cudaGetLastError());
        return WT_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return WT_ERR_RUNTIME;
    }
}

int wt_save_audio_pcm16(int32_t device, const float* wav, int64_t B, int64_t T, int32_t mode, float* peak,
                        float* limited_out, int16_t* pcm_out, void* stream) {
    using namespace wt;
    try {
        if (!wav || !pcm_out) throw Error(WT_ERR_VALUE, "wt_save_audio_pcm16: null buffer");
        if (B < 0 || T < 0) throw Error(WT_ERR_VALUE, "wt_save_audio_pcm16: bad sizes");
        if (mode < 0 || mode > 2) throw Error(WT_ERR_VALUE, "wt_save_audio_pcm16: mode must be 0 (none), 1 (clamp) or 2 (rescale)");
        if (mode == 2 && !peak) throw Error(WT_ERR_VALUE, "wt_save_audio_pcm16: rescale needs the per-file peak buffer");
        if (B > 65535) throw Error(WT_ERR_VALUE, "wt_save_audio_pcm16: more than 65535 files per call");
        if (B == 0 || T == 0) return WT_OK;
        WT_CUDA(cudaSetDevice(device));
        cudaStream_t s = (cudaStream_t)stream;
        const long long per_block = (long long)PCM_THREADS * PCM_PER_THREAD;
        dim3 grid((unsigned)((T + per_block - 1) / per_block), (unsigned)B);
        if (peak && mode == 2) {
            WT_CUDA(cudaMemsetAsync(peak, 0, (size_t)B * sizeof(float), s));
            absmax_rows_kernel<<<grid, PCM_THREADS, 0, s>>>(wav, T, reinterpret_cast<int*>(peak));
        }
        const bool vec = T % PCM_PER_THREAD == 0 && ((uintptr_t)wav & 15) == 0 && ((uintptr_t)pcm_out & 15) == 0 &&
                         (!limited_out || ((uintptr_t)limited_out & 15) == 0);
        if (vec) limit_pcm16_kernel<true><<<grid, PCM_THREADS, 0, s>>>(wav, T, mode, peak, limited_out, pcm_out);
        else limit_pcm16_kernel<false><<<grid, PCM_THREADS, 0, s>>>(wav, T, mode, peak, limited_out, pcm_out);
        WT_CUDA(cudaGetLastError());
        return WT_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return WT_ERR_RUNTIME;
    }
}

}  // extern "C"
