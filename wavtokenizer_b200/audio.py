"""`convert_audio`: the pre-processing step in front of the hot path.

Mirror of reference encoder/utils.py:79-92 (`from encoder.utils import convert_audio` in README.md:50-112,
infer.py:31-70): same name, argument meaning, assertion texts and output shape. The channel mix and
torchaudio.transforms.Resample(sr, target_sr) run as ONE CUDA kernel behind the C ABI (`wt_convert_audio`,
csrc/audio_ops.cu); there is no CPU / torchaudio fallback: the input must live on a CUDA device.
One difference, documented: mono -> 2 channels works here; the reference hands torchaudio a non-contiguous
`expand()` view and raises.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native


def convert_audio(wav: torch.Tensor, sr: int, target_sr: int, target_channels: int) -> torch.Tensor:
    assert wav.dim() >= 2, "Audio tensor must have at least 2 dimensions"
    assert wav.shape[-2] in [1, 2], "Audio must be mono or stereo."
    *shape, channels, length = wav.shape
    if target_channels not in (1, 2) and channels != 1:
        raise RuntimeError(f"Impossible to convert from {channels} to {target_channels}")
    if target_channels not in (1, 2) and shape:
        # the reference's `wav.expand(target_channels, -1)` only accepts a 2-D [1, T] input here
        raise RuntimeError(f"expand({target_channels}, -1): the number of sizes provided must match the tensor's dims")
    if not wav.is_cuda:
        raise RuntimeError("wavtokenizer_b200.convert_audio runs on a CUDA device only (no CPU fallback)")
    if not wav.is_floating_point():
        raise TypeError(f"Expected floating point type for waveform tensor, but received {wav.dtype}.")
    lib = _native.lib()
    x = wav.to(torch.float32).contiguous()
    B = 1
    for d in shape:
        B *= int(d)
    t_out = int(lib.wt_convert_audio_length(length, int(sr), int(target_sr)))
    out = torch.empty(*shape, target_channels, t_out, dtype=torch.float32, device=wav.device)
    if out.numel():
        stream = ctypes.c_void_p(torch.cuda.current_stream(wav.device).cuda_stream)
        dev = wav.device.index if wav.device.index is not None else torch.cuda.current_device()
        _native.check(lib.wt_convert_audio(dev, x.data_ptr(), B, channels, length, int(sr),
                                           int(target_sr), int(target_channels), out.data_ptr(), stream))
    return out.to(wav.dtype)


_MODES = {"none": 0, "clamp": 1, "rescale": 2}


def pcm16(wav: torch.Tensor, limiter: str = "none", return_limited: bool = False):
    """Rows of ``wav`` [B, T] (fp32, CUDA; one file per row) -> int16 PCM [B, T] through ``wt_save_audio_pcm16``:
    the limiter of reference encoder/utils.py:95-103 (``"clamp"``: rescale=False, ``"rescale"``: rescale=True,
    ``"none"``: the bare torchaudio.save of infer.py:70) fused with the PCM_S-16 conversion of torchaudio.save."""
    if limiter not in _MODES:
        raise ValueError(f"limiter must be one of {sorted(_MODES)}")
    if wav.dim() != 2 or wav.dtype != torch.float32:
        raise ValueError(f"expected float32 [B, T], got {wav.dtype} {tuple(wav.shape)}")
    if not wav.is_cuda:
        raise RuntimeError("wavtokenizer_b200.pcm16 runs on a CUDA device only (no CPU fallback)")
    x = wav.contiguous()
    B, T = x.shape
    out = torch.empty(B, T, dtype=torch.int16, device=x.device)
    peak = torch.empty(B, dtype=torch.float32, device=x.device)
    limited = torch.empty_like(x) if return_limited else None
    if out.numel():
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        _native.check(_native.lib().wt_save_audio_pcm16(dev, x.data_ptr(), B, T, _MODES[limiter], peak.data_ptr(),
                                                        limited.data_ptr() if limited is not None else None,
                                                        out.data_ptr(), stream))
    return (out, limited) if return_limited else out


def save_audio(wav: torch.Tensor, path, sample_rate: int, rescale: bool = False) -> None:
    """Mirror of reference encoder/utils.py:95-103: limit (clamp to +-0.99, or rescale by min(0.99 / max|wav|, 1)
    taken over the WHOLE [channels, T] tensor) and write a 16-bit signed PCM WAV file (channels interleaved, as
    torchaudio.save(..., encoding='PCM_S', bits_per_sample=16) does). The arithmetic runs on the GPU; only the
    RIFF container is written on the host (stdlib ``wave``)."""
    import wave

    if wav.dim() == 1:
        wav = wav.unsqueeze(0)
    if wav.dim() != 2:
        raise ValueError(f"Expected 2D tensor [channels, T], got {wav.dim()}D")
    C, T = wav.shape
    # one file = one row for the kernel, so that the peak spans every channel like `wav.abs().max()`
    q = pcm16(wav.to(torch.float32).reshape(1, C * T), "rescale" if rescale else "clamp").reshape(C, T)
    data = q.t().contiguous().cpu().numpy().tobytes()  # little-endian int16, frame-interleaved
    with wave.open(str(path), "wb") as f:
        f.setnchannels(C)
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(data)
