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
