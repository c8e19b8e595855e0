"""Generate tests/golden/save_audio.npz from the UNMODIFIED reference `encoder.utils.save_audio`.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_save_audio.py

`torchaudio.save` is intercepted (the installed torchaudio cannot write files without torchcodec, and the file
writer is third-party code anyway): the fixture holds the float tensor the reference hands to it, i.e. the output of
the reference's own limiter lines (encoder/utils.py:97-102). Inputs are regenerated from seeds.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
# the reference's own decoder/ and encoder/ packages must win over the repo's import-path shims of the same name
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
warnings.filterwarnings("ignore")

# (name, channels, T, amplitude, rescale)
CASES = [
    ("loud_clamp", 1, 5003, 1.7, False),
    ("loud_rescale", 1, 5003, 1.7, True),
    ("quiet_clamp", 1, 4096, 0.3, False),
    ("quiet_rescale", 1, 4096, 0.3, True),      # 0.99 / max > 1 -> multiplied by the python int 1
    ("stereo_rescale", 2, 2500, 2.5, True),     # the peak spans both channels
    ("edge_rescale", 1, 777, 0.99, True),
]


def make_input(i: int, channels: int, T: int, amp: float) -> torch.Tensor:
    g = torch.Generator().manual_seed(4321 + i)
    x = torch.randn(channels, T, generator=g) * (amp / 3.0)
    x[0, T // 2] = amp  # the peak is exactly `amp`
    return x


def main() -> None:
    import torchaudio
    import encoder.utils as U  # the reference, unmodified

    captured = {}

    def fake_save(path, wav, sample_rate, encoding=None, bits_per_sample=None, **kw):
        assert encoding == "PCM_S" and bits_per_sample == 16
        captured["wav"], captured["sr"] = wav.detach().clone(), sample_rate

    real = torchaudio.save
    torchaudio.save = fake_save
    out = {}
    try:
        for i, (name, ch, T, amp, rescale) in enumerate(CASES):
            x = make_input(i, ch, T, amp)
            U.save_audio(x, "/tmp/unused.wav", 24000, rescale=rescale)
            assert captured["sr"] == 24000
            out[name] = captured["wav"].numpy().astype(np.float32)
            print(name, tuple(x.shape), "max|x|", float(x.abs().max()), "-> max|y|", float(np.abs(out[name]).max()))
    finally:
        torchaudio.save = real
    path = os.path.join(ROOT, "tests", "golden", "save_audio.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
