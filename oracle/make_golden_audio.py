"""Generate tests/golden/convert_audio.npz from the UNMODIFIED reference `encoder.utils.convert_audio`.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_audio.py

/root/reference does not exist on the GPU box, so the outputs are committed as a small fixture. Inputs are
regenerated from the seeds stored alongside. TEST INFRASTRUCTURE ONLY.
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

# (name, sr, channels, T, target_sr, target_channels, leading batch dims)
CASES = [
    ("cd_stereo_to_mono", 44100, 2, 13247, 24000, 1, (2,)),
    ("48k_mono", 48000, 1, 9601, 24000, 1, (3,)),
    ("16k_up", 16000, 1, 4001, 24000, 1, ()),
    ("22k_stereo_keep", 22050, 2, 5000, 24000, 2, (1,)),
    # (mono -> 2 channels is not in the fixture: the reference's `wav.expand(...)` hands torchaudio a non-contiguous
    #  tensor and `_apply_sinc_resample_kernel`'s `.view` raises; the native path supports it, tested vs the oracle)
    ("same_rate_mix", 24000, 2, 2500, 24000, 1, (2,)),
    ("short", 44100, 1, 37, 24000, 1, ()),
]


def make_input(i: int, lead, channels: int, T: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(1234 + i)
    return torch.randn(*lead, channels, T, generator=g).clamp(-1, 1)


def main() -> None:
    from encoder.utils import convert_audio  # the reference, unmodified
    out = {}
    for i, (name, sr, ch, T, tsr, tch, lead) in enumerate(CASES):
        x = make_input(i, lead, ch, T)
        y = convert_audio(x, sr, tsr, tch)
        out[name] = y.numpy().astype(np.float32)
        print(name, tuple(x.shape), "->", tuple(y.shape))
    path = os.path.join(ROOT, "tests", "golden", "convert_audio.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
