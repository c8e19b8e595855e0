"""Generate tests/golden/golden_seanet_dec_<tag>.npz by running the UNMODIFIED reference SEANet decoder.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_seanet_dec.py

``feature_extractor.encodec.decoder`` of the reference model (encoder/modules/seanet.py:147-238, built by
decoder/feature_extractors.py:76-79) is loaded with this repo's seeded synthetic decoder weights through its own
``load_state_dict`` and called on seeded latents. TEST INFRASTRUCTURE ONLY (see oracle/wavtok_oracle.py header).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REF)  # the reference's own decoder/ and encoder/ packages win over the repo's shims
sys.path.insert(1, ROOT)
warnings.filterwarnings("ignore")

from wavtokenizer_b200 import spec  # noqa: E402

CONFIGS = {
    "small600": "wavtokenizer_smalldata_frame40_3s_nq1_code4096_dim512_kmeans200_attn.yaml",
    "small320": "wavtokenizer_smalldata_frame75_3s_nq1_code4096_dim512_kmeans200_attn.yaml",
}
SEED = 7


def latents(B: int, L: int, seed: int) -> torch.Tensor:
    """Encoder-output-like latents: a constant vector plus a small signal (std 0.03)."""
    g = torch.Generator().manual_seed(seed)
    return 0.03 * torch.randn(B, 512, L, generator=g) + 0.02 * torch.randn(1, 512, 1, generator=g)


def main() -> None:
    torch.set_num_threads(8)
    from decoder.pretrained import WavTokenizer as Ref
    for tag, cfg_file in CONFIGS.items():
        path = os.path.join(ROOT, "wavtokenizer_b200", "configs", cfg_file)
        cfg = spec.load_config(path)
        ref = Ref.from_hparams0802(path).eval()
        dec = ref.feature_extractor.encodec.decoder
        sd = spec.synthetic_seanet_decoder(cfg, SEED)
        dec.load_state_dict({k[len(spec.UNUSED_PREFIX):]: v for k, v in sd.items()})
        out = {"weights_seed": np.int64(SEED)}
        for name, (B, L) in {"a": (2, 9), "b": (1, 40), "c": (3, 1)}.items():
            z = latents(B, L, 100 + L)
            with torch.inference_mode():
                y = dec(z)
            assert y.shape == (B, 1, L * cfg.hop_length), y.shape
            out[f"{name}_shape"] = np.array([B, L], dtype=np.int64)
            out[f"{name}_seed"] = np.int64(100 + L)
            out[f"{name}_audio"] = y.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"golden_seanet_dec_{tag}.npz"), **out)
        print(tag, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


if __name__ == "__main__":
    main()
