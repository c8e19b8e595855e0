"""SEANet decoder (reference encoder/modules/seanet.py:147-238, conv.py:214-253): the oracle
restatement against goldens produced by the unmodified reference (oracle/make_golden_seanet_dec.py), and - on the GPU - the
native ``feature_extractor.encodec.decoder`` entry (wt_seanet_decoder) against the same goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import wavtok_oracle as O  # checker only
from tests import helpers
from wavtokenizer_b200 import spec

TAGS = ["small600", "small320"]


def latents(B: int, L: int, seed: int) -> torch.Tensor:  # same generator as oracle/make_golden_seanet_dec.py
    g = torch.Generator().manual_seed(seed)
    return 0.03 * torch.randn(B, 512, L, generator=g) + 0.02 * torch.randn(1, 512, 1, generator=g)


def golden(tag):
    return dict(np.load(os.path.join(helpers.GOLDEN, f"golden_seanet_dec_{tag}.npz")))


@pytest.mark.parametrize("tag", TAGS)
def test_spec_matches_reference_layout(tag):
    cfg = spec.load_config(helpers.config_path(tag))
    shapes = spec.seanet_decoder_spec(cfg)
    assert len(shapes) == 2 * 3 + 8 + 4 * (3 + 9)  # two plain convs, 2-layer LSTM, four (convtr + ResBlock) stages
    first = spec.UNUSED_PREFIX + "model.3.convtr.convtr."
    s0 = list(reversed(cfg.strides))[0]
    assert shapes[first + "weight_v"] == (512, 256, 2 * s0) and shapes[first + "weight_g"] == (512, 1, 1)
    assert shapes[spec.UNUSED_PREFIX + "model.15.conv.conv.weight_v"] == (1, 32, 7)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_goldens(tag):
    cfg = spec.load_config(helpers.config_path(tag))
    g = golden(tag)
    sd = spec.synthetic_seanet_decoder(cfg, int(g["weights_seed"]))
    for name in "abc":
        B, L = (int(v) for v in g[f"{name}_shape"])
        with torch.inference_mode():
            y = O.seanet_decoder(sd, cfg, latents(B, L, int(g[f"{name}_seed"])))
        ref = torch.from_numpy(g[f"{name}_audio"])
        assert y.shape == ref.shape == (B, 1, L * cfg.hop_length)
        assert helpers.snr_db(ref, y) >= 110.0, (name, helpers.snr_db(ref, y))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_native_matches_reference_goldens(tag):
    from tests.gpu_util import native_model
    cfg = spec.load_config(helpers.config_path(tag))
    g = golden(tag)
    m = native_model(tag, 1)
    with pytest.raises(RuntimeError):  # no SEANet-decoder weights loaded yet: a loud error, not a silent fallback
        m.feature_extractor.encodec.decoder(latents(1, 4, 1).cuda())
    full = dict(m.state_dict())
    full.update(spec.synthetic_seanet_decoder(cfg, int(g["weights_seed"])))
    m.load_state_dict(full)
    for name in "abc":
        B, L = (int(v) for v in g[f"{name}_shape"])
        z = latents(B, L, int(g[f"{name}_seed"])).cuda()
        y = m.feature_extractor.encodec.decoder(z)
        ref = torch.from_numpy(g[f"{name}_audio"])
        assert tuple(y.shape) == tuple(ref.shape) and y.device.type == "cuda"
        assert helpers.snr_db(ref, y.cpu()) >= 100.0, (name, helpers.snr_db(ref, y.cpu()))
    # the hot path is untouched by the extra weights
    wav = spec.synthetic_audio(1, 4800, seed=3).cuda()
    f1, c1 = m.encode_infer(wav, bandwidth_id=torch.tensor([0]).cuda())
    assert c1.shape[-1] == cfg.frames_for(4800)
