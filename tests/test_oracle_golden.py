"""Pin the CPU oracle (oracle/wavtok_oracle.py) against outputs of the unmodified reference
committed under tests/golden/ by oracle/make_golden.py."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import wavtok_oracle as O
from wavtokenizer_b200 import spec
from tests.helpers import GOLDEN, TAGS, golden, model, snr_db


@pytest.mark.parametrize("tag", TAGS)
def test_weights_reproduce(tag):
    cfg, sd = model(tag)
    g = golden(tag)
    names = ["feature_extractor.encodec.encoder.model.0.conv.conv.weight_v", "backbone.embed.weight",
             "head.out.weight", "feature_extractor.encodec.encoder.model.13.lstm.weight_hh_l1"]
    got = np.array([float(sd[n].double().sum()) for n in names])
    np.testing.assert_allclose(got, g["weight_checksum"], rtol=0, atol=1e-9)
    cb = sd[spec.CODEBOOK_PREFIX + "0._codebook.embed"]
    assert abs(float(cb.double().sum()) - float(g["codebook_checksum"])) < 1e-9


@pytest.mark.parametrize("tag", TAGS)
def test_state_keys_match_reference(tag):
    cfg, sd = model(tag)
    with open(os.path.join(GOLDEN, f"state_keys_{tag}.json")) as f:
        ref_keys = json.load(f)
    hot = {k: v for k, v in ref_keys.items() if not k.startswith(spec.UNUSED_PREFIX)}
    assert set(hot) == set(sd)
    for k, shape in hot.items():
        assert list(sd[k].shape) == shape, k


@pytest.mark.parametrize("tag", TAGS)
def test_e2e_matches_reference(tag):
    cfg, sd = model(tag)
    g = golden(tag)
    T = int(g["e2e_T"])
    wav = spec.synthetic_audio(2, T, seed=11)
    bw = torch.tensor([2])
    with torch.inference_mode():
        z = O.seanet_encoder(sd, cfg, wav.unsqueeze(1))
        feats, codes = O.vq_infer(sd, z)
        audio = O.decode(sd, cfg, feats, bw)
    assert snr_db(torch.from_numpy(g["e2e_z"]), z) > 110
    ref_codes = torch.from_numpy(g["e2e_codes"].astype(np.int64))
    rep = O.vq_tie_report(z.permute(0, 2, 1).reshape(-1, 512), sd[spec.CODEBOOK_PREFIX + "0._codebook.embed"],
                          codes, ref_codes)
    assert rep["hard_mismatches"] == 0, rep
    assert rep["match_pct"] >= 99.0, rep
    assert codes.shape == (1, 2, cfg.frames_for(T)) and codes.dtype == torch.int64
    assert torch.equal(O.codes_to_features(sd, cfg, codes), feats)
    # decode on the reference's own codes so a tie flip does not leak into the waveform check
    with torch.inference_mode():
        audio = O.decode(sd, cfg, O.codes_to_features(sd, cfg, ref_codes), bw)
    assert audio.shape == g["e2e_audio"].shape
    assert snr_db(torch.from_numpy(g["e2e_audio"]), audio) > 100


@pytest.mark.parametrize("tag", TAGS)
def test_edges_and_bandwidths(tag):
    cfg, sd = model(tag)
    g = golden(tag)
    for T_e in [int(x) for x in g["edge_lengths"]]:
        w = spec.synthetic_audio(3, T_e, seed=13 + T_e)
        with torch.inference_mode():
            feats, codes = O.encode_infer(sd, cfg, w)
            ref_codes = torch.from_numpy(g[f"edge{T_e}_codes"].astype(np.int64))
            audio = O.decode(sd, cfg, O.codes_to_features(sd, cfg, ref_codes), torch.tensor(1))
        assert codes.shape == ref_codes.shape
        assert (codes != ref_codes).float().mean() <= 0.02
        assert audio.shape == g[f"edge{T_e}_audio"].shape == (3, cfg.frames_for(T_e) * cfg.hop_length)
        assert snr_db(torch.from_numpy(g[f"edge{T_e}_audio"]), audio) > 100
    rc = torch.from_numpy(g["bw_codes"].astype(np.int64))
    with torch.inference_mode():
        rf = O.codes_to_features(sd, cfg, rc)
        outs = [O.decode(sd, cfg, rf, torch.tensor([b])) for b in range(4)]
    for b in range(4):
        assert snr_db(torch.from_numpy(g[f"bw{b}_audio_sub4"]), outs[b][:, ::4]) > 100
    assert snr_db(outs[0], outs[1]) < 40  # bandwidth_id is observable


def _reference_present() -> bool:
    from oracle import fetch_ref
    try:
        fetch_ref.reference_root()
        return True
    except ImportError:
        return False


@pytest.mark.skipif(not _reference_present(), reason="reference checkout (/root/reference or oracle/_ref) not present")
def test_live_reference_small():
    """Re-run the comparison against the live, unmodified reference (this container: /root/reference; the GPU box:
    the copy oracle/fetch_ref.py made under oracle/_ref). The repo's own `decoder` shim package must not shadow it."""
    import decoder.pretrained as shim  # the drop-in shim ...
    from oracle import fetch_ref
    from oracle import wavtok_oracle as O
    from tests.helpers import config_path, model
    from wavtokenizer_b200 import WavTokenizer, spec
    cfg, sd = model("small600")
    ref = fetch_ref.load_reference(config_path("small600"), sd)
    assert type(ref).__module__ == "decoder.pretrained" and type(ref) is not WavTokenizer  # ... and the real one
    import decoder.pretrained as shim_again
    assert shim_again is shim and shim.WavTokenizer is WavTokenizer  # the shim is back in place afterwards
    wav = spec.synthetic_audio(2, 7001, seed=99)
    with torch.inference_mode():
        f, c = ref.encode_infer(wav, bandwidth_id=torch.tensor([3]))
        a = ref.decode(f, bandwidth_id=torch.tensor([3]))
        f2, c2 = O.encode_infer(sd, cfg, wav)
        a2 = O.decode(sd, cfg, f, torch.tensor([3]))
    assert (c != c2).float().mean() <= 0.02, (c != c2).float().mean()
    err = (a - a2).abs().max().item()
    assert err < 1e-5 * max(1.0, a.abs().max().item()), err
