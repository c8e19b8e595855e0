"""Host-side logic of the drop-in API (no GPU): module tree / state-dict contract, loaders,
argument checks that mirror the reference's error behaviour, and the no-fallback rule."""
import json
import os

import pytest
import torch

from wavtokenizer_b200 import WavTokenizer, spec
from tests.helpers import GOLDEN, TAGS, config_path, model


@pytest.mark.parametrize("tag", TAGS)
def test_state_dict_keys_and_shapes_match_reference(tag):
    m = WavTokenizer.from_hparams0802(config_path(tag))
    with open(os.path.join(GOLDEN, f"state_keys_{tag}.json")) as f:
        ref = json.load(f)
    hot = {k: v for k, v in ref.items() if not k.startswith(spec.UNUSED_PREFIX)}
    sd = m.state_dict()
    assert set(sd) == set(hot)
    for k, shape in hot.items():
        assert list(sd[k].shape) == shape, k
    assert not m.training


def test_config_values():
    a = spec.load_config(config_path("small600"))
    b = spec.load_config(config_path("small320"))
    assert a.strides == (4, 5, 5, 6) and a.hop == 600 and a.n_fft == 2400 and a.hop_length == 600
    assert b.strides == (2, 4, 5, 8) and b.hop == 320 and b.n_fft == 1280
    assert a.frames_for(72000) == 120 and b.frames_for(72000) == 225
    assert a.frames_for(71999) == 120 and a.frames_for(72001) == 121 and a.frames_for(5) == 1


def test_unsupported_class_path_is_an_error(tmp_path):
    import yaml
    cfg = yaml.safe_load(open(config_path("small600")))
    cfg["model"]["init_args"]["head"]["class_path"] = "decoder.heads.IMDCTSymExpHead"
    p = tmp_path / "bad.yaml"
    p.write_text(yaml.safe_dump(cfg))
    with pytest.raises(ValueError, match="unsupported head.class_path"):
        WavTokenizer.from_hparams0802(str(p))
    cfg = yaml.safe_load(open(config_path("small600")))
    cfg["model"]["init_args"]["head"]["init_args"]["padding"] = "bogus"
    p.write_text(yaml.safe_dump(cfg))
    with pytest.raises(ValueError, match="Padding must be"):
        WavTokenizer.from_hparams0802(str(p))


def test_checkpoint_round_trip_with_foreign_keys(tmp_path):
    cfg, sd = model("small600")
    ckpt = dict(sd)
    ckpt["feature_extractor.encodec.decoder.model.0.conv.conv.bias"] = torch.zeros(7)  # dead SEANet decoder
    ckpt["multiperioddisc.discriminators.0.convs.0.bias"] = torch.zeros(3)             # filtered by prefix
    path = tmp_path / "synthetic.ckpt"
    torch.save({"state_dict": ckpt, "epoch": 3}, path)
    m = WavTokenizer.from_pretrained0802(config_path("small600"), str(path))
    got = m.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    # folder averaging (from_pretrained0911) of two identical checkpoints is the identity
    torch.save({"state_dict": ckpt}, tmp_path / "b.ckpt")
    os.rename(path, tmp_path / "a.ckpt")
    m2 = WavTokenizer.from_pretrained0911(config_path("small600"), str(tmp_path))
    assert torch.equal(m2.state_dict()["head.out.weight"], sd["head.out.weight"])


def test_strict_load_rejects_missing_hot_path_key():
    cfg, sd = model("small600")
    bad = dict(sd)
    bad.pop("head.out.bias")
    with pytest.raises(RuntimeError, match="Missing key"):
        WavTokenizer(cfg).load_state_dict(bad)


def test_reference_attribute_surface():
    cfg, sd = model("small600")
    m = WavTokenizer(cfg)
    m.load_state_dict(sd)
    q = m.feature_extractor.encodec.quantizer
    assert q.bins == 4096
    books = [vq.codebook for vq in q.vq.layers]
    assert len(books) == 1 and books[0].shape == (4096, 512)
    assert torch.equal(books[0], sd[spec.CODEBOOK_PREFIX + "0._codebook.embed"])
    assert m.feature_extractor.bandwidths == [6.6, 6.6, 6.6, 6.6]


def test_no_cpu_fallback():
    cfg, sd = model("small600")
    m = WavTokenizer(cfg)
    m.load_state_dict(sd)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.encode_infer(torch.zeros(1, 1000), bandwidth_id=torch.tensor([0]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.decode(torch.zeros(1, 512, 4), bandwidth_id=torch.tensor([0]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.codes_to_features(torch.zeros(1, 4, dtype=torch.int64))
    with pytest.raises(NotImplementedError):
        m.train()


def test_bandwidth_id_semantics():
    cfg, _ = model("small600")
    m = WavTokenizer(cfg)
    assert m._bandwidth_index({"bandwidth_id": torch.tensor([2])}, True) == 2
    assert m._bandwidth_index({"bandwidth_id": torch.tensor(3)}, False) == 3
    with pytest.raises(TypeError):   # list indexed by a multi-element tensor (feature_extractors.py:137)
        m._bandwidth_index({"bandwidth_id": torch.tensor([0, 0, 0])}, True)
    with pytest.raises(IndexError):
        m._bandwidth_index({"bandwidth_id": torch.tensor([4])}, True)
    with pytest.raises(IndexError):
        m._bandwidth_index({"bandwidth_id": torch.tensor([4])}, False)
    with pytest.raises(AssertionError):  # decoder/models.py:227
        m._bandwidth_index({}, False)
    with pytest.raises(TypeError):
        m._bandwidth_index({}, True)
