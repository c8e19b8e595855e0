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


def test_from_pretrained0911_averages_the_three_best_vocos_checkpoints(tmp_path, monkeypatch):
    """Reference decoder/pretrained.py:117-156: only 'vocos_*' files, ranked by the val-loss string in the file name
    ([-11:-5]), best three averaged; anything else in the folder is ignored."""
    cfg, sd = model("small600")
    names = {"vocos_checkpoint_epoch=1_step=10_val_loss=5.3000.ckpt": 3.0,
             "vocos_checkpoint_epoch=2_step=20_val_loss=5.1000.ckpt": 1.0,
             "vocos_checkpoint_epoch=3_step=30_val_loss=9.9000.ckpt": 100.0,   # fourth best: dropped
             "vocos_checkpoint_epoch=4_step=40_val_loss=5.2000.ckpt": 2.0,
             "last.ckpt": 1000.0, "notes.txt": 1000.0}                          # no 'vocos_' prefix: ignored
    for n in names:
        (tmp_path / n).write_bytes(b"")
    real_load = torch.load

    def fake_load(path, *a, **k):
        base = os.path.basename(str(path))
        if base not in names:
            return real_load(path, *a, **k)
        d = {k2: v.clone() for k2, v in sd.items()}
        d["head.out.bias"] = torch.full_like(sd["head.out.bias"], names[base])
        d["discriminator.x"] = torch.zeros(1)
        return {"state_dict": d}
    monkeypatch.setattr(torch, "load", fake_load)
    m = WavTokenizer.from_pretrained0911(config_path("small600"), str(tmp_path))
    got = m.state_dict()
    assert torch.equal(got["head.out.bias"], torch.full_like(sd["head.out.bias"], 2.0))  # mean of 1, 2, 3
    assert torch.allclose(got["head.out.weight"], sd["head.out.weight"], rtol=0, atol=1e-7)
    assert not m.training
    empty = tmp_path / "empty"
    empty.mkdir()
    (empty / "model.ckpt").write_bytes(b"")
    with pytest.raises(FileNotFoundError):
        WavTokenizer.from_pretrained0911(config_path("small600"), str(empty))


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


def test_reference_import_paths_resolve_to_the_native_classes():
    """`from decoder.pretrained import WavTokenizer` / `from encoder.utils import convert_audio` (reference
    README.md:50-53, infer.py:5-6) work unedited through the repo's shim packages."""
    import importlib
    import wavtokenizer_b200
    dp = importlib.import_module("decoder.pretrained")
    eu = importlib.import_module("encoder.utils")
    assert dp.WavTokenizer is wavtokenizer_b200.WavTokenizer
    assert eu.convert_audio is wavtokenizer_b200.convert_audio and eu.save_audio is wavtokenizer_b200.save_audio
    for name in ("from_pretrained0802", "from_pretrained0911", "from_hparams0802", "encode_infer", "codes_to_features",
                 "decode", "encode", "forward"):
        assert hasattr(dp.WavTokenizer, name), name
