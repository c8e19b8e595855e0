"""Model configuration and checkpoint contract of the WavTokenizer inference path.

Mirrors what the reference resolves from YAML through ``instantiate_class``
(reference decoder/pretrained.py:13-29, 81-92) and the hard-wired SEANet encoder
arguments in ``EncodecFeatures.__init__`` (decoder/feature_extractors.py:55-96).
Nothing here imports the reference; the state-dict key set below is checked against
the reference's own ``state_dict()`` by tests/test_oracle_golden.py::test_state_keys_match_reference (golden key list).
"""
from __future__ import annotations

import dataclasses
import math
import zlib
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch
import yaml

# class paths the reference YAMLs name; anything else is an error.
_FE_CLASS = "decoder.feature_extractors.EncodecFeatures"
_BB_CLASS = "decoder.models.VocosBackbone"
_HEAD_CLASS = "decoder.heads.ISTFTHead"

ENC_PREFIX = "feature_extractor.encodec.encoder.model."
CODEBOOK_PREFIX = "feature_extractor.encodec.quantizer.vq.layers."
UNUSED_PREFIX = "feature_extractor.encodec.decoder."  # SEANet decoder: in checkpoints, never run


@dataclasses.dataclass(frozen=True)
class ModelConfig:
    """Everything the native path needs to know about one WavTokenizer variant."""

    ratios: Tuple[int, ...]            # `dowmsamples` as written in YAML (encoder applies them reversed)
    bandwidths: Tuple[float, ...]
    num_quantizers: int = 1
    vq_bins: int = 4096
    vq_kmeans: int = 200
    n_filters: int = 32
    dimension: int = 512
    lstm_layers: int = 2
    input_channels: int = 512
    dim: int = 768
    intermediate_dim: int = 2304
    num_layers: int = 12
    adanorm_num_embeddings: int = 4
    n_fft: int = 1280
    hop_length: int = 320
    padding: str = "same"
    sample_rate: int = 24000

    @property
    def strides(self) -> Tuple[int, ...]:
        """Down-sampling strides in execution order (seanet.py:100 reverses the YAML list)."""
        return tuple(reversed(self.ratios))

    @property
    def hop(self) -> int:
        return int(math.prod(self.ratios))

    def frames_for(self, num_samples: int) -> int:
        """L = ceil(ceil(ceil(ceil(T/s1)/s2)/s3)/s4) (conv.py:54-61, 195-211)."""
        n = int(num_samples)
        for s in self.strides:
            n = -(-n // s)
        return n


def _section(cfg: dict, key: str, expected_class: str) -> dict:
    node = cfg["model"]["init_args"][key]
    if node.get("class_path") != expected_class:
        raise ValueError(
            f"unsupported {key}.class_path {node.get('class_path')!r}: the native path implements "
            f"{expected_class} only (no fallback)")
    return node.get("init_args", {}) or {}


def load_config(config_path: str) -> ModelConfig:
    """Parse a reference-style YAML (decoder/pretrained.py:86-90)."""
    with open(config_path, "r") as f:
        cfg = yaml.safe_load(f)
    fe = _section(cfg, "feature_extractor", _FE_CLASS)
    bb = _section(cfg, "backbone", _BB_CLASS)
    hd = _section(cfg, "head", _HEAD_CLASS)
    if fe.get("encodec_model", "encodec_24khz") != "encodec_24khz":
        # feature_extractors.py:87-90
        raise ValueError(f"Unsupported encodec_model: {fe.get('encodec_model')}. "
                         "Supported options are 'encodec_24khz'.")
    padding = hd.get("padding", "same")
    if padding not in ("center", "same"):
        raise ValueError("Padding must be 'center' or 'same'.")  # spectral_ops.py:24-25
    if padding != "same":
        raise ValueError("the native ISTFT implements padding='same' only (all reference configs)")
    ada = bb.get("adanorm_num_embeddings")
    if not ada:
        raise ValueError("the native backbone requires adanorm_num_embeddings (all reference configs set 4)")
    mc = ModelConfig(
        ratios=tuple(int(r) for r in fe.get("dowmsamples", [6, 5, 5, 4])),
        bandwidths=tuple(float(b) for b in fe.get("bandwidths", [1.5, 3.0, 6.0, 12.0])),
        num_quantizers=int(fe.get("num_quantizers", 1)),
        vq_bins=int(fe.get("vq_bins", 16384)),
        vq_kmeans=int(fe.get("vq_kmeans", 800)),
        input_channels=int(bb["input_channels"]),
        dim=int(bb["dim"]),
        intermediate_dim=int(bb["intermediate_dim"]),
        num_layers=int(bb["num_layers"]),
        adanorm_num_embeddings=int(ada),
        n_fft=int(hd["n_fft"]),
        hop_length=int(hd["hop_length"]),
        padding=padding,
    )
    if len(mc.ratios) != 4:
        raise ValueError("expected 4 down-sampling ratios")
    if mc.input_channels != mc.dimension or int(hd["dim"]) != mc.dim:
        raise ValueError("inconsistent channel sizes between feature extractor, backbone and head")
    if mc.dim % 32 != 0:
        raise ValueError("GroupNorm(32) requires dim % 32 == 0")
    return mc


# --------------------------------------------------------------------------------------
# checkpoint contract
# --------------------------------------------------------------------------------------

def encoder_layout(cfg: ModelConfig) -> List[dict]:
    """The 16-entry ``encoder.model`` Sequential (seanet.py:105-141) as plain records."""
    out: List[dict] = []
    ch = cfg.n_filters
    out.append(dict(idx=0, kind="conv", cin=1, cout=ch, k=7, stride=1))
    idx = 1
    for s in cfg.strides:
        out.append(dict(idx=idx, kind="resblock", dim=ch, hidden=ch // 2))
        out.append(dict(idx=idx + 1, kind="elu"))
        out.append(dict(idx=idx + 2, kind="conv", cin=ch, cout=2 * ch, k=2 * s, stride=s))
        ch *= 2
        idx += 3
    out.append(dict(idx=idx, kind="lstm", dim=ch, layers=cfg.lstm_layers))
    out.append(dict(idx=idx + 1, kind="elu"))
    out.append(dict(idx=idx + 2, kind="conv", cin=ch, cout=cfg.dimension, k=7, stride=1))
    return out


def state_spec(cfg: ModelConfig) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """name -> (shape, 'param'|'buffer') for every tensor the hot path reads.

    Key names follow the reference's ``state_dict()`` (old-style weight_norm:
    ``weight_g``/``weight_v``, conv.py:25-34). The SEANet *decoder* keys
    (``feature_extractor.encodec.decoder.*``) are accepted and ignored by the loader.
    """
    spec: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()

    def wn_conv(prefix: str, cout: int, cin: int, k: int) -> None:
        spec[prefix + "conv.conv.bias"] = ((cout,), "param")
        spec[prefix + "conv.conv.weight_g"] = ((cout, 1, 1), "param")
        spec[prefix + "conv.conv.weight_v"] = ((cout, cin, k), "param")

    for ent in encoder_layout(cfg):
        p = f"{ENC_PREFIX}{ent['idx']}."
        if ent["kind"] == "conv":
            wn_conv(p, ent["cout"], ent["cin"], ent["k"])
        elif ent["kind"] == "resblock":
            wn_conv(p + "block.1.", ent["hidden"], ent["dim"], 3)
            wn_conv(p + "block.3.", ent["dim"], ent["hidden"], 1)
            wn_conv(p + "shortcut.", ent["dim"], ent["dim"], 1)
        elif ent["kind"] == "lstm":
            d = ent["dim"]
            for layer in range(ent["layers"]):
                spec[f"{p}lstm.weight_ih_l{layer}"] = ((4 * d, d), "param")
                spec[f"{p}lstm.weight_hh_l{layer}"] = ((4 * d, d), "param")
                spec[f"{p}lstm.bias_ih_l{layer}"] = ((4 * d,), "param")
                spec[f"{p}lstm.bias_hh_l{layer}"] = ((4 * d,), "param")
    for q in range(cfg.num_quantizers):
        p = f"{CODEBOOK_PREFIX}{q}._codebook."
        spec[p + "inited"] = ((1,), "buffer")
        spec[p + "cluster_size"] = ((cfg.vq_bins,), "buffer")
        spec[p + "embed"] = ((cfg.vq_bins, cfg.dimension), "buffer")
        spec[p + "embed_avg"] = ((cfg.vq_bins, cfg.dimension), "buffer")

    D, H, E = cfg.dim, cfg.intermediate_dim, cfg.adanorm_num_embeddings
    spec["backbone.embed.weight"] = ((D, cfg.input_channels, 7), "param")
    spec["backbone.embed.bias"] = ((D,), "param")
    spec["backbone.norm.scale.weight"] = ((E, D), "param")
    spec["backbone.norm.shift.weight"] = ((E, D), "param")
    for i in range(cfg.num_layers):
        p = f"backbone.convnext.{i}."
        spec[p + "gamma"] = ((D,), "param")
        spec[p + "dwconv.weight"] = ((D, 1, 7), "param")
        spec[p + "dwconv.bias"] = ((D,), "param")
        spec[p + "norm.scale.weight"] = ((E, D), "param")
        spec[p + "norm.shift.weight"] = ((E, D), "param")
        spec[p + "pwconv1.weight"] = ((H, D), "param")
        spec[p + "pwconv1.bias"] = ((H,), "param")
        spec[p + "pwconv2.weight"] = ((D, H), "param")
        spec[p + "pwconv2.bias"] = ((D,), "param")
    spec["backbone.final_layer_norm.weight"] = ((D,), "param")
    spec["backbone.final_layer_norm.bias"] = ((D,), "param")
    for i in range(5):
        p = f"backbone.pos_net.{i}."
        if i == 2:
            spec[p + "norm.weight"] = ((D,), "param")
            spec[p + "norm.bias"] = ((D,), "param")
            for nm in ("q", "k", "v", "proj_out"):
                spec[p + nm + ".weight"] = ((D, D, 1), "param")
                spec[p + nm + ".bias"] = ((D,), "param")
            continue
        for j in (1, 2):
            spec[p + f"norm{j}.weight"] = ((D,), "param")
            spec[p + f"norm{j}.bias"] = ((D,), "param")
            spec[p + f"conv{j}.weight"] = ((D, D, 3), "param")
            spec[p + f"conv{j}.bias"] = ((D,), "param")
    spec["backbone.pos_net.5.weight"] = ((D,), "param")
    spec["backbone.pos_net.5.bias"] = ((D,), "param")
    spec["head.out.weight"] = ((cfg.n_fft + 2, D), "param")
    spec["head.out.bias"] = ((cfg.n_fft + 2,), "param")
    spec["head.istft.window"] = ((cfg.n_fft,), "buffer")
    # canonical order of the reference: feature_extractor, backbone, head — already so.
    return spec


# --------------------------------------------------------------------------------------
# synthetic (random-init) weights: no checkpoints are available offline
# --------------------------------------------------------------------------------------

def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFF)
    return g


def synthetic_state_dict(cfg: ModelConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic random-init weights with the reference's initial statistics.

    Distributions follow the reference constructors (kaiming-uniform convs/LSTM with
    bound 1/sqrt(fan_in); trunc-normal(0.02) backbone, models.py:218-221; pos_net keeps
    torch defaults because it is created after ``apply(_init_weights)``, models.py:196-216),
    with every trivially-initialised affine term (AdaLN scale/shift, layer-scale gamma,
    LN/GN affine, zero biases) perturbed so that it is observable in parity tests
    (SURVEY.md section 7.1). Each tensor has its own generator keyed by (seed, name), so
    the result does not depend on construction order. The codebook is left at zeros with
    ``inited = 0`` exactly like a fresh reference model: callers must install one
    (see ``install_codebook``) before encoding.
    """
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, (shape, _kind) in state_spec(cfg).items():
        g = _gen(seed, name)

        def uni(bound: float) -> torch.Tensor:
            return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

        def nrm(std: float, mean: float = 0.0) -> torch.Tensor:
            return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean

        leaf = name.rsplit(".", 1)[1]
        if name.startswith(ENC_PREFIX):
            if leaf == "weight_v":
                fan_in = shape[1] * shape[2]
                t = uni(1.0 / math.sqrt(fan_in))
            elif leaf == "weight_g":
                t = None  # filled after weight_v (g = ||v|| * (1 + 0.1 n))
            elif leaf == "bias":
                vshape = state_spec_shape_cache(cfg, name[: -len("bias")] + "weight_v")
                t = uni(1.0 / math.sqrt(vshape[1] * vshape[2]))
            else:  # lstm
                t = uni(1.0 / math.sqrt(cfg.dimension))
        elif name.startswith(CODEBOOK_PREFIX):
            t = torch.zeros(shape, dtype=torch.float32)
        elif name == "head.istft.window":
            t = torch.hann_window(cfg.n_fft)  # spectral_ops.py:30 (periodic)
        elif ".pos_net." in name:
            if "norm" in name or name.startswith("backbone.pos_net.5."):
                t = nrm(0.1, 1.0) if leaf == "weight" else nrm(0.1)
            elif leaf == "weight":
                t = uni(1.0 / math.sqrt(shape[1] * shape[2]))
            else:
                wshape = state_spec_shape_cache(cfg, name[: -len("bias")] + "weight")
                t = uni(1.0 / math.sqrt(wshape[1] * wshape[2]))
        elif leaf == "gamma":
            t = nrm(0.02, 1.0 / cfg.num_layers)
        elif ".scale." in name:
            t = nrm(0.1, 1.0)
        elif ".shift." in name:
            t = nrm(0.1)
        elif name.startswith("backbone.final_layer_norm."):
            t = nrm(0.1, 1.0) if leaf == "weight" else nrm(0.1)
        elif leaf == "weight":
            t = torch.nn.init.trunc_normal_(torch.empty(shape, dtype=torch.float32), std=0.02, generator=g)
        elif leaf == "bias":
            t = nrm(0.02)
        else:
            raise AssertionError(name)
        sd[name] = t
    for name in list(sd):
        if name.endswith("weight_g"):
            v = sd[name[: -len("weight_g")] + "weight_v"]
            norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
            sd[name] = norm * (1.0 + 0.1 * torch.randn(norm.shape, generator=_gen(seed, name)))
    return sd


_SPEC_CACHE: Dict[ModelConfig, "OrderedDict[str, Tuple[Tuple[int, ...], str]]"] = {}


def state_spec_shape_cache(cfg: ModelConfig, name: str) -> Tuple[int, ...]:
    if cfg not in _SPEC_CACHE:
        _SPEC_CACHE[cfg] = state_spec(cfg)
    return _SPEC_CACHE[cfg][name][0]


def install_codebook(sd: Dict[str, torch.Tensor], codebook: torch.Tensor, layer: int = 0) -> None:
    """Set the VQ codebook and mark it initialised.

    A fresh reference model has ``embed = 0`` and ``inited = 0`` and would run k-means
    inside ``infer`` (core_vq.py:140-151); every user of random-init weights installs a
    codebook first (SURVEY.md section 8(c) caveat 1).
    """
    p = f"{CODEBOOK_PREFIX}{layer}._codebook."
    cb = codebook.detach().to(torch.float32).cpu().contiguous()
    if tuple(cb.shape) != tuple(sd[p + "embed"].shape):
        raise ValueError(f"codebook shape {tuple(cb.shape)} != {tuple(sd[p + 'embed'].shape)}")
    sd[p + "embed"] = cb.clone()
    sd[p + "embed_avg"] = cb.clone()
    sd[p + "cluster_size"] = torch.ones(cb.shape[0], dtype=torch.float32)
    sd[p + "inited"] = torch.ones(1, dtype=torch.float32)


def synthetic_audio(batch: int, num_samples: int, seed: int = 1) -> torch.Tensor:
    """Unit-variance noise clamped to [-1, 1] (SURVEY.md section 8(d) synthetic inputs)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return torch.randn(batch, num_samples, generator=g, dtype=torch.float32).clamp_(-1.0, 1.0)


def expand_codebook(base_rows_bf16: torch.Tensor, bins: int, seed: int = 5, jitter: float = 1e-3) -> torch.Tensor:
    """Build a [bins, D] fp32 codebook from a small table of bf16-exact base rows.

    Row i = base[i % n_base] + jitter * N(0, 1) (seeded). The base rows are encoder-output
    frames sampled from a calibration batch — the first step of the reference's own k-means
    initialisation (core_vq.py:63-71, 77) — stored in bf16 so that test fixtures stay small;
    a plain Gaussian codebook is degenerate on random-init weights.
    """
    base = base_rows_bf16.to(torch.float32)
    n_base, dim = base.shape
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    noise = torch.randn(bins, dim, generator=g, dtype=torch.float32) * jitter
    idx = torch.arange(bins) % n_base
    return (base[idx] + noise).contiguous()

# ----------------------------------------------------------------------------------
# SEANet decoder: feature_extractor.encodec.decoder, the waveform decoder every reference
# checkpoint carries (decoder/feature_extractors.py:76-79) and the fork's enhancement experiments call. NOT on the
# WavTokenizer hot path: its weights are optional (loaded when the checkpoint has them, run through wt_seanet_decoder).
# ----------------------------------------------------------------------------------
def seanet_decoder_layout(cfg: ModelConfig) -> List[dict]:
    """The 16-entry ``decoder.model`` Sequential (reference encoder/modules/seanet.py:189-238) as plain records; the
    decoder walks the ratios in the constructor's order, i.e. the encoder's strides reversed."""
    out: List[dict] = []
    ch = cfg.n_filters * 2 ** len(cfg.strides)
    out.append(dict(idx=0, kind="conv", cin=cfg.dimension, cout=ch, k=7))
    out.append(dict(idx=1, kind="lstm", dim=ch, layers=cfg.lstm_layers))
    idx = 2
    for s in reversed(list(cfg.strides)):
        out.append(dict(idx=idx, kind="elu"))
        out.append(dict(idx=idx + 1, kind="convtr", cin=ch, cout=ch // 2, k=2 * s, stride=s))
        out.append(dict(idx=idx + 2, kind="resblock", dim=ch // 2, hidden=ch // 4))
        ch //= 2
        idx += 3
    out.append(dict(idx=idx, kind="elu"))
    out.append(dict(idx=idx + 1, kind="conv", cin=ch, cout=1, k=7))
    return out


def seanet_decoder_spec(cfg: ModelConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape of the SEANet decoder's parameters, named as in the reference's ``state_dict()``
    (old-style weight_norm; for ConvTranspose1d the norm runs over dim 0 = INPUT channels, conv.py:25-34, 125-139)."""
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def wn_conv(prefix: str, cout: int, cin: int, k: int) -> None:
        spec[prefix + "conv.conv.bias"] = (cout,)
        spec[prefix + "conv.conv.weight_g"] = (cout, 1, 1)
        spec[prefix + "conv.conv.weight_v"] = (cout, cin, k)

    for ent in seanet_decoder_layout(cfg):
        p = f"{UNUSED_PREFIX}model.{ent['idx']}."
        if ent["kind"] == "conv":
            wn_conv(p, ent["cout"], ent["cin"], ent["k"])
        elif ent["kind"] == "convtr":
            spec[p + "convtr.convtr.bias"] = (ent["cout"],)
            spec[p + "convtr.convtr.weight_g"] = (ent["cin"], 1, 1)
            spec[p + "convtr.convtr.weight_v"] = (ent["cin"], ent["cout"], ent["k"])
        elif ent["kind"] == "resblock":
            wn_conv(p + "block.1.", ent["hidden"], ent["dim"], 3)
            wn_conv(p + "block.3.", ent["dim"], ent["hidden"], 1)
            wn_conv(p + "shortcut.", ent["dim"], ent["dim"], 1)
        elif ent["kind"] == "lstm":
            d = ent["dim"]
            for layer in range(ent["layers"]):
                spec[f"{p}lstm.weight_ih_l{layer}"] = (4 * d, d)
                spec[f"{p}lstm.weight_hh_l{layer}"] = (4 * d, d)
                spec[f"{p}lstm.bias_ih_l{layer}"] = (4 * d,)
                spec[f"{p}lstm.bias_hh_l{layer}"] = (4 * d,)
    return spec


def synthetic_seanet_decoder(cfg: ModelConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random-init SEANet-decoder weights with the constructors' statistics (uniform +-1/sqrt(fan_in);
    weight_g = ||weight_v|| per norm slice, jittered so that the fold is observable)."""
    shapes = seanet_decoder_spec(cfg)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in shapes.items():
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[1]
        if leaf == "weight_g":
            continue
        if leaf == "weight_v":
            fan_in = shape[1] * shape[2]
        elif leaf == "bias":
            v = shapes[name[: -len("bias")] + "weight_v"]
            fan_in = v[1] * v[2]
        else:
            fan_in = shape[-1] if leaf.startswith("weight") else shapes[name.replace("bias_", "weight_")][-1]
        sd[name] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) / math.sqrt(fan_in)
    for name, shape in shapes.items():
        if name.endswith("weight_g"):
            v = sd[name[: -len("weight_g")] + "weight_v"]
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape)
            g = _gen(seed, name)
            sd[name] = nrm * (1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32))
    return OrderedDict((k, sd[k]) for k in shapes)
