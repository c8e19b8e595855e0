"""CPU oracle for the WavTokenizer inference hot path.  TEST INFRASTRUCTURE ONLY.

A functional, state-dict-driven restatement (plain PyTorch CPU ops, fp32 or fp64) of the
reference path  encode_infer -> VQ -> codes_to_features -> decode.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product (``wavtokenizer_b200``) never does and fails loudly when
its CUDA library is missing.

Parity status: PINNED.  The reference ships no golden vectors of its own (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself:
``oracle/make_golden.py`` imports the unmodified reference from /root/reference, runs it
on seeded weights/inputs and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those files, and its
``test_live_reference_small`` re-runs the live comparison whenever the reference is present
(/root/reference, or the copy ``oracle/fetch_ref.py`` places under the git-ignored oracle/_ref).

Every function cites the reference file:line it restates (paths relative to
/root/reference).  Layout convention here is the reference's: activations [B, C, T].
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
ENC = "feature_extractor.encodec.encoder.model."
CB = "feature_extractor.encodec.quantizer.vq.layers."


def _w(sd: Dict[str, Tensor], name: str, dtype) -> Tensor:
    return sd[name].to(dtype)


# ----------------------------------------------------------------------------------
# encoder
# ----------------------------------------------------------------------------------

def weight_norm_fold(g: Tensor, v: Tensor) -> Tensor:
    """Old-style ``torch.nn.utils.weight_norm`` (dim=0): w = g * v / ||v||, norm over
    (C_in, k) per output channel (encoder/modules/conv.py:25-34, 115)."""
    return g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)


def extra_padding(length: int, kernel_size: int, stride: int, padding_total: int) -> int:
    """encoder/modules/conv.py:54-61 (float division + ceil, kept as written)."""
    n_frames = (length - kernel_size + padding_total) / stride + 1
    ideal_length = (math.ceil(n_frames) - 1) * stride + (kernel_size - padding_total)
    return ideal_length - length


def pad1d_reflect(x: Tensor, left: int, right: int) -> Tensor:
    """encoder/modules/conv.py:79-96: reflect pad; when the signal is not longer than the
    pad, zero-pad on the right first and crop afterwards."""
    length = x.shape[-1]
    max_pad = max(left, right)
    extra = 0
    if length <= max_pad:
        extra = max_pad - length + 1
        x = F.pad(x, (0, extra))
    padded = F.pad(x, (left, right), mode="reflect")
    end = padded.shape[-1] - extra
    return padded[..., :end]


def sconv1d(sd: Dict[str, Tensor], prefix: str, x: Tensor, stride: int = 1) -> Tensor:
    """SConv1d.forward, non-causal branch (encoder/modules/conv.py:195-211) on a
    weight-normed Conv1d (conv.py:108-122). Dilation is 1 everywhere on this path
    (seanet.py:115: dilation_base ** 0)."""
    dt = x.dtype
    w = weight_norm_fold(_w(sd, prefix + "conv.conv.weight_g", dt), _w(sd, prefix + "conv.conv.weight_v", dt))
    b = _w(sd, prefix + "conv.conv.bias", dt)
    k = w.shape[-1]
    padding_total = k - stride
    extra = extra_padding(x.shape[-1], k, stride, padding_total)
    right = padding_total // 2
    left = padding_total - right
    x = pad1d_reflect(x, left, right + extra)
    return F.conv1d(x, w, b, stride=stride)


def seanet_resblock(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """SEANetResnetBlock.forward (encoder/modules/seanet.py:45-63): shortcut 1x1 conv of x
    plus conv1x1(ELU(conv_k3(ELU(x)))); true_skip=False (feature_extractors.py:71-74)."""
    h = sconv1d(sd, prefix + "block.1.", F.elu(x))
    h = sconv1d(sd, prefix + "block.3.", F.elu(h))
    return sconv1d(sd, prefix + "shortcut.", x) + h


def slstm(sd: Dict[str, Tensor], prefix: str, x: Tensor, layers: int = 2) -> Tensor:
    """SLSTM.forward (encoder/modules/lstm.py:31-39): nn.LSTM(512, 512, 2), zero initial
    state, gate order i,f,g,o, biases b_ih + b_hh, then skip connection y + x."""
    B, C, T = x.shape
    dt = x.dtype
    seq = x.permute(2, 0, 1)  # [T, B, C]
    for layer in range(layers):
        w_ih = _w(sd, f"{prefix}lstm.weight_ih_l{layer}", dt)
        w_hh = _w(sd, f"{prefix}lstm.weight_hh_l{layer}", dt)
        bias = _w(sd, f"{prefix}lstm.bias_ih_l{layer}", dt) + _w(sd, f"{prefix}lstm.bias_hh_l{layer}", dt)
        xin = seq @ w_ih.t() + bias  # hoisted input projection
        h = torch.zeros(B, C, dtype=dt)
        c = torch.zeros(B, C, dtype=dt)
        outs = []
        for t in range(T):
            gates = xin[t] + h @ w_hh.t()
            i, f, g, o = gates.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        seq = torch.stack(outs, dim=0)
    return seq.permute(1, 2, 0) + x


def slstm_library(sd: Dict[str, Tensor], prefix: str, x: Tensor, layers: int = 2) -> Tensor:
    """Same SLSTM through ``torch.nn.LSTM`` itself — the library op the reference calls
    (encoder/modules/lstm.py:20) — used when the oracle is timed as the CPU baseline so that the
    baseline is not slowed down by the explicit Python time loop of ``slstm``."""
    C = x.shape[1]
    lstm = torch.nn.LSTM(C, C, layers)
    with torch.no_grad():
        for layer in range(layers):
            for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                getattr(lstm, f"{nm}_l{layer}").copy_(sd[f"{prefix}lstm.{nm}_l{layer}"])
    lstm = lstm.to(x.dtype)
    y, _ = lstm(x.permute(2, 0, 1))
    return y.permute(1, 2, 0) + x


def seanet_encoder(sd: Dict[str, Tensor], cfg, audio: Tensor, library_lstm: bool = False) -> Tensor:
    """SEANetEncoder.forward (encoder/modules/seanet.py:105-144) on audio [B, 1, T];
    returns z [B, 512, L]."""
    x = sconv1d(sd, ENC + "0.", audio)
    idx = 1
    for s in cfg.strides:
        x = seanet_resblock(sd, f"{ENC}{idx}.", x)
        x = sconv1d(sd, f"{ENC}{idx + 2}.", F.elu(x), stride=s)
        idx += 3
    x = (slstm_library if library_lstm else slstm)(sd, f"{ENC}{idx}.", x, cfg.lstm_layers)
    x = sconv1d(sd, f"{ENC}{idx + 2}.", F.elu(x))
    return x


def sconvtr1d(sd: Dict[str, Tensor], prefix: str, x: Tensor, stride: int) -> Tensor:
    """SConvTranspose1d.forward, non-causal branch (encoder/modules/conv.py:232-253) on a weight-normed
    ConvTranspose1d (conv.py:125-139; weight_norm over dim 0 = input channels): full transposed conv, then the fixed
    padding kernel - stride is trimmed, right = total // 2, left = total - right."""
    dt = x.dtype
    w = weight_norm_fold(_w(sd, prefix + "convtr.convtr.weight_g", dt), _w(sd, prefix + "convtr.convtr.weight_v", dt))
    b = _w(sd, prefix + "convtr.convtr.bias", dt)
    k = w.shape[-1]
    y = F.conv_transpose1d(x, w, b, stride=stride)
    total = k - stride
    right = total // 2
    left = total - right
    return y[..., left: y.shape[-1] - right]


def seanet_decoder(sd: Dict[str, Tensor], cfg, z: Tensor, library_lstm: bool = True) -> Tensor:
    """SEANetDecoder.forward (encoder/modules/seanet.py:189-238) on z [B, 512, L]; returns audio [B, 1, L * hop].
    SURVEY.md section 8(f) row 4: next to the hot path, not on it."""
    DEC = "feature_extractor.encodec.decoder.model."
    x = sconv1d(sd, DEC + "0.", z)
    x = (slstm_library if library_lstm else slstm)(sd, DEC + "1.", x, cfg.lstm_layers)
    idx = 2
    for s in reversed(list(cfg.strides)):
        x = sconvtr1d(sd, f"{DEC}{idx + 1}.", F.elu(x), s)
        x = seanet_resblock(sd, f"{DEC}{idx + 2}.", x)
        idx += 3
    return sconv1d(sd, f"{DEC}{idx + 1}.", F.elu(x))


# ----------------------------------------------------------------------------------
# vector quantiser
# ----------------------------------------------------------------------------------

def vq_scores(x: Tensor, embed: Tensor) -> Tensor:
    """EuclideanCodebook.quantize (encoder/quantization/core_vq.py:175-182): the negative
    expanded squared distance, term order as written."""
    e = embed.t()
    return -(x.pow(2).sum(1, keepdim=True) - 2 * x @ e + e.pow(2).sum(0, keepdim=True))


def vq_quantize(x: Tensor, embed: Tensor, chunk: int = 8192) -> Tensor:
    """argmax of ``vq_scores`` along the codebook (core_vq.py:182; first index on ties).
    Chunked over rows only to bound memory; rows are independent."""
    out = []
    for i in range(0, x.shape[0], chunk):
        out.append(vq_scores(x[i:i + chunk], embed).max(dim=-1).indices)
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.int64)


def vq_infer(sd: Dict[str, Tensor], z: Tensor) -> Tuple[Tensor, Tensor]:
    """ResidualVectorQuantizer.infer with n_q forced to 1 (encoder/quantization/vq.py:115-140)
    -> LanguageVectorQuantization.forward (core_vq.py:378-401) -> VectorQuantization.forward
    eval branch (core_vq.py:294-315) -> EuclideanCodebook.forward (core_vq.py:206-231).
    z [B, 512, L] -> (quantized [B, 512, L], codes [1, B, L] int64)."""
    embed = _w(sd, CB + "0._codebook.embed", z.dtype)
    B, D, L = z.shape
    flat = z.permute(0, 2, 1).reshape(B * L, D)
    idx = vq_quantize(flat, embed)
    quant = F.embedding(idx, embed).view(B, L, D).permute(0, 2, 1)
    return quant, idx.view(1, B, L)


def codes_to_features(sd: Dict[str, Tensor], cfg, codes: Tensor) -> Tensor:
    """WavTokenizer.codes_to_features (decoder/pretrained.py:209-239): offset embedding
    lookup in the concatenated codebooks, summed over K, transposed to [B, 512, L]."""
    if codes.dim() == 2:
        codes = codes.unsqueeze(1)
    K = codes.shape[0]
    n_bins = cfg.vq_bins
    offsets = torch.arange(0, n_bins * K, n_bins)
    table = torch.cat([sd[f"{CB}{q}._codebook.embed"] for q in range(cfg.num_quantizers)], dim=0)
    feats = F.embedding(codes + offsets.view(-1, 1, 1), table).sum(dim=0)
    return feats.transpose(1, 2)


def encode_infer(sd: Dict[str, Tensor], cfg, audio: Tensor, bandwidth_id=None,
                 dtype=torch.float32, library_lstm: bool = False) -> Tuple[Tensor, Tensor]:
    """EncodecFeatures.infer (decoder/feature_extractors.py:131-142) behind
    WavTokenizer.encode_infer (decoder/pretrained.py:186-189). ``bandwidth_id`` only
    indexes a Python list there and the looked-up value is ignored (vq.py:126-137)."""
    z = seanet_encoder(sd, cfg, audio.to(dtype).unsqueeze(1), library_lstm)
    return vq_infer(sd, z)


# ----------------------------------------------------------------------------------
# decoder backbone
# ----------------------------------------------------------------------------------

def _group_norm(sd, prefix: str, x: Tensor) -> Tensor:
    """Normalize (decoder/models.py:15-16): GroupNorm(32, C, eps=1e-6, affine)."""
    return F.group_norm(x, 32, _w(sd, prefix + "weight", x.dtype), _w(sd, prefix + "bias", x.dtype), eps=1e-6)


def _swish(x: Tensor) -> Tensor:
    """decoder/models.py:10-12."""
    return x * torch.sigmoid(x)


def resnet_block(sd, prefix: str, x: Tensor) -> Tensor:
    """ResnetBlock.forward with temb=None, in==out channels, dropout in eval
    (decoder/models.py:58-78)."""
    dt = x.dtype
    h = _swish(_group_norm(sd, prefix + "norm1.", x))
    h = F.conv1d(h, _w(sd, prefix + "conv1.weight", dt), _w(sd, prefix + "conv1.bias", dt), padding=1)
    h = _swish(_group_norm(sd, prefix + "norm2.", h))
    h = F.conv1d(h, _w(sd, prefix + "conv2.weight", dt), _w(sd, prefix + "conv2.bias", dt), padding=1)
    return x + h


def attn_block(sd, prefix: str, x: Tensor) -> Tensor:
    """AttnBlock.forward (decoder/models.py:107-127): single-head softmax attention over
    time with scale C^-0.5, written with the same bmm/permute order."""
    dt = x.dtype
    h = _group_norm(sd, prefix + "norm.", x)
    q = F.conv1d(h, _w(sd, prefix + "q.weight", dt), _w(sd, prefix + "q.bias", dt))
    k = F.conv1d(h, _w(sd, prefix + "k.weight", dt), _w(sd, prefix + "k.bias", dt))
    v = F.conv1d(h, _w(sd, prefix + "v.weight", dt), _w(sd, prefix + "v.bias", dt))
    c = q.shape[1]
    w_ = torch.bmm(q.permute(0, 2, 1), k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2).permute(0, 2, 1)
    h = torch.bmm(v, w_)
    h = F.conv1d(h, _w(sd, prefix + "proj_out.weight", dt), _w(sd, prefix + "proj_out.bias", dt))
    return x + h


def ada_layer_norm(sd, prefix: str, x: Tensor, bandwidth_id: Tensor) -> Tensor:
    """AdaLayerNorm.forward (decoder/modules.py:81-86) on [B, L, C]; one id for the batch."""
    dt = x.dtype
    scale = F.embedding(bandwidth_id, _w(sd, prefix + "scale.weight", dt))
    shift = F.embedding(bandwidth_id, _w(sd, prefix + "shift.weight", dt))
    x = F.layer_norm(x, (x.shape[-1],), eps=1e-6)
    return x * scale + shift


def convnext_block(sd, prefix: str, x: Tensor, bandwidth_id: Tensor) -> Tensor:
    """ConvNeXtBlock.forward (decoder/modules.py:43-60): depthwise k7 -> AdaLN ->
    Linear -> exact-erf GELU -> Linear -> gamma -> residual."""
    dt = x.dtype
    C = x.shape[1]
    h = F.conv1d(x, _w(sd, prefix + "dwconv.weight", dt), _w(sd, prefix + "dwconv.bias", dt), padding=3, groups=C)
    h = ada_layer_norm(sd, prefix + "norm.", h.transpose(1, 2), bandwidth_id)
    h = F.linear(h, _w(sd, prefix + "pwconv1.weight", dt), _w(sd, prefix + "pwconv1.bias", dt))
    h = F.gelu(h)
    h = F.linear(h, _w(sd, prefix + "pwconv2.weight", dt), _w(sd, prefix + "pwconv2.bias", dt))
    h = _w(sd, prefix + "gamma", dt) * h
    return x + h.transpose(1, 2)


def vocos_backbone(sd, cfg, features: Tensor, bandwidth_id: Tensor) -> Tensor:
    """VocosBackbone.forward (decoder/models.py:223-235): [B, 512, L] -> [B, L, 768]."""
    dt = features.dtype
    x = F.conv1d(features, _w(sd, "backbone.embed.weight", dt), _w(sd, "backbone.embed.bias", dt), padding=3)
    x = resnet_block(sd, "backbone.pos_net.0.", x)
    x = resnet_block(sd, "backbone.pos_net.1.", x)
    x = attn_block(sd, "backbone.pos_net.2.", x)
    x = resnet_block(sd, "backbone.pos_net.3.", x)
    x = resnet_block(sd, "backbone.pos_net.4.", x)
    x = _group_norm(sd, "backbone.pos_net.5.", x)
    x = ada_layer_norm(sd, "backbone.norm.", x.transpose(1, 2), bandwidth_id).transpose(1, 2)
    for i in range(cfg.num_layers):
        x = convnext_block(sd, f"backbone.convnext.{i}.", x, bandwidth_id)
    x = F.layer_norm(x.transpose(1, 2), (cfg.dim,), _w(sd, "backbone.final_layer_norm.weight", dt),
                     _w(sd, "backbone.final_layer_norm.bias", dt), eps=1e-6)
    return x


# ----------------------------------------------------------------------------------
# head
# ----------------------------------------------------------------------------------

def window_envelope(window: Tensor, frames: int, hop: int) -> Tensor:
    """Overlap-added squared window over ``frames`` frames, trimmed by (win-hop)//2 each
    side (decoder/spectral_ops.py:65-69), computed by direct accumulation."""
    n_fft = window.shape[0]
    pad = (n_fft - hop) // 2
    size = (frames - 1) * hop + n_fft
    env = torch.zeros(size, dtype=window.dtype)
    wsq = window.square()
    for t in range(frames):
        env[t * hop:t * hop + n_fft] += wsq
    return env[pad:size - pad]


def istft_same(spec: Tensor, window: Tensor, n_fft: int, hop: int) -> Tensor:
    """ISTFT.forward, padding='same' (decoder/spectral_ops.py:47-75): irfft(norm
    'backward') * window -> overlap-add (fold) -> trim -> divide by the window envelope."""
    B, N, T = spec.shape
    pad = (n_fft - hop) // 2
    frames = torch.fft.irfft(spec, n_fft, dim=1, norm="backward") * window[None, :, None]
    size = (T - 1) * hop + n_fft
    y = torch.zeros(B, size, dtype=frames.dtype)
    for t in range(T):
        y[:, t * hop:t * hop + n_fft] += frames[:, :, t]
    y = y[:, pad:size - pad]
    env = window_envelope(window, T, hop)
    assert (env > 1e-11).all()  # spectral_ops.py:72
    return y / env


def istft_head(sd, cfg, x: Tensor) -> Tensor:
    """ISTFTHead.forward (decoder/heads.py:42-67): Linear -> (log-mag | phase) halves ->
    exp, clip(max=1e2) -> mag*(cos p + i sin p) -> ISTFT."""
    dt = x.dtype
    y = F.linear(x, _w(sd, "head.out.weight", dt), _w(sd, "head.out.bias", dt)).transpose(1, 2)
    mag, p = y.chunk(2, dim=1)
    mag = torch.clip(torch.exp(mag), max=1e2)
    spec = torch.complex(mag * torch.cos(p), mag * torch.sin(p))
    return istft_same(spec, _w(sd, "head.istft.window", dt), cfg.n_fft, cfg.hop_length)


def decode(sd, cfg, features: Tensor, bandwidth_id: Tensor, dtype=torch.float32) -> Tensor:
    """WavTokenizer.decode (decoder/pretrained.py:192-207): head(backbone(features))."""
    bw = torch.as_tensor(bandwidth_id).reshape(-1)[:1].to(torch.int64)
    x = vocos_backbone(sd, cfg, features.to(dtype), bw)
    return istft_head(sd, cfg, x)


# ----------------------------------------------------------------------------------
# tie accounting for the code-match metric
# ----------------------------------------------------------------------------------

def vq_tie_report(z_rows: Tensor, embed: Tensor, codes_a: Tensor, codes_b: Tensor,
                  rel_gap: float = 1e-5) -> dict:
    """Classify disagreements between two code assignments of the same rows (a = the path under test, b = the fp32
    reference / oracle). True squared distances are evaluated in fp64.

    Two definitions of a near-tie are reported side by side:
    * term magnitude (`near_ties`, `worst_rel_gap`): |d(a) - d(b)| < rel_gap * (||x||^2 + ||c||^2). The expanded
      formula the reference evaluates in fp32 (core_vq.py:177-181) cancels at the scale of its TERMS, not of the
      distance, so this is the resolution below which the reference's own choice is rounding noise (SURVEY.md
      section 7, hard part 2). `hard_mismatches` counts the flips above it.
    * distance (`near_ties_rel_distance`, `worst_rel_gap_distance`): |d(a) - d(b)| < rel_gap * min(d(a), d(b)),
      BASELINE.json north_star's literal wording ("distance gap < 1e-5 relative").
    `a_closer_fp64` counts the flips where the code under test is at least as close to the frame as the reference's
    code in fp64, i.e. where the fp32 reference itself did not pick the true nearest code.
    """
    a = codes_a.reshape(-1).to(torch.int64)
    b = codes_b.reshape(-1).to(torch.int64)
    n = a.numel()
    diff = (a != b).nonzero().flatten()
    near = near_d = closer = 0
    worst = worst_d = 0.0
    if diff.numel():
        x = z_rows[diff].double()
        ca = embed[a[diff]].double()
        cb = embed[b[diff]].double()
        da = (x - ca).pow(2).sum(1)
        db = (x - cb).pow(2).sum(1)
        mag = x.pow(2).sum(1) + torch.maximum(ca.pow(2).sum(1), cb.pow(2).sum(1))
        gap = (da - db).abs()
        rel = gap / mag
        rel_d = gap / torch.minimum(da, db).clamp_min(1e-300)
        near = int((rel < rel_gap).sum())
        near_d = int((rel_d < rel_gap).sum())
        closer = int((da <= db).sum())
        worst = float(rel.max())
        worst_d = float(rel_d.max())
    return dict(frames=n, mismatches=int(diff.numel()), near_ties=near,
                hard_mismatches=int(diff.numel()) - near,
                match_pct=100.0 * (n - diff.numel()) / max(n, 1), worst_rel_gap=worst,
                near_ties_rel_distance=near_d, worst_rel_gap_distance=worst_d, a_closer_fp64=closer)
