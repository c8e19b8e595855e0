"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference (pure Python/PyTorch) is imported as-is, loaded with this repo's seeded
synthetic weights through its own ``load_state_dict`` and driven through its own public API
(``encode_infer`` / ``codes_to_features`` / ``decode``, decoder/pretrained.py:186-239).
Outputs are committed as small fixtures because /root/reference does not exist on the GPU box.
TEST INFRASTRUCTURE ONLY (see oracle/wavtok_oracle.py header).
"""
from __future__ import annotations

import json
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

from wavtokenizer_b200 import spec  # noqa: E402

CONFIGS = {
    "small600": ("wavtokenizer_smalldata_frame40_3s_nq1_code4096_dim512_kmeans200_attn.yaml", 0),
    "small320": ("wavtokenizer_smalldata_frame75_3s_nq1_code4096_dim512_kmeans200_attn.yaml", 1),
    "medium": ("wavtokenizer_mediumdata_music_audio_frame75_3s_nq1_code4096_dim512_kmeans200_attn.yaml", 2),
}
N_BASE = 512
TAP_SAMPLES = 384


def sub(t: torch.Tensor, n: int = TAP_SAMPLES) -> np.ndarray:
    """Deterministic strided subsample of a flattened tensor."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].to(torch.float32).numpy().copy()


def build_reference(cfg_file: str, seed: int):
    from decoder.pretrained import WavTokenizer as Ref
    path = os.path.join(ROOT, "wavtokenizer_b200", "configs", cfg_file)
    cfg = spec.load_config(path)
    ref = Ref.from_hparams0802(path).eval()
    sd = spec.synthetic_state_dict(cfg, seed)
    full = dict(ref.state_dict())
    full.update(sd)
    ref.load_state_dict(full)
    return cfg, ref, sd


def main() -> None:
    torch.set_num_threads(8)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for tag, (cfg_file, seed) in CONFIGS.items():
        cfg, ref, sd = build_reference(cfg_file, seed)
        rsd = ref.state_dict()
        with open(os.path.join(out_dir, f"state_keys_{tag}.json"), "w") as f:
            json.dump({k: list(v.shape) for k, v in rsd.items()}, f, indent=0)

        # ---- codebook: sampled encoder frames of a calibration batch (reference encoder) ----
        cal = spec.synthetic_audio(36, 72000, seed=7)
        with torch.inference_mode():
            zc = torch.cat([ref.feature_extractor.encodec.encoder(cal[i:i + 4].unsqueeze(1)) for i in range(0, 36, 4)])
        frames = zc.permute(0, 2, 1).reshape(-1, cfg.dimension)
        g = torch.Generator().manual_seed(5)
        base = frames[torch.randperm(frames.shape[0], generator=g)[:N_BASE]].to(torch.bfloat16)
        codebook = spec.expand_codebook(base, cfg.vq_bins, seed=5)
        spec.install_codebook(sd, codebook)
        full = dict(ref.state_dict())
        full.update(sd)
        ref.load_state_dict(full)

        gold = {"codebook_base_bf16": base.view(torch.int16).numpy().copy(),
                "weights_seed": np.int64(seed)}
        # weight checksums: detect RNG drift between machines
        names = ["feature_extractor.encodec.encoder.model.0.conv.conv.weight_v", "backbone.embed.weight",
                 "head.out.weight", "feature_extractor.encodec.encoder.model.13.lstm.weight_hh_l1"]
        gold["weight_checksum"] = np.array([float(sd[n].double().sum()) for n in names])
        gold["codebook_checksum"] = np.float64(codebook.double().sum())

        # ---- taps through forward hooks on the reference modules ----
        taps = {}
        hooks = []
        enc = ref.feature_extractor.encodec.encoder.model
        for i, m in enumerate(enc):
            hooks.append(m.register_forward_hook(lambda _m, _i, o, i=i: taps.__setitem__(f"enc{i}", o)))
        bb = ref.backbone
        hooks.append(bb.embed.register_forward_hook(lambda _m, _i, o: taps.__setitem__("dec_embed", o)))
        for i, m in enumerate(bb.pos_net):
            hooks.append(m.register_forward_hook(lambda _m, _i, o, i=i: taps.__setitem__(f"dec_pos{i}", o)))
        hooks.append(bb.norm.register_forward_hook(lambda _m, _i, o: taps.__setitem__("dec_norm", o.transpose(1, 2))))
        for i, m in enumerate(bb.convnext):
            hooks.append(m.register_forward_hook(lambda _m, _i, o, i=i: taps.__setitem__(f"dec_cnx{i}", o)))
        hooks.append(bb.final_layer_norm.register_forward_hook(
            lambda _m, _i, o: taps.__setitem__("dec_final", o.transpose(1, 2))))
        hooks.append(ref.head.out.register_forward_hook(
            lambda _m, _i, o: taps.__setitem__("dec_headlin", o.transpose(1, 2))))

        # ---- case "e2e": two ~1 s clips, odd length, bandwidth_id 2 ----
        T = 24123
        wav = spec.synthetic_audio(2, T, seed=11)
        bw = torch.tensor([2])
        with torch.inference_mode():
            feats, codes = ref.encode_infer(wav, bandwidth_id=bw)
            c2f = ref.codes_to_features(codes)
            audio = ref.decode(feats, bandwidth_id=bw)
        assert torch.equal(c2f, feats)
        gold["e2e_T"] = np.int64(T)
        gold["e2e_codes"] = codes.numpy().astype(np.int16)
        gold["e2e_audio"] = audio.numpy()
        gold["e2e_z"] = taps[f"enc{len(enc) - 1}"].numpy()
        tap_names = sorted(taps.keys())
        gold["tap_names"] = np.array(tap_names)
        for k in tap_names:
            gold["tap_" + k] = sub(taps[k])
            gold["tapshape_" + k] = np.array(taps[k].shape, dtype=np.int64)
        for h in hooks:
            h.remove()

        # ---- case "3s": one full 3 s clip (BASELINE.json configs[0] shape) ----
        wav3 = spec.synthetic_audio(1, 72000, seed=12)
        with torch.inference_mode():
            f3, c3 = ref.encode_infer(wav3, bandwidth_id=torch.tensor([0]))
            a3 = ref.decode(f3, bandwidth_id=torch.tensor([0]))
        gold["c3s_codes"] = c3.numpy().astype(np.int16)
        gold["c3s_audio_sub16"] = a3.numpy()[:, ::16].copy()
        gold["c3s_audio_absmax"] = np.float64(a3.abs().max())

        # ---- edge lengths: T < pad (L = 1), short, T = k*hop + 1 ----
        for T_e in (5, 1000, 4 * cfg.hop_length + 1):
            w = spec.synthetic_audio(3, T_e, seed=13 + T_e)
            with torch.inference_mode():
                fe, ce = ref.encode_infer(w, bandwidth_id=torch.tensor([1]))
                ae = ref.decode(fe, bandwidth_id=torch.tensor(1))  # 0-dim id is accepted too
            gold[f"edge{T_e}_codes"] = ce.numpy().astype(np.int16)
            gold[f"edge{T_e}_audio"] = ae.numpy()
        gold["edge_lengths"] = np.array([5, 1000, 4 * cfg.hop_length + 1], dtype=np.int64)

        # ---- bandwidth ids 0..3 on the same features; decode-only from random codes ----
        g2 = torch.Generator().manual_seed(21)
        rc = torch.randint(0, cfg.vq_bins, (1, 2, 50), generator=g2)
        with torch.inference_mode():
            rf = ref.codes_to_features(rc)
            for b in range(4):
                gold[f"bw{b}_audio_sub4"] = ref.decode(rf, bandwidth_id=torch.tensor([b])).numpy()[:, ::4].copy()
            rf2 = ref.codes_to_features(rc[:, 0])  # 2-D [K, L] form
        gold["bw_codes"] = rc.numpy().astype(np.int16)
        gold["c2f_2d_equal_3d"] = np.bool_(torch.equal(rf2[0], rf[0]))

        path = os.path.join(out_dir, f"golden_{tag}.npz")
        np.savez_compressed(path, **gold)
        print(tag, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB",
              "unique codes e2e:", int(codes.unique().numel()), "/", codes.numel())


if __name__ == "__main__":
    main()
