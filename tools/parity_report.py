"""Parity report at benchmark scale: the CUDA path vs the CPU oracle on the same seeded clips (one JSON line per config).

    python tools/parity_report.py [--clips 256] [--plan 2] [--tags small320,small600,medium]

Reports BASELINE.json's metric (iii): code match % with near-tie accounting (fp64 distances, gap < 1e-5 relative),
feature / waveform max-abs error and SNR against the fp32 oracle. The oracle runs on the host in chunks of 16 clips.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import wavtok_oracle as O  # noqa: E402  (checker)
from tests import helpers  # noqa: E402
from tests.gpu_util import native_model  # noqa: E402
from wavtokenizer_b200 import spec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=256)
    ap.add_argument("--plan", type=int, default=2)
    ap.add_argument("--tags", default="small320,small600,medium")
    ap.add_argument("--seconds", type=float, default=3.0)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    T = int(args.seconds * 24000)
    for tag in args.tags.split(","):
        cfg, sd = helpers.model(tag)
        m = native_model(tag, args.plan)
        cb = sd[spec.CODEBOOK_PREFIX + "0._codebook.embed"]
        wav = spec.synthetic_audio(args.clips, T, seed=2024)
        bw_id = 1
        bw = torch.tensor([bw_id])
        feats, codes = m.encode_infer(wav.cuda(), bandwidth_id=bw.cuda())
        audio = m.decode(feats, bandwidth_id=bw.cuda())
        z_nat = m._encoder_forward(wav.cuda()).cpu()
        codes, audio = codes.cpu(), audio.cpu()
        t0 = time.perf_counter()
        zs, cs, auds = [], [], []
        with torch.inference_mode():
            for i in range(0, args.clips, 16):
                z = O.seanet_encoder(sd, cfg, wav[i:i + 16].unsqueeze(1), library_lstm=True)
                _, c = O.vq_infer(sd, z)
                # decode the NATIVE codes with the oracle: isolates the decoder's own error from code flips
                a = O.decode(sd, cfg, O.codes_to_features(sd, cfg, codes[:, i:i + 16]), bw)
                zs.append(z), cs.append(c), auds.append(a)
        z, c_ref, a_ref = torch.cat(zs), torch.cat(cs, dim=1), torch.cat(auds)
        rep = O.vq_tie_report(z.permute(0, 2, 1).reshape(-1, cfg.dimension), cb, codes, c_ref)
        err = (a_ref - audio).double()
        zerr = (z - z_nat).double()
        line = {"config": tag, "plan": args.plan, "clips": args.clips, "frames": int(rep["frames"]),
                "code_match_pct": round(float(rep["match_pct"]), 5), "mismatches": int(rep["mismatches"]),
                "near_ties": int(rep["near_ties"]), "hard_mismatches": int(rep["hard_mismatches"]),
                "unique_codes": int(codes.unique().numel()),
                "latent_snr_db": round(helpers.snr_db(z, z_nat), 2), "latent_max_abs_err": float(zerr.abs().max()),
                "waveform_snr_db": round(helpers.snr_db(a_ref, audio), 2), "waveform_max_abs_err": float(err.abs().max()),
                "waveform_peak": float(a_ref.abs().max()), "oracle_seconds": round(time.perf_counter() - t0, 1)}
        print(json.dumps(line), flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
