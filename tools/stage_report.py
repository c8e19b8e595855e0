"""Per-stage parity report on the GPU: every tapped stage of the native path against the
reference goldens (tests/golden), without stopping at the first mismatch.

    python tools/stage_report.py [tag ...] [--plan N]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import wavtok_oracle as O  # noqa: E402  (checker)
from tests import helpers  # noqa: E402
from tests.gpu_util import Taps, golden_sub, native_model  # noqa: E402
from wavtokenizer_b200 import spec  # noqa: E402


def main():
    argv = sys.argv[1:]
    plan = 0
    if "--plan" in argv:
        i = argv.index("--plan")
        plan = int(argv[i + 1])
        del argv[i:i + 2]
    args = argv
    tags = args or ["small600"]
    for tag in tags:
        cfg, sd = helpers.model(tag)
        g = helpers.golden(tag)
        m = native_model(tag, plan)
        names = [str(n) for n in g["tap_names"]]
        skip = {"enc2", "enc5", "enc8", "enc11", "enc14"}  # ELU outputs are fused into the next loader
        taps = Taps(m, [n for n in names if n not in skip])
        T = int(g["e2e_T"])
        wav = spec.synthetic_audio(2, T, seed=11).cuda()
        bw = torch.tensor([2]).cuda()
        feats, codes = m.encode_infer(wav, bandwidth_id=bw)
        ref_codes = torch.from_numpy(g["e2e_codes"].astype(np.int64))
        audio = m.decode(m.codes_to_features(ref_codes.cuda()), bandwidth_id=bw)
        torch.cuda.synchronize()
        print(f"== {tag} plan {plan}: T={T} L={codes.shape[-1]}")
        for n in names:
            if n in skip:
                continue
            try:
                t = taps.get(n)
            except KeyError as e:
                print(f"  {n:12s} MISSING {e}")
                continue
            ref = torch.from_numpy(g["tap_" + n])
            shape = tuple(int(x) for x in g["tapshape_" + n])
            ok_shape = tuple(t.shape) == shape
            got = golden_sub(t)
            snr = helpers.snr_db(ref, got) if ok_shape else float("nan")
            print(f"  {n:12s} shape {tuple(t.shape)} {'ok ' if ok_shape else 'BAD'+str(shape)} "
                  f"SNR {snr:7.1f} dB  max|err| {(ref - got).abs().max().item() if ok_shape else float('nan'):.3e} "
                  f"max|ref| {ref.abs().max().item():.3e}")
        z = taps.get("enc15")
        print(f"  z vs golden z: SNR {helpers.snr_db(torch.from_numpy(g['e2e_z']), z):.1f} dB")
        rep = O.vq_tie_report(z.permute(0, 2, 1).reshape(-1, cfg.dimension),
                              sd[spec.CODEBOOK_PREFIX + "0._codebook.embed"], codes.cpu(), ref_codes)
        print("  codes:", rep)
        print(f"  features == codebook[codes]: "
              f"{torch.equal(feats.cpu(), O.codes_to_features(sd, cfg, codes.cpu()))}")
        ga = torch.from_numpy(g["e2e_audio"])
        print(f"  audio: shape {tuple(audio.shape)} SNR {helpers.snr_db(ga, audio.cpu()):.1f} dB "
              f"max|err| {(ga - audio.cpu()).abs().max().item():.3e} max|ref| {ga.abs().max().item():.3e}")
        print(f"  launches: {m.launch_count()}")
        taps.close()


if __name__ == "__main__":
    main()
