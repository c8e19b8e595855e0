"""One encode+decode step for profiling under ncu (cudaProfilerStart/Stop bracket the measured step).

    python tools/profile_step.py [--clips B] [--plan P] [--tag small320] [--seconds S]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.gpu_util import native_model  # noqa: E402
from wavtokenizer_b200 import spec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--plan", type=int, default=2)
    ap.add_argument("--tag", default="small320")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--pcm", action="store_true", help="also run the save_audio back-end (clamp, rescale) on the output")
    args = ap.parse_args()
    m = native_model(args.tag, args.plan)
    T = int(args.seconds * 24000)
    wav = spec.synthetic_audio(args.clips, T, seed=1).cuda()
    bw = torch.tensor([0]).cuda()
    for _ in range(2):
        f, c = m.encode_infer(wav, bandwidth_id=bw)
        a = m.decode(m.codes_to_features(c), bandwidth_id=bw)
    torch.cuda.synchronize()
    n0 = m.launch_count()
    torch.cuda.cudart().cudaProfilerStart()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    f, c = m.encode_infer(wav, bandwidth_id=bw)
    a = m.decode(m.codes_to_features(c), bandwidth_id=bw)
    e1.record()
    if args.pcm:
        from wavtokenizer_b200 import pcm16
        pcm16(a, "clamp")
        pcm16(a, "rescale")
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"step: {e0.elapsed_time(e1):.2f} ms for {args.clips} clips x {args.seconds} s, plan {args.plan}, "
          f"{m.launch_count() - n0} launches")


if __name__ == "__main__":
    main()
