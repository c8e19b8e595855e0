"""Worker of tests/test_multi_gpu.py (launched by torch.distributed.run, one rank per GPU, NCCL):
sharded encode + all-gather of the codes must equal the single-GPU codes bit for bit (SURVEY.md section 4 item 4,
section 8(e))."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.gpu_util import native_model  # noqa: E402
from wavtokenizer_b200 import spec  # noqa: E402
from wavtokenizer_b200.shard import gather_codes, shard_range  # noqa: E402


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_items = int(sys.argv[1]) if len(sys.argv) > 1 else 37  # not a multiple of the world size: ragged shards
    tag = sys.argv[2] if len(sys.argv) > 2 else "small320"
    from tests import helpers
    cfg, sd = helpers.model(tag)
    from wavtokenizer_b200 import WavTokenizer
    m = WavTokenizer(cfg)
    m.load_state_dict(sd)
    m = m.to(dev)
    m.set_plan(2)
    wav = spec.synthetic_audio(n_items, 24000, seed=321)
    bw = torch.tensor([0], device=dev)
    s, e = shard_range(n_items, rank, world)
    _, local_codes = m.encode_infer(wav[s:e].to(dev), bandwidth_id=bw)
    allc = gather_codes(local_codes, n_items, bins=cfg.vq_bins)
    ok = torch.ones(1, device=dev)
    if rank == 0:
        _, single = m.encode_infer(wav.to(dev), bandwidth_id=bw)
        same = torch.equal(allc, single)
        print(f"MGPU_EQUAL world={world} clips={n_items} frames={single.numel()} equal={same}", flush=True)
        ok[0] = 1.0 if same else 0.0
    dist.broadcast(ok, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
