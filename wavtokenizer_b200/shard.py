"""Clip sharding across the GPUs of one box.

Every clip is independent end to end, so rank r of W simply owns a contiguous slice of the
clips; nothing crosses GPUs on the data path except one all-gather of the codes at the end
(NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests). There is no reference
counterpart: the reference runs one file at a time on one device (reference infer.py:44-54).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of the items rank ``rank`` owns; sizes differ by at most one, lower ranks first."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(n_items), world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_codes(local_codes: torch.Tensor, n_items: int, group=None, bins: int | None = None) -> torch.Tensor:
    """All-gather per-rank codes [1, B_r, L] (int64) into [1, n_items, L] on every rank.

    ``bins`` is the codebook size the codes index (``cfg.vq_bins``). Up to 32768 bins the codes fit a signed 16-bit
    word and travel as int16, widened afterwards: the payload of BASELINE.json's config 3 (1024 x 225 codes per rank)
    shrinks from 1.84 MB to 0.46 MB. Larger codebooks (the reference's EncodecFeatures default is 16384, but the YAML
    is free) and callers that do not say travel as int32, so an id can never wrap.
    Ragged shards (n_items % world != 0) are padded to the largest shard on the wire.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_codes
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    K, b_local, L = local_codes.shape
    s, e = shard_range(n_items, rank, world)
    if b_local != e - s:
        raise ValueError(f"rank {rank} holds {b_local} clips, expected {e - s}")
    wire_dtype = torch.int16 if bins is not None and 0 < int(bins) <= 32768 else torch.int32
    width = torch.empty((), dtype=wire_dtype).element_size()
    b_max = -(-n_items // world)
    wire = torch.zeros(K, b_max, L, dtype=wire_dtype, device=local_codes.device)
    wire[:, :b_local] = local_codes.to(wire_dtype)
    # neither NCCL nor gloo has a 16-bit integer type: ship the payload as raw bytes
    flat = torch.empty(world * K, b_max, width * L, dtype=torch.uint8, device=local_codes.device)
    dist.all_gather_into_tensor(flat, wire.view(torch.uint8), group=group)
    out = flat.view(wire_dtype).view(world, K, b_max, L)
    parts = []
    for r in range(world):
        rs, re = shard_range(n_items, r, world)
        parts.append(out[r, :, : re - rs])
    return torch.cat(parts, dim=1).to(torch.int64)
