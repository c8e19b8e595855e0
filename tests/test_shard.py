"""Clip sharding and the codes all-gather (world_size 2, gloo, CPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wavtokenizer_b200.shard import gather_codes, shard_range


@pytest.mark.parametrize("n,world", [(8192, 8), (256, 2), (7, 4), (3, 8), (0, 2), (1000, 3)])
def test_shard_range_partitions(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0
    sizes = [e - s for s, e in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _worker(rank, world, port, n_items, L, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        ok = True
        # 4096 bins: int16 on the wire; 65536 bins (ids above 32767 would wrap in int16) and "bins unknown": int32
        for bins, arg in ((4096, 4096), (65536, 65536), (65536, None)):
            full = torch.randint(0, bins, (1, n_items, L), generator=g)
            full[0, 0, 0] = bins - 1
            s, e = shard_range(n_items, rank, world)
            out = gather_codes(full[:, s:e].clone(), n_items, bins=arg)
            ok = ok and bool(torch.equal(out, full)) and out.dtype == torch.int64
        q.put((rank, ok, True))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [6, 7])
def test_gather_codes_world2_gloo(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_items) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)


def test_single_process_is_identity():
    c = torch.zeros(1, 3, 5, dtype=torch.int64)
    assert gather_codes(c, 3) is c
