"""Variable-length batches.

The reference has no batching story: ``infer.py:44-54`` and ``README.md:50-112`` push one file of
arbitrary length through ``encode_infer`` / ``decode`` at a time, so the result of clip i is by
definition what a batch-of-one call returns (reflect padding at ITS end, conv.py:195-211; GroupNorm /
attention over ITS frames, models.py:10-16,107-127; iSTFT "same" trimming of ITS length,
spectral_ops.py:58-73). Zero- or reflect-padding clips to a common length changes those results, so
ragged input is served by LENGTH BUCKETS instead: clips of equal length are stacked and go through the
C ABI as one batch (clips are independent end to end), buckets are visited longest first
so the workspace arena is sized once, and results are scattered back to input order. A workload of
files cut to a handful of lengths (or framed to a multiple of the hop) runs at batch speed; all-distinct
lengths degrade to the reference's own one-call-per-file pattern, never to different numbers.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch


def length_buckets(lengths: Sequence[int], max_bucket: int = 0) -> List[Tuple[int, List[int]]]:
    """[(length, [input indices in input order])], longest length first; buckets larger than
    ``max_bucket`` (> 0) are split into consecutive pieces of at most that many clips."""
    by_len: dict = {}
    for i, n in enumerate(lengths):
        n = int(n)
        if n < 1:
            raise ValueError(f"clip {i} is empty (length {n})")
        by_len.setdefault(n, []).append(i)
    out: List[Tuple[int, List[int]]] = []
    for n in sorted(by_len, reverse=True):
        idx = by_len[n]
        step = max_bucket if max_bucket > 0 else len(idx)
        for j in range(0, len(idx), step):
            out.append((n, idx[j:j + step]))
    return out


def assign_lanes(buckets: Sequence[Tuple[int, List[int]]], n_lanes: int) -> List[int]:
    """Lane (stream) of every bucket: greedy longest-processing-time on samples per bucket (buckets arrive longest
    first), so that the lanes finish together."""
    load = [0] * max(1, n_lanes)
    lanes = []
    for n, idx in buckets:
        k = min(range(len(load)), key=lambda i: (load[i], i))
        lanes.append(k)
        load[k] += n * len(idx)
    return lanes


def run_bucketed(items: Sequence[torch.Tensor], fn, max_bucket: int = 0, streams=None) -> List[Tuple[torch.Tensor, ...]]:
    """Stack ``items`` (tensors whose LAST dimension is the ragged one, equal leading shape) per length bucket, call
    ``fn(batch)`` -> tuple of ``(tensor, batch_dim)`` pairs, and return the per-item slices in input order.

    With ``streams`` (a list of CUDA streams) ``fn`` is a list of callables, one per stream: bucket j runs
    ``fn[lane](batch)`` on ``streams[lane]`` (``assign_lanes``), so that small buckets - which cannot fill the GPU on
    their own - overlap. Every lane first waits for the caller's stream and the caller's stream waits for every lane."""
    if len(items) == 0:
        return []
    lead = items[0].shape[:-1]
    for i, t in enumerate(items):
        if t.shape[:-1] != lead:
            raise ValueError(f"item {i} has shape {tuple(t.shape)}, expected {tuple(lead)} + [length]")
    results: List = [None] * len(items)
    buckets = length_buckets([t.shape[-1] for t in items], max_bucket)
    if not streams:
        for _, idx in buckets:
            outs = fn(torch.stack([items[i] for i in idx], dim=0))
            for k, i in enumerate(idx):
                results[i] = tuple(o.select(d, k) for o, d in outs)
        return results
    cur = torch.cuda.current_stream(items[0].device)
    batches = [torch.stack([items[i] for i in idx], dim=0) for _, idx in buckets]  # on the caller's stream
    for st in streams:
        st.wait_stream(cur)  # fork: the lanes start once the stacked inputs exist
    for (_, idx), lane, batch in zip(buckets, assign_lanes(buckets, len(streams)), batches):
        st = streams[lane]
        batch.record_stream(st)
        with torch.cuda.stream(st):
            outs = fn[lane](batch)
        for o, _ in outs:
            o.record_stream(cur)  # allocated on the lane, consumed on the caller's stream
        for k, i in enumerate(idx):
            results[i] = tuple(o.select(d, k) for o, d in outs)
    for st in streams:
        cur.wait_stream(st)  # join
    return results
