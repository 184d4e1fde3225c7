"""Multi-GPU sharding of the render path: one process per GPU, torch.distributed for the plumbing.

Two natural shardings exist and neither needs a collective on the data path:
  * a batch of independent clips  -> static split of clip indices over ranks, balanced by output length;
  * one long mask-free render     -> contiguous ranges of overlap-save blocks per rank, each with an (L - 1)-frame
                                     input halo (only valid when no exact-N spectral mask is active).
Collectives are used only for the small things: gathering per-clip metrics, max-reducing the peak-guard scalars
of a block-sharded render, and timing barriers.
"""
from __future__ import annotations

import numpy as np


def partition_clips(lengths, world: int):
    """Longest-processing-time greedy split of clip indices over `world` ranks.  -> list of index lists
    (each sorted ascending, so a rank renders its clips in the caller's order)."""
    loads = [0] * world
    parts = [[] for _ in range(world)]
    for idx in sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i)):
        r = min(range(world), key=lambda j: (loads[j], j))
        parts[r].append(idx)
        loads[r] += int(lengths[idx])
    return [sorted(p) for p in parts]


def block_ranges(n_blocks: int, world: int):
    """Contiguous [lo, hi) ranges of overlap-save output blocks per rank (sizes differ by at most one)."""
    base, extra = divmod(int(n_blocks), world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def frame_range_for_blocks(lo: int, hi: int, block: int, ir_len: int, n_in: int):
    """Input frames rank needs for output blocks [lo, hi): output frames [lo*B, hi*B) depend on input frames
    [lo*B - (L - 1), hi*B).  -> (first_input_frame, last_input_frame_exclusive), clipped to the clip."""
    first = max(0, lo * block - (ir_len - 1))
    last = min(n_in, hi * block)
    return first, max(first, last)


def gather_metrics(local: dict, group=None):
    """All-gather {clip index: (lufs or None, true_peak_dbfs, rms_dbfs)} from every rank into one dict."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    boxes = [None] * dist.get_world_size(group)
    dist.all_gather_object(boxes, local, group=group)
    out = {}
    for b in boxes:
        out.update(b)
    return out


def reduce_peak(local_peak: float, group=None, device=None):
    """Max over ranks of a peak-guard scalar (the three np.max(np.abs(.)) normalisers of a block-sharded render)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(local_peak)
    t = torch.tensor([float(local_peak)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def render_batch_sharded(jobs, *, rank=None, world=None, group=None, **kw):
    """Render this rank's share of `jobs` (same list on every rank) with the pipelined batch call and gather the
    metrics of all clips.  -> (results for the local clips as {index: result}, metrics of every clip)."""
    import torch.distributed as dist
    from . import raytracer_studio as rs
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    lengths = [int(np.shape(j["samples"])[0]) for j in jobs]
    mine = partition_clips(lengths, world)[rank]
    res = rs.render_batch([jobs[i] for i in mine], **kw)
    local = {i: r for i, r in zip(mine, res)}
    met = {i: (r["metrics"]["lufs"], r["metrics"]["true_peak_dbfs"], r["metrics"]["rms_dbfs"])
           for i, r in local.items() if r["metrics"] is not None}
    return local, gather_metrics(met, group)
