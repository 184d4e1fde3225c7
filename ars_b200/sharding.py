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


# ------------------------------------------------------------------------------------------------
# One long mask-free render split by overlap-save block ranges (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------
class LongRenderRank:
    """Device work of ONE rank of a block-sharded long render, phase by phase, so that the collectives between the
    phases can be real (torch.distributed over NCCL) or emulated (tests run several `ranks` on one GPU).

        convolve() -> reduce state words [0:4] (MAX) -> pan_max() -> reduce [4] (MAX) -> map_max() -> reduce [5] (MAX)
        -> final() -> reduce [8:10] (MAX) and the float64 at byte 48 (SUM) -> gather PCM / loudness-feed slices
    """

    def __init__(self, samples, rate, ir, settings, rank, world, *, want_float=False):
        import torch
        from . import _capi, raytracer_studio as rs
        self.torch, self.capi, self.rs = torch, _capi, rs
        self.lib = _capi.init()
        x = rs._as_frames(samples)
        self.n, self.cin = x.shape
        ir = np.ascontiguousarray(ir, dtype=np.float32)
        assert ir.ndim == 2 and ir.shape[1] == 2, "the long-render path takes a stereo (L, 2) IR"
        self.L = ir.shape[0]
        self.p, _ = rs.make_render_params(rate, external_ir=True, want_lufs=True, **settings)
        self.N = self.n + self.L - 1
        layout = settings.get("target_channel_layout", rs.DEFAULT_CHANNEL_LAYOUT)
        self.C = rs.CHANNEL_LAYOUTS[layout]["channels"]
        self.B = int(self.lib.ars_ols_block_frames())
        nblk = -(-self.N // self.B)
        P = -(-self.L // self.B)
        self.lo, self.hi = block_ranges(nblk, world)[rank]
        delay = {"7.1 (Surround)": int(int(rate) * 12 / 1000), "5.1.2 (Atmos Light)": int(int(rate) * 18 / 1000)}.get(layout, 0)
        halo = -(-delay // self.B) if delay else 0
        self.clo = max(0, self.lo - halo)                         # blocks computed here (incl. the tail's delay halo)
        seg0 = max(0, self.clo - (P - 1))
        self.x_lo = max(0, (seg0 - 1) * self.B)
        self.x_hi = max(self.x_lo, min(self.n, self.hi * self.B))
        self.f_lo, self.f_hi = self.lo * self.B, min(self.N, self.hi * self.B)
        self.y0 = self.clo * self.B
        dev = torch.device("cuda", torch.cuda.current_device())
        xs = x[self.x_lo:self.x_hi] if self.x_hi > self.x_lo else np.zeros((1, self.cin), np.float32)
        self.d_x = torch.from_numpy(np.ascontiguousarray(xs)).to(dev)
        self.d_ir = torch.from_numpy(ir).to(dev)
        self.d_y = torch.empty((max(1, self.hi * self.B - self.y0), 2), dtype=torch.float32, device=dev)
        self.state = torch.zeros(int(self.lib.ars_state_bytes()), dtype=torch.uint8, device=dev)
        nf = max(0, self.f_hi - self.f_lo)
        self.d_pcm = torch.empty((max(1, nf), self.C), dtype=torch.int16, device=dev)
        self.d_mono = torch.empty(max(1, nf), dtype=torch.float32, device=dev)
        self.d_f32 = torch.empty((max(1, nf), self.C), dtype=torch.float32, device=dev) if want_float else None
        torch.cuda.synchronize()

    # views of the state block for the collectives
    def words(self):
        return self.state.view(self.torch.int32)

    def sumsq(self):
        return self.state.view(self.torch.float64)[6:7]

    def _sync(self):
        self.capi.check(self.lib.ars_sync(), "ars_sync")

    def convolve(self):
        self.capi.check(self.lib.ars_long_convolve_dev(
            self.p, self.d_x.data_ptr(), self.x_lo, self.x_hi - self.x_lo, self.n, self.cin, self.d_ir.data_ptr(), self.L,
            None, 0, self.clo, self.hi, self.d_y.data_ptr(), self.y0, self.state.data_ptr()), "ars_long_convolve_dev")
        self._sync()

    def _tail(self, phase):
        if self.f_hi <= self.f_lo:
            return
        self.torch.cuda.synchronize()
        self.capi.check(self.lib.ars_long_tail_dev(
            self.p, phase, self.d_y.data_ptr(), self.y0, self.f_lo, self.f_hi, self.N, self.state.data_ptr(),
            self.d_f32.data_ptr() if (phase == 2 and self.d_f32 is not None) else None,
            self.d_pcm.data_ptr() if phase == 2 else None, self.d_mono.data_ptr() if phase == 2 else None),
            "ars_long_tail_dev")
        self._sync()

    def pan_max(self):
        self._tail(0)

    def map_max(self):
        self._tail(1)

    def final(self):
        self._tail(2)

    def frames(self):
        return max(0, self.f_hi - self.f_lo)


def finish_long_render(rank0: "LongRenderRank", d_mono_all, count):
    """Loudness of the gathered feed + metrics read-back on the gathering rank."""
    import ctypes as C
    lib, capi = rank0.lib, rank0.capi
    rank0.torch.cuda.synchronize()
    status = C.c_int32(0)
    capi.check(lib.ars_loudness_dev(d_mono_all.data_ptr(), int(d_mono_all.numel()), float(rank0.p.rate),
                                    rank0.state.data_ptr(), C.byref(status)), "ars_loudness_dev")
    m = capi.ArsMetrics()
    capi.check(lib.ars_state_metrics(rank0.state.data_ptr(), int(count), status.value, m), "ars_state_metrics")
    return rank0.rs._metrics_dict(m)


def render_long_sharded(samples, rate, external_ir_data, *, group=None, **settings):
    """One long mask-free render over all ranks of the process group (one process per GPU, NCCL).  The stereo IR is
    broadcast from rank 0, every rank convolves its range of overlap-save blocks, the peak-guard maxima are
    max-reduced, the PCM and loudness-feed segments are gathered on rank 0.
    -> on rank 0: dict(pcm, metrics, names); on the other ranks: None."""
    import torch
    import torch.distributed as dist
    from . import raytracer_studio as rs
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    ir = torch.from_numpy(np.ascontiguousarray(external_ir_data, dtype=np.float32)).to(dev)
    if world > 1:
        dist.broadcast(ir, src=0, group=group)              # IR broadcast over NVLink
    r = LongRenderRank(samples, rate, ir.cpu().numpy(), settings, rank, world)
    MAX, SUM = (dist.ReduceOp.MAX, dist.ReduceOp.SUM) if world > 1 else (None, None)

    def red(t, op):
        if world > 1:
            torch.cuda.synchronize()
            dist.all_reduce(t, op=op, group=group)
            torch.cuda.synchronize()

    r.convolve()
    red(r.words()[0:4], MAX)
    r.pan_max()
    red(r.words()[4:5], MAX)
    r.map_max()
    red(r.words()[5:6], MAX)
    r.final()
    red(r.words()[8:10], MAX)
    red(r.sumsq(), SUM)
    layout = settings.get("target_channel_layout", rs.DEFAULT_CHANNEL_LAYOUT)
    names = rs.CHANNEL_LAYOUTS[layout]["names"]
    if world == 1:
        metrics = finish_long_render(r, r.d_mono[:r.frames()], r.N * r.C)
        return {"pcm": r.d_pcm[:r.frames()].cpu().numpy(), "metrics": metrics, "names": names}
    # gather the variable-length segments, padded to the longest
    counts = [max(0, min(r.N, hi * r.B) - lo * r.B) for lo, hi in block_ranges(-(-r.N // r.B), world)]
    pad = max(counts)
    pcm_pad = torch.zeros((pad, r.C), dtype=torch.int16, device=dev)
    mono_pad = torch.zeros(pad, dtype=torch.float32, device=dev)
    pcm_pad[:r.frames()] = r.d_pcm[:r.frames()]
    mono_pad[:r.frames()] = r.d_mono[:r.frames()]
    pcm_bytes = pcm_pad.view(torch.uint8)                   # NCCL has no int16: ship the PCM frames as bytes
    byte_list = [torch.empty_like(pcm_bytes) for _ in range(world)] if rank == 0 else None
    mono_list = [torch.empty_like(mono_pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pcm_bytes, byte_list, dst=0, group=group)    # output segments gathered over NVLink
    dist.gather(mono_pad, mono_list, dst=0, group=group)
    if rank != 0:
        return None
    pcm = torch.cat([t.view(torch.int16)[:c] for t, c in zip(byte_list, counts)], dim=0)
    mono = torch.cat([t[:c] for t, c in zip(mono_list, counts)], dim=0).contiguous()
    metrics = finish_long_render(r, mono, r.N * r.C)
    return {"pcm": pcm.cpu().numpy(), "metrics": metrics, "names": names}
