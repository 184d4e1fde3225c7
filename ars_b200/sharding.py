"""Multi-GPU sharding of the render path: one process per GPU, torch.distributed for the plumbing.

Two natural shardings exist and neither needs a collective on the data path:
  * a batch of independent clips  -> static split of clip indices over ranks, balanced by output length;
  * one long mask-free render     -> contiguous ranges of overlap-save blocks per rank, each with an (L - 1)-frame
                                     input halo (only valid when no exact-N spectral mask is active).
Collectives are used only for the small things: gathering per-clip metrics, max-reducing the peak-guard scalars
of a block-sharded render, and timing barriers.
"""
from __future__ import annotations

import os

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
Y_HALO = 3 * 8192        # frames of the previous rank's stage output a rank needs: the loudness filters' warm-up (>= any layout delay)


class LongRenderRank:
    """Device work of ONE rank of a block-sharded long render, phase by phase, so that the exchanges between the phases
    can be real collectives (torch.distributed over NCCL, `render_long_sharded`) or emulated (the tests run several
    ranks one after another on one GPU).

        convolve()  -> MAX words [0:4]; every rank's last Y_HALO stage-output frames go to its successor (set_halo)
        pan_max()   -> MAX word [4]        (map_max() -> MAX word [5], Stereo layout only)
        final()     -> PCM of the rank's own frames; MAX words [8:10], SUM of the float64 at byte 48
        loudness_hops() -> SUM of the hop-energy vectors; loudness_gate() on whoever wants the number

    Only the rank's own slice of the clip (plus the convolution's input halo) is uploaded; the IR is a device tensor.
    """

    def __init__(self, samples, rate, d_ir, settings, rank, world, *, want_float=False, x_is_slice_from=None):
        import ctypes as C
        import torch
        from . import _capi, raytracer_studio as rs
        self.torch, self.capi, self.rs, self.C = torch, _capi, rs, C
        self.lib = _capi.init()
        self.rank, self.world = rank, world
        assert d_ir.is_cuda and d_ir.dtype == torch.float32 and d_ir.dim() == 2 and d_ir.shape[1] == 2, \
            "the long-render path takes a stereo (L, 2) float32 IR on the device"
        self.d_ir = d_ir.contiguous()
        self.L = int(d_ir.shape[0])
        self.p, _ = rs.make_render_params(rate, external_ir=True, want_lufs=True, **settings)
        layout = settings.get("target_channel_layout", rs.DEFAULT_CHANNEL_LAYOUT)
        self.C_out = rs.CHANNEL_LAYOUTS[layout]["channels"]
        self.stereo_layout = layout == "Stereo"
        # `samples`: the whole clip (n, cin), or -- x_is_slice_from=(first frame, total frames) -- just a slice of it
        x = rs._as_frames(samples) if isinstance(samples, np.ndarray) else samples
        if x_is_slice_from is None:
            self.n, self.cin = int(x.shape[0]), int(x.shape[1])
            x_first = 0
        else:
            x_first, self.n = int(x_is_slice_from[0]), int(x_is_slice_from[1])
            self.cin = int(x.shape[1])
        plan = _capi.ArsLongPlan()
        _capi.check(self.lib.ars_long_plan(self.p, self.n, self.L, plan), "ars_long_plan")
        self.plan = plan
        self.N, self.B = int(plan.frames_out), int(plan.block_frames)
        self.lo, self.hi = block_ranges(int(plan.n_blocks), world)[rank]
        self.f_lo, self.f_hi = min(self.N, self.lo * self.B), min(self.N, self.hi * self.B)
        self.x_lo = max(0, self.f_lo - int(plan.halo_frames))
        self.x_hi = max(self.x_lo, min(self.n, self.f_hi))
        self.y0 = max(0, self.f_lo - Y_HALO)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        if isinstance(x, np.ndarray):
            xs = x[self.x_lo - x_first:self.x_hi - x_first]
            assert xs.shape[0] == self.x_hi - self.x_lo, "the slice handed in does not cover this rank's input range"
            if xs.shape[0] == 0:
                xs = np.zeros((1, self.cin), np.float32)
            self.d_x = torch.from_numpy(np.ascontiguousarray(xs)).to(dev)
        else:                                             # a device (or pinned host) tensor holding exactly [x_lo, x_hi)
            self.d_x = x.to(dev, non_blocking=True)
        nf = self.frames()
        self.d_y = torch.zeros((max(1, self.f_hi - self.y0), 2), dtype=torch.float32, device=dev)
        self.state = torch.zeros(int(self.lib.ars_state_bytes()), dtype=torch.uint8, device=dev)
        self.d_pcm = torch.empty((max(1, nf), self.C_out), dtype=torch.int16, device=dev)
        self.d_f32 = torch.empty((max(1, nf), self.C_out), dtype=torch.float32, device=dev) if want_float else None
        self.n_hops = int(plan.hop_count)
        self.d_hops = torch.zeros(max(1, self.n_hops), dtype=torch.float64, device=dev)

    # ---- views of the state block for the collectives
    def words(self):
        return self.state.view(self.torch.int32)

    def sumsq(self):
        return self.state.view(self.torch.float64)[6:7]

    def frames(self):
        return max(0, self.f_hi - self.f_lo)

    def y_tail(self):
        """The last Y_HALO frames of this rank's stage output (zero-padded in front when the rank holds fewer)."""
        t = self.torch.zeros((Y_HALO, 2), dtype=self.torch.float32, device=self.dev)
        k = min(Y_HALO, self.frames())
        if k:
            t[Y_HALO - k:] = self.d_y[self.f_hi - self.y0 - k:self.f_hi - self.y0]
        return t

    def set_halo(self, prev_tail):
        """Frames [y0, f_lo) of the stage output = the end of the predecessor's tail."""
        k = self.f_lo - self.y0
        if k > 0:
            self.d_y[:k] = prev_tail[Y_HALO - k:]

    # ---- phases (everything is enqueued on the library's stream; nothing here waits for the GPU)
    def convolve(self):
        if self.hi <= self.lo:
            return                                        # more ranks than blocks: the zeroed state is the identity
        self.capi.check(self.lib.ars_long_convolve_dev(
            self.p, self.d_x.data_ptr(), self.x_lo, self.x_hi - self.x_lo, self.n, self.cin, self.d_ir.data_ptr(), self.L,
            None, 0, self.lo, self.hi, self.d_y[self.f_lo - self.y0:].data_ptr(), self.f_lo, self.state.data_ptr()),
            "ars_long_convolve_dev")

    def _tail(self, phase):
        if self.f_hi <= self.f_lo:
            return
        self.capi.check(self.lib.ars_long_tail_dev(
            self.p, phase, self.d_y.data_ptr(), self.y0, self.f_lo, self.f_hi, self.N, self.state.data_ptr(),
            self.d_f32.data_ptr() if (phase == 2 and self.d_f32 is not None) else None,
            self.d_pcm.data_ptr() if phase == 2 else None, None), "ars_long_tail_dev")

    def pan_max(self):
        self._tail(0)

    def map_max(self):
        if self.stereo_layout:
            self._tail(1)

    def final(self):
        self._tail(2)

    def loudness_hops(self):
        if self.n_hops == 0:
            return
        self.capi.check(self.lib.ars_long_loudness_hops_dev(
            self.p, self.d_y.data_ptr(), self.y0, self.f_lo, self.f_hi, self.N, self.state.data_ptr(),
            self.d_hops.data_ptr(), self.n_hops), "ars_long_loudness_hops_dev")

    def loudness_gate(self):
        status = self.C.c_int32(self.capi.LUFS_NONE)
        if self.n_hops:
            self.capi.check(self.lib.ars_long_loudness_gate_dev(self.p, self.d_hops.data_ptr(), self.n_hops, self.N,
                                                                self.state.data_ptr(), self.C.byref(status)),
                            "ars_long_loudness_gate_dev")
        return status.value

    def metrics(self, lufs_status):
        m = self.capi.ArsMetrics()
        self.capi.check(self.lib.ars_state_metrics(self.state.data_ptr(), int(self.N * self.C_out), lufs_status, m),
                        "ars_state_metrics")
        return self.rs._metrics_dict(m)


def lib_stream():
    """The library's CUDA stream as a torch stream: collectives issued under it are ordered with the library's kernels
    on the device, so no phase needs a host-side synchronize."""
    import torch
    from . import _capi
    ptr = _capi.init().ars_stream()
    return torch.cuda.ExternalStream(int(ptr))


_SIDE_STREAMS = {}


def side_stream(dev, index=0):
    """Extra streams per device for the PCM gather (kept: torch's allocator caches blocks per stream)."""
    import torch
    key = (int(dev.index if dev.index is not None else torch.cuda.current_device()), int(index))
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


PEER_PUSH_PARTS = 1       # a segment may travel as several concurrent copies (ARS_PEER_PUSH_PARTS); measured at N = 2: 1, 2, 4, 8 parts all 5.22-5.24 ms -- the link is the limit, not the copy engine


class _DeviceBlock:
    """A raw device allocation as something torch can wrap without a copy (torch.as_tensor reads this protocol)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


class PeerArray:
    """The whole-render PCM array of a block-sharded render: allocated on rank 0 (ars_peer_alloc), mapped into every other
    rank of the node through its CUDA IPC handle (ars_peer_open) -- the ranks then PUSH their segments into it over NVLink
    (ars_peer_push), instead of NCCL sends that rank 0 has to receive.  Collective: every rank of `group` must call it."""

    def __init__(self, nbytes, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _capi
        self.lib, self.capi = _capi.init(), _capi
        self.nbytes, self.group = int(nbytes), group
        self.rank = dist.get_rank(group)
        self.owner = self.rank == 0
        dev = torch.device("cuda", torch.cuda.current_device())
        handle = C.create_string_buffer(64)
        ptr = C.c_void_p()
        if self.owner:
            _capi.check(self.lib.ars_peer_alloc(self.nbytes, C.byref(ptr), handle), "ars_peer_alloc")
        h = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(dev)
        dist.broadcast(h, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if not self.owner:
            raw = bytes(h.cpu().numpy().tobytes())
            _capi.check(self.lib.ars_peer_open(C.create_string_buffer(raw, 64), C.byref(ptr)), "ars_peer_open")
        self.ptr = int(ptr.value)

    def push(self, byte_offset, d_src_ptr, nbytes, stream):
        """This rank's bytes into place, enqueued on `stream` (a torch stream)."""
        self.capi.check(self.lib.ars_peer_push(self.ptr + int(byte_offset), int(d_src_ptr), int(nbytes), int(stream.cuda_stream)),
                        "ars_peer_push")

    def tensor(self, shape, dtype_str="<i2"):
        """(owner) the array as a torch tensor over the same memory."""
        import torch
        return torch.as_tensor(_DeviceBlock(self.ptr, shape, dtype_str), device=torch.device("cuda", torch.cuda.current_device()))

    def release(self):
        """Collective: the other ranks unmap, then the owner frees."""
        import torch
        import torch.distributed as dist
        if self.ptr:
            torch.cuda.synchronize()
            if not self.owner:
                self.capi.check(self.lib.ars_peer_close(self.ptr), "ars_peer_close")
            dist.barrier(group=self.group)
            if self.owner:
                self.capi.check(self.lib.ars_peer_free(self.ptr), "ars_peer_free")
            self.ptr = 0


_PEER_ARRAYS = {}


def peer_array(nbytes, group=None):
    """One PeerArray per size, kept like the library's workspaces: the next sharded render of the same size on this
    process group reuses it (so copy the previous result first if it is still needed).  Collective."""
    key = (int(nbytes), id(group))
    if key not in _PEER_ARRAYS:
        _PEER_ARRAYS[key] = PeerArray(nbytes, group)
    return _PEER_ARRAYS[key]


def release_peer_arrays():
    """Collective: unmap / free every cached whole-render array (call before the process group is destroyed)."""
    for k in sorted(_PEER_ARRAYS, key=lambda kv: kv[0]):
        _PEER_ARRAYS[k].release()
    _PEER_ARRAYS.clear()


def render_long_sharded(samples, rate, external_ir_data, *, group=None, gather=True, to_host=True, x_is_slice_from=None,
                        timings=None, **settings):
    """One long mask-free render over all ranks of the process group (one process per GPU, NCCL).

    Every rank uploads only its own slice of the clip; the stereo IR is broadcast from rank 0 and stays on the device;
    three small all-reduces carry the peak-guard words (4 after the convolution, 1 after the pan maximum, 1 more for the
    Stereo layout), one all-gather passes every rank's last stage-output frames to its successor (layout delay + the
    loudness filters' warm-up), the hop energies of the loudness meter are summed, and the PCM segments travel to rank 0
    over NVLink on a second stream while the meter runs -- pushed by their ranks straight into their place in rank 0's
    whole-render array through a CUDA IPC peer mapping (gather=True, PeerArray), or as NCCL point-to-point sends
    (gather="nccl", the round-2 mid-state form).  Everything is ordered on the device: the host waits once, at the end.
    -> dict(metrics, names, rank_pcm (this rank's frames, device), frames, pcm_device (rank 0, gather=True: the whole
       render on the device), pcm (rank 0, gather and to_host: numpy))"""
    import torch
    import torch.distributed as dist
    from . import raytracer_studio as rs
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    stream = lib_stream()
    ev = []

    def mark(name):
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            ev.append((name, e))

    with torch.cuda.stream(stream):
        if isinstance(external_ir_data, torch.Tensor):
            d_ir = external_ir_data.to(dev)
        else:
            d_ir = torch.from_numpy(np.ascontiguousarray(external_ir_data, dtype=np.float32)).to(dev)
        mark("start")
        if world > 1:
            dist.broadcast(d_ir, src=0, group=group)                 # IR broadcast over NVLink; it stays on the device
        r = LongRenderRank(samples, rate, d_ir, settings, rank, world, x_is_slice_from=x_is_slice_from)
        MAX, SUM = (dist.ReduceOp.MAX, dist.ReduceOp.SUM) if world > 1 else (None, None)
        mark("upload + broadcast")
        r.convolve()
        mark("convolve")
        if world > 1:
            dist.all_reduce(r.words()[0:4], op=MAX, group=group)
            tails = torch.empty((world, Y_HALO, 2), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(tails, r.y_tail(), group=group)
            if rank > 0:
                r.set_halo(tails[rank - 1])
        mark("reduce maxima + halo exchange")
        r.pan_max()
        if world > 1:
            dist.all_reduce(r.words()[4:5], op=MAX, group=group)
        if r.stereo_layout:
            r.map_max()
            if world > 1:
                dist.all_reduce(r.words()[5:6], op=MAX, group=group)
        mark("pan / map maxima")
        r.final()
        mark("final pass")
        # the meter's and the metrics' small reductions first (NCCL runs a communicator's operations in issue order: behind
        # the gigabytes of the gather they would wait for it), then the PCM segments on the second stream
        r.loudness_hops()
        if world > 1:
            dist.all_reduce(r.d_hops, op=SUM, group=group)
            dist.all_reduce(r.words()[8:10], op=MAX, group=group)
            dist.all_reduce(r.sumsq(), op=SUM, group=group)
        pcm_all = None
        use_peer = world > 1 and bool(gather) and gather != "nccl"
        if use_peer:
            # every rank pushes its segment into rank 0's array over the peer mapping (a plain device-to-device copy over
            # NVLink on the second stream, next to the meter); the small all-reduce behind the pushes tells rank 0 that
            # all of them have landed
            done_final = torch.cuda.Event()
            done_final.record(stream)
            row = r.C_out * 2
            whole = peer_array(r.N * row, group)
            parts = max(1, int(os.environ.get("ARS_PEER_PUSH_PARTS", PEER_PUSH_PARTS)))
            total = r.frames() * row
            chunk = (-(-total // parts) + 4095) // 4096 * 4096            # (ceil(total / parts), rounded up to 4 KB)
            done_gather = []
            for k in range(parts):
                off = k * chunk
                if off >= total:
                    break
                side = side_stream(dev, k)
                side.wait_event(done_final)
                whole.push(r.f_lo * row + off, r.d_pcm.data_ptr() + off, min(chunk, total - off), side)
                e = torch.cuda.Event()
                e.record(side)
                done_gather.append(e)
            if rank == 0:
                pcm_all = whole.tensor((r.N, r.C_out))
        elif world > 1 and gather:
            side = side_stream(dev)
            done_final = torch.cuda.Event()
            done_final.record(stream)
            ranges = [(min(r.N, lo * r.B), min(r.N, hi * r.B)) for lo, hi in block_ranges(int(r.plan.n_blocks), world)]
            with torch.cuda.stream(side):
                side.wait_event(done_final)
                ops = []
                if rank == 0:
                    pcm_all = torch.empty((r.N, r.C_out), dtype=torch.int16, device=dev)
                    pcm_all[r.f_lo:r.f_hi] = r.d_pcm[:r.frames()]
                    for src, (a, b) in enumerate(ranges):
                        if src != 0 and b > a:             # NCCL has no int16: the frames travel as bytes
                            ops.append(dist.P2POp(dist.irecv, pcm_all[a:b].view(torch.uint8), src, group=group))
                elif r.frames() > 0:
                    ops.append(dist.P2POp(dist.isend, r.d_pcm[:r.frames()].view(torch.uint8), 0, group=group))
                if ops:
                    for w_ in dist.batch_isend_irecv(ops):
                        w_.wait()
                done_gather = torch.cuda.Event()
                done_gather.record(side)
        status = r.loudness_gate()
        mark("loudness + metric reductions")
        if world > 1 and gather:
            for e in (done_gather if isinstance(done_gather, list) else [done_gather]):
                stream.wait_event(e)
        if use_peer:
            landed = torch.zeros(1, dtype=torch.int32, device=dev)
            dist.all_reduce(landed, op=SUM, group=group)             # (behind every rank's push: rank 0 may read the array)
        mark("pcm gather (part not hidden by the meter)")
        metrics = r.metrics(status)                                  # (the one host wait of the render)
    if timings is not None:
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(ev[:-1], ev[1:]):
            timings[name] = timings.get(name, 0.0) + e0.elapsed_time(e1)
    layout = settings.get("target_channel_layout", rs.DEFAULT_CHANNEL_LAYOUT)
    names = rs.CHANNEL_LAYOUTS[layout]["names"]
    out = {"metrics": metrics, "names": names, "rank_pcm": r.d_pcm[:r.frames()], "frames": (r.f_lo, r.f_hi), "rank": r,
           "pcm_device": None, "pcm": None}
    if gather and rank == 0:
        out["pcm_device"] = r.d_pcm[:r.frames()] if world == 1 else pcm_all
        if to_host:
            out["pcm"] = out["pcm_device"].cpu().numpy()
    return out


def rank_ranges(plan, n_in, rank, world):
    """-> (first output frame, end output frame, first input frame, end input frame) of `rank` under `plan`."""
    lo, hi = block_ranges(int(plan.n_blocks), world)[rank]
    N, B = int(plan.frames_out), int(plan.block_frames)
    f_lo, f_hi = min(N, lo * B), min(N, hi * B)
    x_lo = max(0, f_lo - int(plan.halo_frames))
    return f_lo, f_hi, x_lo, max(x_lo, min(int(n_in), f_hi))


def long_clip_slice(lo, hi, cin=2, amp=0.1, seed=5):
    """Frames [lo, hi) of the synthetic long clip of the benchmark: generated in 2^20-frame chunks seeded by the chunk
    index, so that a rank can make its own slice without generating the whole hour."""
    CHK = 1 << 20
    out = np.empty((max(0, hi - lo), cin), np.float32)
    for c in range(lo // CHK, (max(lo, hi - 1)) // CHK + 1):
        a, b = max(lo, c * CHK), min(hi, (c + 1) * CHK)
        if b <= a:
            continue
        blk = np.random.default_rng([seed, c]).standard_normal((CHK, cin), dtype=np.float32) * np.float32(amp)
        out[a - lo:b - lo] = blk[a - c * CHK:b - c * CHK]
    return out


def bench_long(args, *, make_ir, load_peaks, ClockSampler, **_):
    """bench.py --workload long [--gpus N]: BASELINE configs[4], ONE long 48 kHz stereo render (x) a dense stereo IR
    (default 1 h (x) 20 s), EQ flat, 5.1 out, split by overlap-save block ranges over the N ranks.  Strong scaling:
    the work is fixed, `value` = clip seconds / max-over-ranks device time of the whole render including the
    collectives and the gather of the PCM segments on rank 0."""
    import json
    import os
    import time
    import torch
    import torch.distributed as dist
    from . import _capi, raytracer_studio as rs
    rate = 48000
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _capi.init(local)
    for kv in args.opt:
        _capi.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    seconds = args.seconds or 3600.0
    ir_seconds = args.ir_seconds or 20.0
    settings = dict(dry_wet=.5, dry_wet_kill_start=.5, bass_gain=1.0, treble_gain=1.0, x_pos=.5, y_pos=.5, z_pos=.5,
                    target_channel_layout="5.1 (Standard)")
    n = int(seconds * rate)
    ir = make_ir(ir_seconds)
    L = ir.shape[0]
    p, _ = rs.make_render_params(rate, external_ir=True, want_lufs=True, **settings)
    plan = _capi.ArsLongPlan()
    _capi.check(lib.ars_long_plan(p, n, L, plan), "ars_long_plan")
    f_lo, f_hi, x_lo, x_hi = rank_ranges(plan, n, rank, world)
    h_x = torch.from_numpy(long_clip_slice(x_lo, x_hi)).pin_memory()
    d_x = h_x.cuda()
    d_ir = torch.from_numpy(ir).cuda()
    h_pcm = torch.empty((max(1, f_hi - f_lo), 6), dtype=torch.int16).pin_memory()
    stream = lib_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_dev(timings=None, gather=True):
        return render_long_sharded(d_x, rate, d_ir, x_is_slice_from=(x_lo, n), gather=gather, to_host=False, timings=timings,
                                   **settings)

    def step_host():
        with torch.cuda.stream(stream):
            dx = h_x.to("cuda", non_blocking=True)
        out = render_long_sharded(dx, rate, d_ir, x_is_slice_from=(x_lo, n), gather=False, to_host=False, **settings)
        with torch.cuda.stream(stream):
            h_pcm[:out["rank_pcm"].shape[0]].copy_(out["rank_pcm"], non_blocking=True)
        stream.synchronize()
        return out

    for _ in range(max(1, args.warmup)):
        out = step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = int(lib.ars_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        out = step_dev()
    e1.record(stream)
    torch.cuda.synchronize()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = int(lib.ars_launch_count()) - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    timings = {}
    for _ in range(args.steps):
        step_dev(timings=timings)
    barrier()
    # the same render with the segments gathered by NCCL point-to-point sends instead of peer pushes (comparison)
    nccl_ms = None
    if world > 1:
        step_dev(gather="nccl")
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_dev(gather="nccl")
        e1.record(stream)
        torch.cuda.synchronize()
        nccl_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        barrier()
    # end to end: every rank uploads its slice from pinned host memory and downloads its own PCM segment
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_ms = max_over_ranks(1000 * (time.perf_counter() - t0))
    metrics = out["metrics"]
    if rank == 0:
        peaks = load_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        N = int(plan.frames_out)
        phases = {k: v / args.steps for k, v in timings.items()}
        algo = (8 + 2 * 6) * N
        step_ms = dev_ms / args.steps
        pcm_bytes_to_rank0 = 12 * (N - (f_hi - f_lo)) if world > 1 else 0
        line = {"metric": "audio-seconds rendered per second (x realtime) @48kHz, one long render (x) dense stereo IR",
                "value": seconds * args.steps / (dev_ms * 1e-3), "unit": "audio-seconds/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[4]: single long 48 kHz stereo render (x) dense external stereo IR, EQ flat, "
                                       "dw 0.5, 5.1 pan/map, metrics, int16 PCM; overlap-save blocks split over the ranks",
                           "clip_seconds": seconds, "frames_in": n, "ir_frames": L, "frames_out": N, "channels_out": 6,
                           "block_frames": int(plan.block_frames), "blocks": int(plan.n_blocks),
                           "route": "big-block overlap-save" if plan.route == 1 else "partitioned overlap-save",
                           "parallelism": f"block-range sharded x{world} (rank r: blocks {block_ranges(int(plan.n_blocks), world)})"},
                "phases_ms_rank0": phases,
                "ms_per_step_without_pcm_gather": sum(v for k, v in phases.items() if not k.startswith("pcm gather")),
                "value_without_pcm_gather": seconds / (1e-3 * sum(v for k, v in phases.items() if not k.startswith("pcm gather"))),
                "collectives": {"ir_broadcast_bytes": L * 8, "maxima_allreduce_bytes": [16, 4], "halo_allgather_bytes": world * Y_HALO * 8,
                                "hop_energy_allreduce_bytes": int(plan.hop_count) * 8, "metric_allreduce_bytes": [8, 8],
                                "pcm_gather_bytes_into_rank0": pcm_bytes_to_rank0,
                                "pcm_gather": ("peer push: every rank copies its segment into rank 0's array through a CUDA IPC "
                                               "mapping over NVLink (ars_peer_push), one 4-byte all-reduce behind the pushes")
                                if world > 1 else None,
                                "ms_per_step_with_nccl_p2p_gather": nccl_ms,
                                "limiting": "PCM gather into rank 0 (one GPU's NVLink ingress)" if world > 1 else None},
                "e2e": {"value": seconds * args.steps / (e2e_ms * 1e-3), "unit": "audio-seconds/s",
                        "ms_per_step": e2e_ms / args.steps,
                        "call": "sharding.render_long_sharded: every rank uploads its own slice from pinned host memory and "
                                "downloads its own PCM segment",
                        "h2d_bytes_per_step": int(h_x.numel() * 4 * (world if world > 1 else 1)),
                        "d2h_bytes_per_step": int(N * 12)},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "whole render (every kernel and collective of the step)",
                             "achieved": algo / (step_ms * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                             "frac": algo / (step_ms * 1e-3) / 1e9 / (peak * world), "algorithmic_bytes_per_step": algo,
                             "traffic": None},
                "metrics_of_last_render": metrics}
        print(json.dumps(line))
    if world > 1:
        release_peer_arrays()
        dist.destroy_process_group()
