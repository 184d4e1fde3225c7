#!/usr/bin/env python
"""Benchmark of the render hot path (BASELINE.json metric: audio-seconds rendered per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg1]

One "step" = one complete render of the workload clip: procedural IR synthesis -> exact-N spectral
filter (convolution + air absorption + dry/wet + EQ) -> pan -> layout map -> metrics -> int16 PCM.

  value : clip-seconds rendered per second with the input clip already resident in HBM and the PCM
          result left in HBM (ars_render_dev), timed with CUDA events on the library's stream.
  e2e   : the same metric through the host-buffer C-ABI call (ars_render): pinned host input copied
          to the device, PCM frames + metrics copied back, every step, inside the timed region.
  N > 1 : one process per GPU (torchrun), every rank renders its own clip of the same shape
          (clips are independent: no collective on the data path), barrier + max over ranks.

`--impl reference` times the CPU oracle (oracle/ars_oracle.py: a numpy/scipy restatement of the
reference, which is a Python file and cannot travel) on the host cores, one clip per process.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE = 48000
WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (48 kHz, 8 s IR, 5.1 bed)
    "cfg3": dict(desc="configs[2]: 5 min 48 kHz 6-ch clip (ch0-1 used, as the reference does), internal hall "
                      "'Cathedral' / Stein, 8 s procedural IR, air 0.1, EQ flat, dw 0.5, 5.1 pan -> 5.1.2 map, "
                      "LUFS/peak/RMS metrics, int16 PCM out",
                 seconds=300, cin=6, amp=0.2, data_seed=2, np_seed=3,
                 settings=dict(hall_type="Cathedral", room_size=20000., ir_duration=8.0, diffusion=.5,
                               air_absorption=.1, base_early_level=.8, base_late_level=.6, dry_wet=.5,
                               dry_wet_kill_start=.5, bass_gain=1.0, treble_gain=1.0, x_pos=.5, y_pos=.5, z_pos=.5,
                               material="Stein", target_channel_layout="5.1.2 (Atmos Light)")),
    "cfg2": dict(desc="configs[1]: 60 s 48 kHz mono clip, internal hall 'Room' / Holz / 200 m^3, air 0.1, "
                      "bass 1.5 / treble 0.8, dw 0.6, pan (.3,.4,.6) -> 5.1, metrics, int16 PCM out",
                 seconds=60, cin=1, amp=0.3, data_seed=0, np_seed=11,
                 settings=dict(hall_type="Room", room_size=200., diffusion=.5, air_absorption=.1, base_early_level=.8,
                               base_late_level=.6, dry_wet=.6, dry_wet_kill_start=.5, bass_gain=1.5, treble_gain=.8,
                               x_pos=.3, y_pos=.4, z_pos=.6, material="Holz",
                               target_channel_layout="5.1 (Standard)")),
    # BASELINE.json configs[4] (single-GPU slice): long stereo render (x) dense stereo IR, EQ flat -> the
    # partitioned overlap-save path; default 600 s (x) 8 s, override with --seconds / --ir-seconds
    "cfg5": dict(desc="configs[4]: long 48 kHz stereo clip (x) dense external stereo IR, EQ flat, dw 0.5, 5.1 out "
                      "(overlap-save convolution path)",
                 seconds=600, cin=2, amp=0.1, data_seed=5, np_seed=0, ext_ir_seconds=8.0,
                 settings=dict(dry_wet=.5, dry_wet_kill_start=.5, bass_gain=1.0, treble_gain=1.0, x_pos=.5, y_pos=.5,
                               z_pos=.5, target_channel_layout="5.1 (Standard)")),
}
ORACLE_KW = {"hall_type": "hall", "room_size": "room_size", "diffusion": "diffusion", "air_absorption": "air",
             "base_early_level": "early", "base_late_level": "late", "dry_wet": "dry_wet_amount",
             "dry_wet_kill_start": "kill_start", "bass_gain": "bass", "treble_gain": "treble", "x_pos": "x",
             "y_pos": "y", "z_pos": "z", "material": "material", "target_channel_layout": "layout"}


def make_ir(seconds):
    """Dense synthetic stereo IR: Gaussian noise under an exponential decay (RT60 ~ 0.6 L), peak-normalised."""
    L = int(seconds * RATE)
    g = np.random.default_rng(6)
    ir = g.standard_normal((L, 2), dtype=np.float32) * np.exp(-6.9 * np.arange(L, dtype=np.float32) / (0.6 * L))[:, None]
    return (ir / np.max(np.abs(ir)) / np.float32(20.0)).astype(np.float32)


def make_clip(w, seconds, seed_offset=0):
    n = int(seconds * RATE)
    g = np.random.default_rng(w["data_seed"] + seed_offset)
    shape = (n, w["cin"]) if w["cin"] > 1 else (n,)
    return (w["amp"] * g.standard_normal(shape, dtype=np.float32)).astype(np.float32)


# ----------------------------------------------------------------------------- clocks ------
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(delete=False, suffix=".csv", prefix="clocks_")
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [s.strip() for s in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax.append(float(p[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.remove(self.path)
        except Exception:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank to the CPU cores next to its GPU (NVML's affinity mask), so that pinned host buffers are
    allocated on the GPU's own NUMA node; with several ranks per box the host side of the copies otherwise crosses
    sockets.  Best effort: silently does nothing when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


# ----------------------------------------------------------------------------- CPU arms -----
def _oracle_render(args):
    w_name, seconds, seed_offset = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ars_oracle as orc
    w = WORKLOADS[w_name]
    x = make_clip(w, seconds, seed_offset)
    kw = {ORACLE_KW[k]: v for k, v in w["settings"].items() if k in ORACLE_KW}
    np.random.seed(w["np_seed"])
    t0 = time.perf_counter()
    if "ir_duration" in w["settings"]:
        # same override the GPU arm uses: call the oracle stages with the forced IR duration
        s = w["settings"]
        dur, refl, mdel, split = orc.shape_params(s["hall_type"], s["room_size"], s["z_pos"])
        dur = s["ir_duration"]
        d = orc.directionality(s["x_pos"], s["y_pos"], s["z_pos"], s["hall_type"], s["diffusion"], s["dry_wet"])
        e_ir, l_ir = orc.generate_ir(RATE, dur, refl, mdel, s["material"], d, split, s["diffusion"])
        e_lvl, l_lvl = orc.adapt_levels(s["dry_wet"], s["base_early_level"], s["base_late_level"])
        xs = x[:, :2] if x.ndim == 2 and x.shape[1] > 2 else x
        stereo = orc.convolve_split(xs, e_ir, l_ir, e_lvl, l_lvl, s["dry_wet"], s["bass_gain"], s["treble_gain"], RATE,
                                    s["dry_wet_kill_start"], s["air_absorption"])
        six = orc.pan_5_1(stereo, s["x_pos"], s["y_pos"], s["z_pos"])
        final, _ = orc.map_layout(six, s["target_channel_layout"], RATE, s["z_pos"])
        orc.metrics(final, RATE)
        orc.pcm16(final)
    else:
        orc.render(x, RATE, **kw)
    return time.perf_counter() - t0


def cpu_baseline(w_name, sample_seconds):
    dt = _oracle_render((w_name, sample_seconds, 0))
    return {"value": sample_seconds / dt, "unit": "audio-seconds/s", "cores": 1, "kind": "port",
            "sample": f"first {sample_seconds} s of the {w_name} clip, same settings, oracle/ars_oracle.py "
                      f"(numpy/scipy restatement of the reference; single-threaded like the reference), {dt:.2f} s"}


def run_reference(args):
    """--impl reference: the CPU oracle on every host core (one clip per process), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample = args.ref_sample_seconds
    jobs = [(args.workload, sample, i) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup if args.warmup < 2 else 1):
            pool.map(_oracle_render, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_oracle_render, jobs)
        dt = time.perf_counter() - t0
    value = cores * sample * args.steps / dt
    line = {"impl": "reference", "metric": "audio-seconds rendered per second (x realtime) @48kHz, 8 s IR",
            "value": value, "unit": "audio-seconds/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "sample": f"{sample} s clips, {cores} at a time"},
            "cpu_baseline": {"value": value, "unit": "audio-seconds/s", "cores": cores, "kind": "port",
                             "sample": f"{cores} processes x {sample} s clips per step (clip-parallel; numpy/scipy FFTs "
                                       "are single-threaded), oracle/ars_oracle.py"},
            "e2e": {"value": value, "unit": "audio-seconds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm ------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from ars_b200 import _capi, raytracer_studio as rs
    from ars_b200._capi import ArsMetrics

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _capi.init(local)
    for kv in args.opt:
        _capi.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    w = WORKLOADS[args.workload]
    if args.cin:
        w = dict(w, cin=args.cin, desc=w["desc"] + " [input channels overridden: %d]" % args.cin)
    if args.air is not None:
        w = dict(w, settings=dict(w["settings"], air_absorption=args.air), desc=w["desc"] + " [air overridden: %g]" % args.air)
    seconds = args.seconds or w["seconds"]
    x = make_clip(w, seconds, rank)
    x2 = x if x.ndim == 2 else x[:, None]
    n, cin = x2.shape
    ext = "ext_ir_seconds" in w
    p, refl = rs.make_render_params(RATE, want_lufs=True, external_ir=ext, **w["settings"])
    L = 0
    h_ir = d_ir = None
    if ext:
        ir = make_ir(args.ir_seconds or w["ext_ir_seconds"])
        L = ir.shape[0]
        h_ir = torch.from_numpy(ir).pin_memory()
        d_ir = h_ir.cuda()
        taps, bases, noise = np.zeros(0, np.int64), np.zeros(0), np.zeros(0)
    else:
        np.random.seed(w["np_seed"] + rank)
        taps, bases, noise = rs.draw_ir_randoms(RATE, p.ir_duration, refl, p.ir_max_delay, p.ir_split_time)
    N = int(lib.ars_render_out_len(p, n, L))
    C = rs.CHANNEL_LAYOUTS[w["settings"]["target_channel_layout"]]["channels"]

    # pinned host buffers (e2e) and device-resident copies (value)
    h_in = torch.from_numpy(np.ascontiguousarray(x2)).pin_memory()
    h_noise = torch.from_numpy(noise).pin_memory()
    h_pcm = torch.empty((N, C), dtype=torch.int16).pin_memory()
    d_in = h_in.cuda()
    d_noise = h_noise.cuda()
    d_pcm = torch.empty((N, C), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    keep = []
    draws_h = _capi.make_draws(taps, bases, h_noise.numpy(), keep)
    draws_d = _capi.make_draws(taps, bases, int(d_noise.data_ptr()), keep)
    draws_d.noise_len = int(noise.size)
    m = ArsMetrics()

    d_ir_ptr = d_ir.data_ptr() if ext else None
    h_ir_ptr = h_ir.data_ptr() if ext else None

    def step_dev():
        _capi.check(lib.ars_render_dev(p, d_in.data_ptr(), n, cin, d_ir_ptr, L, None if ext else draws_d, None, None,
                                       d_pcm.data_ptr(), m), "ars_render_dev")

    def step_host():
        _capi.check(lib.ars_render(p, h_in.data_ptr(), n, cin, h_ir_ptr, L, None if ext else draws_h, None, None,
                                   h_pcm.data_ptr(), m), "ars_render")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _capi.check(lib.ars_sync(), "ars_sync")

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = int(lib.ars_launch_count())
    ms = _capi.C.c_float(0)
    _capi.check(lib.ars_timer_begin(), "timer")
    for _ in range(args.steps):
        step_dev()
    _capi.check(lib.ars_timer_end(_capi.C.byref(ms)), "timer")
    launches = int(lib.ars_launch_count()) - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(float(ms.value))
    metrics_dev = rs._metrics_dict(m)

    # ---- end-to-end timing (host buffers in, host buffers out) ----
    # the public batch call: `steps` clips from pinned host memory, PCM frames + metrics back to pinned host
    # memory; clip i+1's upload and clip i-1's download overlap clip i's compute inside the library
    h_pcm2 = torch.empty((N, C), dtype=torch.int16).pin_memory()
    ms_list = [ArsMetrics() for _ in range(args.steps)]

    def batch_host(count):
        clips = (_capi.ArsClip * count)()
        for i in range(count):
            k = clips[i]
            k.params = _capi.C.pointer(p)
            k.in_ = h_in.data_ptr()
            k.n, k.cin = n, cin
            if ext:
                k.ext_ir, k.ext_ir_len = h_ir_ptr, L
            else:
                k.draws = _capi.C.pointer(draws_h)
            k.out_pcm = (h_pcm if i % 2 == 0 else h_pcm2).data_ptr()
            k.metrics = _capi.C.pointer(ms_list[i])
        _capi.check(lib.ars_render_batch(clips, count), "ars_render_batch")

    for _ in range(max(1, min(args.warmup, 2))):
        step_host()
    batch_host(min(2, args.steps))
    barrier()
    t0 = time.perf_counter()
    batch_host(args.steps)
    e2e_ms = max_over_ranks(1000 * (time.perf_counter() - t0))
    barrier()
    t0 = time.perf_counter()
    step_host()
    single_ms = 1000 * (time.perf_counter() - t0)
    barrier()

    # ---- roofline of the dominant kernels (FFT passes), separate untimed run with per-launch events ----
    folds0 = int(lib.ars_air_fold_count())
    _capi.check(lib.ars_profile_begin(), "profile")
    step_dev()
    air_folds = int(lib.ars_air_fold_count()) > folds0
    pl, pms, pbytes = _capi.C.c_int64(0), _capi.C.c_double(0), _capi.C.c_double(0)
    _capi.check(lib.ars_profile_end(_capi.C.byref(pl), _capi.C.byref(pms), _capi.C.byref(pbytes)), "profile")
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = pbytes.value / (pms.value * 1e-3) / 1e9 if pms.value > 0 else 0.0
    traffic = traffic_src = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        traffic, traffic_src = float(t["dram_bytes_per_launch_mean"]), t["source"]
    except Exception:
        pass
    step_ms = dev_ms / args.steps
    total_seconds = seconds * world * args.steps
    out_bytes_per_frame = 8 + 2 * C
    line = {
        "metric": "audio-seconds rendered per second (x realtime) @48kHz, 8 s IR",
        "value": total_seconds / (dev_ms * 1e-3), "unit": "audio-seconds/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "clip_seconds": seconds, "frames_in": n, "frames_out": N,
                   "channels_out": C, "clips_per_step": world, "parallelism": f"clip-sharded x{world}",
                   "ir_frames": L if ext else int(p.ir_duration * RATE),
                   "route": "folded-air overlap-save" if air_folds else "see DESIGN.md section 2",
                   "l2": "working set (input %d MB, delay-line / FFT work buffers >= %d MB, output %d MB) exceeds the "
                         "126 MB L2; no flush needed" % (n * cin * 4 >> 20, 8 * N >> 19, N * C * 2 >> 20)},
        "e2e": {"value": total_seconds / (e2e_ms * 1e-3), "unit": "audio-seconds/s", "ms_per_step": e2e_ms / args.steps,
                "call": "ars_render_batch (host buffers, copy/compute pipelined across the steps' clips)",
                "single_call_ms": single_ms,
                "h2d_bytes_per_step": int(h_in.numel() * 4 + noise.size * 8 + taps.size * 16 + L * 8),
                "d2h_bytes_per_step": int(h_pcm.numel() * 2 + 56)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "fft pass kernels (pass_contig_kernel block transforms of the overlap-save "
                                               "route / pass_strided_kernel + pass_contig_kernel M-point passes)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "launches_per_step": int(pl.value), "ms_per_step_in_kernel": pms.value,
                     "share_of_step": pms.value / step_ms if step_ms else None,
                     "algorithmic_bytes_per_launch": pbytes.value / max(1, pl.value),
                     "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write, ncu)",
                     "traffic_source": traffic_src},
        "roofline_whole_render": {"algorithmic_bytes_per_step": int(out_bytes_per_frame * N),
                                  "achieved": out_bytes_per_frame * N / (step_ms * 1e-3) / 1e9, "unit": "GB/s",
                                  "frac": out_bytes_per_frame * N / (step_ms * 1e-3) / 1e9 / peak,
                                  "note": "SURVEY 8(d) compulsory bytes of an ideal single fused pass (8 B in + 2 B per "
                                          "output channel per frame) over the whole step, all kernels included"},
        "metrics_of_last_render": metrics_dev,
    }
    if world == 1 and not args.no_cpu and not ext:
        line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_sample_seconds)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cfg4_presets(count, first=0):
    """BASELINE configs[3] / SURVEY 8(d): random presets drawn from the UI ranges, one np.random seed per clip."""
    halls = ["Plate", "Room", "Cathedral"]
    mats = ["Stein", "Holz", "Teppich", "Glas", "Beton", "Vorhang (schwer)"]
    out = []
    for i in range(first, first + count):
        g = np.random.default_rng([4, i])
        u = g.uniform(0, 1, 7)
        out.append(dict(seed=1000 + i, hall_type=halls[int(g.integers(3))], material=mats[int(g.integers(6))],
                        room_size=float(10 * g.integers(1, 101)), diffusion=u[0], air_absorption=u[1], dry_wet=u[2],
                        dry_wet_kill_start=u[3], x_pos=u[4], y_pos=u[5], z_pos=u[6],
                        base_early_level=float(g.uniform(0, 2)), base_late_level=float(g.uniform(0, 2)),
                        bass_gain=1.0 if g.uniform() < .5 else float(g.uniform(.1, 5)),
                        treble_gain=1.0 if g.uniform() < .5 else float(g.uniform(.1, 5)),
                        target_channel_layout="7.1 (Surround)"))
    return out


def run_cfg4(args):
    """Batch of 30 s stereo clips with random presets (configs[3]); clips are split over the ranks, no collective on
    the data path.  Not the headline line: an extra measurement of the small-clip regime (FFT buffers fit the L2)."""
    import torch
    import torch.distributed as dist
    from ars_b200 import _capi, raytracer_studio as rs, sharding as sh
    from ars_b200._capi import ArsMetrics
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _capi.init(local)
    for kv in args.opt:
        _capi.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    total = args.clips
    mine = sh.partition_clips([1] * total, world)[rank]
    presets = cfg4_presets(total)
    n = int(30 * RATE)
    distinct = 8                                       # distinct input buffers, reused round-robin
    h_in = [torch.from_numpy((0.25 * np.random.default_rng(4000 + j).standard_normal((n, 2), dtype=np.float32))
                             .astype(np.float32)).pin_memory() for j in range(distinct)]
    keep, clips_meta = [], []
    for i in mine:
        st = dict(presets[i])
        seed = st.pop("seed")
        p, refl = rs.make_render_params(RATE, want_lufs=True, **st)
        np.random.seed(seed)
        taps, bases, noise = rs.draw_ir_randoms(RATE, p.ir_duration, refl, p.ir_max_delay, p.ir_split_time)
        h_noise = torch.from_numpy(noise).pin_memory()
        draws = _capi.make_draws(taps, bases, h_noise.numpy(), keep)
        N = int(lib.ars_render_out_len(p, n, 0))
        clips_meta.append((p, draws, h_noise, N, i))
    Nmax = max(m[3] for m in clips_meta)
    h_pcm = [torch.empty((Nmax, 8), dtype=torch.int16).pin_memory() for _ in range(2)]
    mets = [ArsMetrics() for _ in clips_meta]

    def batch():
        arr = (_capi.ArsClip * len(clips_meta))()
        for j, (p, draws, h_noise, N, i) in enumerate(clips_meta):
            k = arr[j]
            k.params = _capi.C.pointer(p)
            k.in_ = h_in[i % distinct].data_ptr()
            k.n, k.cin = n, 2
            k.draws = _capi.C.pointer(draws)
            k.out_pcm = h_pcm[j % 2].data_ptr()
            k.metrics = _capi.C.pointer(mets[j])
        _capi.check(lib.ars_render_batch(arr, len(clips_meta)), "ars_render_batch")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch()                                             # warm-up: plans, workspaces
    barrier()
    l0 = int(lib.ars_launch_count())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        batch()
    dt = time.perf_counter() - t0
    launches = int(lib.ars_launch_count()) - l0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        secs = 30.0 * total * args.steps
        print(json.dumps({"metric": "audio-seconds rendered per second (x realtime) @48kHz, batch of 30 s clips",
                          "value": secs / dt, "unit": "audio-seconds/s", "n_gpus": world, "steps": args.steps,
                          "warmup": 1, "ms_per_step": 1000 * dt / args.steps, "ms_per_clip": 1000 * dt / args.steps / total * world,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "gpu_launches": launches,
                          "config": {"workload": f"configs[3]: {total} x 30 s stereo clips, random presets (SURVEY 8d), "
                                                 "5.1 pan -> 7.1 map, metrics, int16 PCM out; host buffers in/out through "
                                                 "ars_render_batch", "clips": total, "parallelism": f"clip-sharded x{world}"},
                          "e2e": {"value": secs / dt, "unit": "audio-seconds/s", "h2d_bytes_per_step": int(len(mine) * n * 8),
                                  "d2h_bytes_per_step": int(sum(m[3] for m in clips_meta) * 16)},
                          "example_metrics": rs._metrics_dict(mets[0])}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["cfg4"])
    ap.add_argument("--clips", type=int, default=128, help="cfg4: clips in the batch (BASELINE: 1024)")
    ap.add_argument("--seconds", type=float, default=0.0, help="override the clip length (default: the config's)")
    ap.add_argument("--ir-seconds", type=float, default=0.0, help="external-IR length for cfg5")
    ap.add_argument("--cpu-sample-seconds", type=float, default=300.0)
    ap.add_argument("--ref-sample-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cin", type=int, default=0, help="experiments: override the clip's channel count")
    ap.add_argument("--air", type=float, default=None, help="experiments: override the air-absorption setting")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=INT",
                    help="library option for experiments (ars_set_option), e.g. --opt air_fold=0")
    args = ap.parse_args()
    if args.workload == "cfg4" and args.impl != "reference":
        run_cfg4(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
