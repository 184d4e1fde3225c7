#!/usr/bin/env python
"""Benchmark of the render hot path (BASELINE.json metric: audio-seconds rendered per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg5|cfg4|long]

One "step" = one complete render of the workload clip: procedural IR synthesis (or an external stereo IR) ->
convolution + air absorption + dry/wet + EQ -> pan -> layout map -> metrics -> int16 PCM.

  value     : clip-seconds rendered per second with the input clip already resident in HBM and the PCM result left in
              HBM (ars_render_dev), timed with CUDA events on the library's stream.
  e2e       : the same metric through the host-buffer C-ABI (ars_render_batch): pinned host input copied to the
              device, PCM frames + metrics copied back, every step, inside the timed region.
  e2e_numpy : the call the drop-in module exposes (raytracer_studio.render_array, pageable numpy in / numpy out).
  roofline  : SURVEY 8(d): algorithmic bytes of one render (8 B in + 2 B per output channel per frame) / ms_per_step
              against the measured HBM copy bandwidth; `kernels` lists every kernel >= 5 % of the step with its own
              event-timed duration, algorithmic bytes and (from the committed ncu capture) DRAM bytes.
  parity    : the GPU result of the timed workload against the CPU oracle's render of the same clip, same draws
              (N = 1 only; the oracle render doubles as `cpu_baseline`).
  N > 1     : one process per GPU (torchrun), every rank renders its own clip of the same shape (clips are
              independent: no collective on the data path), barrier + max over ranks.       scaling = weak
  --workload long --gpus N : ONE long mask-free render (BASELINE configs[4]) split by overlap-save block ranges over
              the N ranks (ars_b200.sharding.render_long_sharded).                             scaling = strong

`--impl reference` times the CPU oracle (oracle/ars_oracle.py: a numpy/scipy restatement of the reference, which is a
Python file and cannot travel) on the host cores, one full-length clip per process, same config keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE = 48000
METRIC = "audio-seconds rendered per second (x realtime) @48kHz, 8 s IR"
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
    # BASELINE.json configs[4] (single-GPU slice): long stereo render (x) dense external stereo IR, EQ flat;
    # default 600 s (x) 8 s, override with --seconds / --ir-seconds
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
LAYOUT_CHANNELS = {"Stereo": 2, "5.1 (Standard)": 6, "7.1 (Surround)": 8, "5.1.2 (Atmos Light)": 8}


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


def workload_config(w, seconds, world, ext_ir_seconds=None):
    """The config keys both arms print (so that the driver sees the same configuration on either side)."""
    s = w["settings"]
    n = int(seconds * RATE)
    ext = "ext_ir_seconds" in w
    L = int((ext_ir_seconds or w.get("ext_ir_seconds", 0)) * RATE) if ext else int(s["ir_duration"] * RATE) if "ir_duration" in s else None
    cfg = {"workload": w["desc"], "clip_seconds": seconds, "frames_in": n,
           "channels_out": LAYOUT_CHANNELS[s["target_channel_layout"]], "clips_per_step": world,
           "parallelism": f"clip-sharded x{world}"}
    if L is not None:
        cfg["ir_frames"] = L
        cfg["frames_out"] = n + L - 1
    return cfg


# ----------------------------------------------------------------------------- clocks ------
class ClockSampler:
    """nvidia-smi sampled every 100 ms while the timed region runs (B200_PROFILING.md recipe)."""
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
    allocated on the GPU's own NUMA node.  Best effort: silently does nothing when NVML is unavailable (and a no-op on
    boxes that expose one NUMA node)."""
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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ----------------------------------------------------------------------------- CPU arms -----
def oracle_render(w_name, seconds, seed_offset=0, keep=False):
    """One render of the workload clip by the CPU oracle.  -> (seconds taken, outputs or None)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ars_oracle as orc
    w = WORKLOADS[w_name]
    x = make_clip(w, seconds, seed_offset)
    s = w["settings"]
    np.random.seed(w["np_seed"] + seed_offset)
    t0 = time.perf_counter()
    if "ext_ir_seconds" in w:
        kw = {ORACLE_KW[k]: v for k, v in s.items() if k in ORACLE_KW}
        out = orc.render(x, RATE, external_ir=make_ir(w["ext_ir_seconds"]), **kw)
        final, pcm, met = out["final"], out["pcm"], out["metrics"]
    elif "ir_duration" in s:
        # same override the GPU arm uses: the reference's stages called with the forced IR duration
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
        met = orc.metrics(final, RATE)
        pcm = orc.pcm16(final)
    else:
        kw = {ORACLE_KW[k]: v for k, v in s.items() if k in ORACLE_KW}
        out = orc.render(x, RATE, **kw)
        final, pcm, met = out["final"], out["pcm"], out["metrics"]
    dt = time.perf_counter() - t0
    return dt, ((final, pcm, met) if keep else None)


def _oracle_job(args):
    return oracle_render(*args)[0]


def run_reference(args):
    """--impl reference: the CPU oracle on every host core (one FULL-LENGTH clip per process and step), the same
    workload and config keys as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    name = "cfg3" if args.workload in ("cfg4", "long") else args.workload
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    if args.ref_procs:
        cores = min(cores, args.ref_procs)
    seconds = args.ref_sample_seconds or args.seconds or w["seconds"]
    jobs = [(name, seconds, i) for i in range(cores)]
    warm = 0                            # no warm-up pass: a step is ~45 s of CPU work on every core and numpy has nothing to warm
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(warm):
            pool.map(_oracle_job, jobs)
        t0 = time.perf_counter()
        per_clip = []
        for _ in range(args.steps):
            per_clip += pool.map(_oracle_job, jobs)
        dt = time.perf_counter() - t0
    value = cores * seconds * args.steps / dt
    cfg = workload_config(w, seconds, 1)
    cfg.update({"clips_per_step": cores, "parallelism": f"{cores} host processes, one clip each (numpy/scipy FFTs are "
                                                         "single-threaded)", "warmup_done": warm})
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "audio-seconds/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "audio-seconds/s", "cores": cores, "kind": "port",
                             "sample": f"{cores} processes x one {seconds:g} s clip per step (the workload's own clip "
                                       f"length), oracle/ars_oracle.py; mean {np.mean(per_clip):.1f} s per clip and core"},
            "e2e": {"value": value, "unit": "audio-seconds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm ------
class Render:
    """One workload set up for repeated rendering: device-resident and pinned-host copies of the inputs."""

    def __init__(self, lib, capi, rs, w, seconds, rank, ir_seconds=None, torch=None, true_peak=False):
        self.lib, self.capi, self.rs, self.w, self.seconds, self.torch = lib, capi, rs, w, seconds, torch
        x = make_clip(w, seconds, rank)
        self.x_np = x
        x2 = x if x.ndim == 2 else x[:, None]
        self.n, self.cin = x2.shape
        self.ext = "ext_ir_seconds" in w
        self.p, refl = rs.make_render_params(RATE, want_lufs=True, external_ir=self.ext, want_true_peak_4x=true_peak,
                                             **w["settings"])
        self.L = 0
        self.h_ir = self.d_ir = None
        self.ir_np = None
        if self.ext:
            self.ir_np = make_ir(ir_seconds or w["ext_ir_seconds"])
            self.L = self.ir_np.shape[0]
            self.h_ir = torch.from_numpy(self.ir_np).pin_memory()
            self.d_ir = self.h_ir.cuda()
            self.taps, self.bases, self.noise = np.zeros(0, np.int64), np.zeros(0), np.zeros(0)
        else:
            np.random.seed(w["np_seed"] + rank)
            self.taps, self.bases, self.noise = rs.draw_ir_randoms(RATE, self.p.ir_duration, refl, self.p.ir_max_delay,
                                                                   self.p.ir_split_time)
        self.N = int(lib.ars_render_out_len(self.p, self.n, self.L))
        self.C = LAYOUT_CHANNELS[w["settings"]["target_channel_layout"]]
        self.h_in = torch.from_numpy(np.ascontiguousarray(x2)).pin_memory()
        self.h_noise = torch.from_numpy(self.noise).pin_memory()
        self.h_pcm = [torch.empty((self.N, self.C), dtype=torch.int16).pin_memory() for _ in range(2)]
        self.d_in = self.h_in.cuda()
        self.d_noise = self.h_noise.cuda()
        self.d_pcm = torch.empty((self.N, self.C), dtype=torch.int16, device="cuda")
        torch.cuda.synchronize()
        self.keep = []
        self.draws_h = capi.make_draws(self.taps, self.bases, self.h_noise.numpy(), self.keep)
        self.draws_d = capi.make_draws(self.taps, self.bases, int(self.d_noise.data_ptr()), self.keep)
        self.draws_d.noise_len = int(self.noise.size)
        self.m = capi.ArsMetrics()

    def step_dev(self, deferred_metrics=None):
        """One device-resident render.  deferred_metrics: an ArsMetrics the library fills in when the stream is next waited
        for (ars_render_dev_async: the host does not wait per render); None: the synchronous form, metrics in self.m."""
        fn = self.lib.ars_render_dev if deferred_metrics is None else self.lib.ars_render_dev_async
        self.capi.check(fn(self.p, self.d_in.data_ptr(), self.n, self.cin,
                           self.d_ir.data_ptr() if self.ext else None, self.L,
                           None if self.ext else self.draws_d, None, None, self.d_pcm.data_ptr(),
                           self.m if deferred_metrics is None else deferred_metrics), "ars_render_dev")

    def step_host(self, out_f32=None, out_pcm=None):
        self.capi.check(self.lib.ars_render(self.p, self.h_in.data_ptr(), self.n, self.cin,
                                            self.h_ir.data_ptr() if self.ext else None, self.L,
                                            None if self.ext else self.draws_h, None, out_f32,
                                            out_pcm if out_pcm is not None else self.h_pcm[0].data_ptr(), self.m),
                        "ars_render")

    def batch_host(self, count, metrics_list):
        C = self.capi.C
        clips = (self.capi.ArsClip * count)()
        for i in range(count):
            k = clips[i]
            k.params = C.pointer(self.p)
            k.in_ = self.h_in.data_ptr()
            k.n, k.cin = self.n, self.cin
            if self.ext:
                k.ext_ir, k.ext_ir_len = self.h_ir.data_ptr(), self.L
            else:
                k.draws = C.pointer(self.draws_h)
            k.out_pcm = self.h_pcm[i % 2].data_ptr()
            k.metrics = C.pointer(metrics_list[i])
        self.capi.check(self.lib.ars_render_batch(clips, count), "ars_render_batch")

    def time_dev(self, steps, warmup):
        warm = [self.capi.ArsMetrics() for _ in range(warmup)]
        for k in range(warmup):
            self.step_dev(warm[k])
        self.capi.check(self.lib.ars_sync(), "ars_sync")
        ms = self.capi.C.c_float(0)
        later = [self.capi.ArsMetrics() for _ in range(steps)]
        self.capi.check(self.lib.ars_timer_begin(), "timer")
        for k in range(steps):
            self.step_dev(later[k])
        self.capi.check(self.lib.ars_timer_end(self.capi.C.byref(ms)), "timer")
        self.m = later[-1]
        return float(ms.value) / steps

    def h2d_bytes(self):
        return int(self.h_in.numel() * 4 + self.noise.size * 8 + self.taps.size * 16 + self.L * 8)

    def d2h_bytes(self):
        return int(self.N * self.C * 2 + 56)

    def algorithmic_bytes(self):
        """compulsory traffic of an ideal fused pass: 8 B in + 2 B per output channel per frame."""
        return int((8 + 2 * self.C) * self.N)


def profile_kernels(lib, capi, r, step_ms, traffic):
    """One untimed render with every kernel bracketed by CUDA events (side stream and lanes off, so the kernels run
    one after another as under ncu).  -> (fft totals, kernels[] for the roofline)."""
    capi.set_option("side_stream", 0)
    capi.set_option("olsb_lanes", 1)
    try:
        capi.check(lib.ars_profile_begin(), "profile")
        r.step_dev()
        pl, pms, pbytes = capi.C.c_int64(0), capi.C.c_double(0), capi.C.c_double(0)
        capi.check(lib.ars_profile_end(capi.C.byref(pl), capi.C.byref(pms), capi.C.byref(pbytes)), "profile")
        rep = json.loads(lib.ars_profile_report().decode())
    finally:
        capi.set_option("side_stream", 1)
        capi.set_option("olsb_lanes", 0)       # 0 = the library's default
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    total = sum(k["ms"] for k in rep) or 1.0
    kernels = []
    for k in sorted(rep, key=lambda k: -k["ms"]):
        us = 1000 * k["ms"]
        e = {"name": k["name"], "launches": k["launches"], "us": us, "share_of_serialised_step": k["ms"] / total,
             "algorithmic_bytes": k["bytes"],
             "achieved_gbs": k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None}
        e["frac"] = e["achieved_gbs"] / peak if e["achieved_gbs"] else None
        t = (traffic or {}).get("per_kernel", {}).get(k["name"])
        if t is not None:
            e["dram_bytes"] = t
        if k["ms"] / total >= 0.05:
            kernels.append(e)
    return {"launches": int(pl.value), "ms": pms.value, "bytes": pbytes.value, "serialised_ms": total}, kernels


def parity_vs_oracle(r, oracle_out):
    """GPU result of the timed workload against the oracle's render of the same clip with the same draws."""
    import torch
    final_ref, pcm_ref, met_ref = oracle_out
    h_f32 = torch.empty((r.N, r.C), dtype=torch.float32).pin_memory()
    h_pcm = torch.empty((r.N, r.C), dtype=torch.int16).pin_memory()
    r.step_host(out_f32=h_f32.data_ptr(), out_pcm=h_pcm.data_ptr())
    got = h_f32.numpy()
    met = r.rs._metrics_dict(r.m)
    scale = max(1.0, float(np.max(np.abs(final_ref))))
    err = 0.0
    num = den = 0.0
    for lo in range(0, r.N, 1 << 20):            # chunked: the arrays are hundreds of MB
        a = got[lo:lo + (1 << 20)].astype(np.float64)
        b = final_ref[lo:lo + (1 << 20)].astype(np.float64)
        d = a - b
        err = max(err, float(np.max(np.abs(d))))
        num += float(np.sum(d * d))
        den += float(np.sum(b * b))
    dp = np.abs(h_pcm.numpy().astype(np.int32) - pcm_ref.astype(np.int32))
    out = {"against": "oracle/ars_oracle.py render of the same clip, same random draws (full size)",
           "frames": int(r.N), "channels": int(r.C), "max_err_fs": err / scale,
           "snr_db": float(10 * np.log10(den / num)) if num > 0 else float("inf"),
           "pcm_lsb_max": int(dp.max()), "pcm_diff_frac": float(np.mean(dp != 0)),
           "peak_db_diff": abs(met["true_peak_dbfs"] - met_ref["true_peak_dbfs"]),
           "rms_db_diff": abs(met["rms_dbfs"] - met_ref["rms_dbfs"]),
           "lufs_abs_diff": (abs(met["lufs"] - met_ref["lufs"]) if met["lufs"] is not None and met_ref.get("lufs") is not None else None),
           "tolerance": {"max_err_fs": 1e-5, "snr_db": 100.0, "pcm_lsb_max": 1}}
    out["ok"] = bool(out["max_err_fs"] <= 1e-5 and out["snr_db"] >= 100.0 and out["pcm_lsb_max"] <= 1)
    return out


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
    r = Render(lib, _capi, rs, w, seconds, rank, args.ir_seconds or None, torch, true_peak=args.true_peak)

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
    # warm-up through the SAME call as the timed loop (asynchronous renders back to back), so that whatever that path sets
    # up on first use -- the second state block / feed buffer, internal streams and events, pinned metric slots, the head
    # start between renders -- exists before the timed region starts
    n_warm = max(args.warmup, 3)          # (asynchronous renders alternate between two slots: three touch everything)
    warm = [ArsMetrics() for _ in range(n_warm)]
    for k in range(n_warm):
        r.step_dev(None if args.sync_steps else warm[k])
    _capi.check(lib.ars_sync(), "ars_sync")
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = int(lib.ars_launch_count())
    ms = _capi.C.c_float(0)
    # every step is one complete render (metrics included); the host enqueues them back to back and the metrics of all
    # steps are finished when the stop event has been waited for (ars_render_dev_async) -- with --sync-steps the host waits
    # for each render's metrics before it prepares the next one, and the GPU idles meanwhile
    later = [ArsMetrics() for _ in range(args.steps)]
    _capi.check(lib.ars_timer_begin(), "timer")
    t_host = time.perf_counter()
    for k in range(args.steps):
        r.step_dev(None if args.sync_steps else later[k])
    host_enqueue_ms = 1000 * (time.perf_counter() - t_host) / args.steps     # (host time to enqueue one render; < ms_per_step: the GPU is the limit)
    _capi.check(lib.ars_timer_end(_capi.C.byref(ms)), "timer")
    if not args.sync_steps:
        assert all(rs._metrics_dict(m) == rs._metrics_dict(later[0]) for m in later), "steps of the same render differ"
        r.m = later[-1]
    launches = int(lib.ars_launch_count()) - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(float(ms.value))
    metrics_dev = rs._metrics_dict(r.m)

    # ---- end-to-end timing (host buffers in, host buffers out) ----
    # the public batch call: `steps` clips from pinned host memory, PCM frames + metrics back to pinned host
    # memory; clip i+1's upload and clip i-1's download overlap clip i's compute inside the library
    ms_list = [ArsMetrics() for _ in range(args.steps)]
    for _ in range(max(1, min(args.warmup, 2))):
        r.step_host()
    r.batch_host(min(2, args.steps), ms_list)
    barrier()
    t0 = time.perf_counter()
    r.batch_host(args.steps, ms_list)
    e2e_ms = max_over_ranks(1000 * (time.perf_counter() - t0))
    barrier()
    t0 = time.perf_counter()
    r.step_host()
    single_ms = 1000 * (time.perf_counter() - t0)
    barrier()

    # ---- the call the drop-in module exposes: pageable numpy in, numpy out (draws, allocation, copies included) ----
    e2e_np = None
    if world == 1 and not args.no_numpy:
        ext_kw = dict(external_ir_data=r.ir_np) if r.ext else {}
        k_np = max(2, min(args.steps, 5))
        # steady state of a caller that keeps the previous result while asking for the next one: two result blocks of the
        # library's pinned pool exist before the clock starts (page-locking a fresh 236 MB block costs ~0.2 s, once)
        warm = []
        for _ in range(2):
            np.random.seed(w["np_seed"])
            warm.append(rs.render_array(r.x_np, RATE, want_float=False, **ext_kw, **w["settings"]))
        res = warm[-1]
        del warm
        t0 = time.perf_counter()
        for _ in range(k_np):
            np.random.seed(w["np_seed"])
            res = rs.render_array(r.x_np, RATE, want_float=False, **ext_kw, **w["settings"])
        np_ms = 1000 * (time.perf_counter() - t0) / k_np
        e2e_np = {"value": seconds / (np_ms * 1e-3), "unit": "audio-seconds/s", "ms_per_step": np_ms, "calls": k_np,
                  "call": "ars_b200.raytracer_studio.render_array(numpy clip, ...) -> numpy PCM + metrics: pageable "
                          "input array (staged through the library's pinned ring), the reference's random draws replayed on the "
                          "host, results returned as numpy arrays over the library's pinned result pool",
                  "lufs": res["metrics"]["lufs"]}
        del res

    # ---- per-kernel event timing of one render (serialised) ----
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload + "_r02")
    except Exception:
        pass
    step_ms = dev_ms / args.steps
    fft_tot, kernels = profile_kernels(lib, _capi, r, step_ms, traffic)
    for kv in args.opt:                                 # (the profile turned two options off and back to their defaults)
        _capi.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    olsb = int(lib.ars_olsb_count()) > 0
    folds = int(lib.ars_air_fold_count()) > 0
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    total_seconds = seconds * world * args.steps
    algo = r.algorithmic_bytes()
    achieved = algo / (step_ms * 1e-3) / 1e9
    cfg = workload_config(w, seconds, world, args.ir_seconds or None)
    cfg.update({"route": ("folded-air " if folds else "") + ("big-block overlap-save (one partition, fused middle pass)" if olsb
                          else "see DESIGN.md section 2"),
                "l2": "working set (input %d MB, FFT work buffers, output %d MB) exceeds the 126 MB L2; no flush needed"
                      % (r.n * r.cin * 4 >> 20, r.N * r.C * 2 >> 20),
                "device_loop": ("ars_render_dev per step, the host waits for each render's metrics" if args.sync_steps else
                                "ars_render_dev_async per step (complete renders enqueued back to back, every step's "
                                "metrics collected once the stop event has been waited for)")})
    line = {
        "metric": METRIC, "value": total_seconds / (dev_ms * 1e-3), "unit": "audio-seconds/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "e2e": {"value": total_seconds / (e2e_ms * 1e-3), "unit": "audio-seconds/s", "ms_per_step": e2e_ms / args.steps,
                "call": "ars_render_batch (pinned host buffers, copy/compute pipelined across the steps' clips)",
                "single_call_ms": single_ms, "h2d_bytes_per_step": r.h2d_bytes(), "d2h_bytes_per_step": r.d2h_bytes()},
        "gpu_launches": launches,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "whole render (every kernel of the step)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "algorithmic_bytes_per_step": algo,
                     "definition": "SURVEY 8(d): 8 B in + 2 B per output channel per output frame of an ideal single "
                                   "fused pass, divided by ms_per_step (all kernels, launch gaps included)",
                     "traffic": (traffic or {}).get("dram_bytes_per_render"),
                     "traffic_unit": "DRAM bytes read + written by one render (ncu --set full, all kernels)",
                     "traffic_source": (traffic or {}).get("source"),
                     "fft_pass_kernels": {"launches_per_step": fft_tot["launches"], "ms": fft_tot["ms"],
                                          "algorithmic_bytes": fft_tot["bytes"],
                                          "achieved_gbs": fft_tot["bytes"] / (fft_tot["ms"] * 1e-3) / 1e9 if fft_tot["ms"] else None},
                     "serialised_step_ms": fft_tot["serialised_ms"],
                     "kernels": kernels},
        "metrics_of_last_render": metrics_dev,
    }
    if e2e_np:
        line["e2e_numpy"] = e2e_np
    if world == 1 and not args.no_cpu:
        dt, out = oracle_render(args.workload, seconds, 0, keep=True)
        line["cpu_baseline"] = {"value": seconds / dt, "unit": "audio-seconds/s", "cores": 1, "kind": "port",
                                "sample": f"the whole {seconds:g} s {args.workload} clip, same settings, oracle/ars_oracle.py "
                                          f"(numpy/scipy restatement of the reference; single-threaded like the "
                                          f"reference), {dt:.2f} s"}
        if not (args.cin or args.air is not None or args.ir_seconds):
            line["parity"] = parity_vs_oracle(r, out)
        del out
    if world == 1 and not args.no_extras and args.workload == "cfg3":
        del r
        torch.cuda.empty_cache()
        line["ir_sweep"], line["dense_ir"] = ir_sweep(lib, _capi, rs, torch, peak)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ir_sweep(lib, capi, rs, torch, peak):
    """BASELINE configs[4] on one GPU: 600 s stereo clip (x) DENSE external stereo IR of 0.1 ... 20 s, EQ flat, 5.1 out,
    device-resident.  -> (list of sweep points, the 8 s point as `dense_ir`)."""
    w = WORKLOADS["cfg5"]
    pts, dense = [], None
    for L in (0.1, 0.5, 2.0, 8.0, 20.0):
        r = Render(lib, capi, rs, w, w["seconds"], 0, L, torch)
        ms = r.time_dev(5, 2)
        algo = r.algorithmic_bytes()
        pt = {"ir_seconds": L, "ir_frames": r.L, "clip_seconds": w["seconds"], "ms_per_step": ms,
              "value": w["seconds"] / (ms * 1e-3), "unit": "audio-seconds/s",
              "roofline_frac": algo / (ms * 1e-3) / 1e9 / peak, "lufs": rs._metrics_dict(r.m)["lufs"]}
        pts.append(pt)
        if L == 8.0:
            dense = dict(pt, workload=w["desc"], note="dense Gaussian x exponential-decay IR (every tap non-zero); one "
                                                      "partition, 2^22-point blocks, no multiply-accumulate kernel: HBM-bound")
        del r
        torch.cuda.empty_cache()
    return pts, dense


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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["cfg4", "long"])
    ap.add_argument("--clips", type=int, default=128, help="cfg4: clips in the batch (BASELINE: 1024)")
    ap.add_argument("--seconds", type=float, default=0.0, help="override the clip length (default: the config's)")
    ap.add_argument("--ir-seconds", type=float, default=0.0, help="external-IR length for cfg5 / long")
    ap.add_argument("--ref-sample-seconds", type=float, default=0.0,
                    help="reference arm: clip length per process (default: the workload's own clip length)")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm: processes (default: every host core)")
    ap.add_argument("--sync-steps", action="store_true",
                    help="device-resident loop with the synchronous ars_render_dev (the host waits for every render's metrics)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle render (cpu_baseline + parity)")
    ap.add_argument("--no-extras", action="store_true", help="skip the IR-length sweep / dense-IR point")
    ap.add_argument("--no-numpy", action="store_true", help="skip the numpy-API timing")
    ap.add_argument("--cin", type=int, default=0, help="experiments: override the clip's channel count")
    ap.add_argument("--air", type=float, default=None, help="experiments: override the air-absorption setting")
    ap.add_argument("--true-peak", action="store_true", help="also compute the 4x-oversampled true peak (add-on metric)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=INT",
                    help="library option for experiments (ars_set_option), e.g. --opt air_fold=0")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg4":
        run_cfg4(args)
    elif args.workload == "long":
        from ars_b200 import sharding
        sharding.bench_long(args, make_clip=make_clip, make_ir=make_ir, load_peaks=load_peaks, ClockSampler=ClockSampler)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
