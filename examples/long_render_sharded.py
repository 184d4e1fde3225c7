#!/usr/bin/env python
"""One long mask-free render split over the GPUs of a box by overlap-save block ranges (BASELINE configs[4]).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        examples/long_render_sharded.py [--seconds 600] [--ir-seconds 8] [--check]

Rank 0 broadcasts the stereo IR (NCCL), every rank uploads and convolves its own block range, the peak-guard maxima
are max-reduced, the ranks' last stage-output frames go to their successors, the loudness meter's hop energies are
summed, and the PCM segments are pushed into rank 0's whole-render array over NVLink (ars_b200.sharding.render_long_sharded).
--check also renders the whole clip on rank 0 alone and requires bit-identical PCM from both gather forms (peer pushes and
NCCL point-to-point sends; prints BIT-IDENTICAL).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from ars_b200 import _capi, raytracer_studio as rs, sharding as sh
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=600.0)
    ap.add_argument("--ir-seconds", type=float, default=8.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _capi.init(local)
    rate = 48000
    n, L = int(a.seconds * rate), int(a.ir_seconds * rate)
    x = (0.1 * np.random.default_rng(5).standard_normal((n, 2), dtype=np.float32)).astype(np.float32)
    g = np.random.default_rng(6)
    ir = g.standard_normal((L, 2), dtype=np.float32) * np.exp(-6.9 * np.arange(L, dtype=np.float32) / (0.6 * L))[:, None]
    ir = (ir / np.max(np.abs(ir)) / np.float32(20.0)).astype(np.float32)
    settings = dict(dry_wet=.5, dry_wet_kill_start=.5, bass_gain=1.0, treble_gain=1.0, x_pos=.5, y_pos=.5, z_pos=.5,
                    target_channel_layout="5.1 (Standard)")
    res = None
    times = []
    for _ in range(a.steps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sh.render_long_sharded(x, rate, ir, **settings)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
    # the segments reach rank 0 by peer pushes over a CUDA IPC mapping (default) or by NCCL point-to-point sends
    res_nccl = sh.render_long_sharded(x, rate, ir, gather="nccl", **settings) if world > 1 and a.check else None
    if world > 1:
        sh.release_peer_arrays()
    if rank == 0:
        out = {"world": world, "clip_seconds": a.seconds, "ir_seconds": a.ir_seconds,
               "wall_ms_incl_upload_and_gather": [round(1000 * t, 2) for t in times[1:]], "metrics": res["metrics"]}
        if a.check:
            whole = rs.render_array(x, rate, external_ir_data=ir, want_float=False, **settings)
            same = bool(np.array_equal(whole["pcm"], res["pcm"]))
            if res_nccl is not None:
                same = same and bool(np.array_equal(whole["pcm"], res_nccl["pcm"])) and res_nccl["metrics"] == res["metrics"]
            out["bit_identical_to_single_gpu"] = same
            out["metrics_single_gpu"] = whole["metrics"]
            lufs_ok = abs(whole["metrics"]["lufs"] - res["metrics"]["lufs"]) < 1e-9
            print("BIT-IDENTICAL" if same and lufs_ok else "MISMATCH")
        print(json.dumps(out))
        if a.check and not (same and lufs_ok):
            sys.exit(1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
