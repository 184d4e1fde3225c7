"""CPU-only checks (-m "not gpu"): the C-ABI library loads and exports every symbol include/ars_b200.h declares,
fails loudly without a GPU, the FFT engine's host emulation agrees with a double-precision reference, the WAV codec
round-trips, and the multi-GPU sharding logic works over a world-size-2 gloo group."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ars_b200 import _capi, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "ars_b200.h")).read()
    declared = set(re.findall(r"ARS_API\s+[\w\s\*]+?\b(ars_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = _capi.load_library()
    for name in declared:
        assert hasattr(lib, name), f"libars_b200.so does not export {name}"
    assert declared == set(_capi.PROTOTYPES), declared ^ set(_capi.PROTOTYPES)
    assert b"sm_100a" in lib.ars_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ars_b200
    from ars_b200 import _capi
    with pytest.raises(ars_b200.ArsError):
        _capi.init(0)
    with pytest.raises(ars_b200.ArsError):
        ars_b200.raytracer_studio.apply_surround_panning_3d(np.zeros((10, 2), np.float32), .5, .5, .5)
    # stage calls made without ars_init report "no device", they do not compute
    lib = _capi.load_library()
    out = np.zeros(12, np.int16)
    assert lib.ars_pcm16(np.zeros(12, np.float32).ctypes.data, 12, out.ctypes.data) == 3


def test_fft_engine_host_emulation():
    src = os.path.join(ROOT, "tests", "host_emul", "fft_emul.cu")
    exe = os.path.join(ROOT, "ars_b200", "build", "fft_emul")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "ars_b200", "csrc", "fft.cuh"))):
        subprocess.run(["nvcc", "-std=c++17", "-O2", "-arch=sm_100a", "-o", exe, src], check=True, capture_output=True)
    r = subprocess.run([exe] + [str(i) for i in range(1, 20)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    r = subprocess.run([exe, "blue", "1", "2", "3", "37", "4099", "6000", "70001"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    r = subprocess.run([exe, "ols"], capture_output=True, text=True)      # overlap-save wiring incl. block ranges
    assert r.returncode == 0, r.stdout[-2000:]
    r = subprocess.run([exe, "irs"], capture_output=True, text=True)      # short-IR spectrum route
    assert r.returncode == 0, r.stdout[-2000:]
    env = dict(os.environ, ARS_FFT_PLAN="6,6,7")
    r = subprocess.run([exe, "19"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout[-2000:]


def test_float32_evaluation_of_float64_gain_products_is_exact_or_flagged():
    """The final pass forms RN32(RN64(x * gain)) (numpy's `audio * np.float64 gain` stored into float32, rs.py:475-494,
    550-553) from float32 pieces of the gain and redoes a frame in float64 only when a tie test fires (epilogue.cu: prod2).
    tests/host_emul/prod_emul.c replays that sequence with libm's correctly rounded fmaf: no unflagged result may differ
    from the float64 evaluation in a single bit -- random operands, operands on and next to float32 rounding boundaries,
    double-rounding traps, signed zeros."""
    src = os.path.join(ROOT, "tests", "host_emul", "prod_emul.c")
    exe = os.path.join(ROOT, "ars_b200", "build", "prod_emul")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True, capture_output=True)
    r = subprocess.run([exe, "20"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "wrong 0" in r.stdout


def test_scalar_prologue_matches_golden_on_cpu(golden):
    from ars_b200 import raytracer_studio as rs
    g = golden("scalars")
    halls = [str(h) for h in g["halls"]]
    for row in g["table"]:
        hall = halls[int(row[0])]
        room, x, y, z, dif, dw, e, l = row[1:9]
        assert tuple(float(v) for v in rs.adjust_parameters_for_3d(hall, room, z)) == tuple(row[9:13])
        assert float(rs.compute_final_directionality_3d(x, y, z, hall, dif, dw)) == row[13]
        ae, al = rs.adapt_early_late_levels(dw, e, l)
        assert (float(ae), float(al)) == (row[14], row[15])


def test_rng_replay_matches_oracle():
    import ars_oracle as orc
    from ars_b200 import raytracer_studio as rs
    for seed, args in ((3, (48000, 1.88552, 42, 0.07135, 0.09514)), (9, (100, 0.4, 20, 0.0175, 0.021))):
        np.random.seed(seed)
        a = rs.draw_ir_randoms(*args)
        np.random.seed(seed)
        b = orc.draw_ir_randoms(*args)
        for u, v in zip(a, b):
            assert np.array_equal(u, v)


def test_wav_codec_round_trip(tmp_path):
    from ars_b200 import wavio
    g = np.random.default_rng(0)
    pcm = g.integers(-32768, 32767, (1000, 6)).astype(np.int16)
    p = str(tmp_path / "a.wav")
    wavio.write_pcm16(p, pcm, 44100)
    x, rate = wavio.read(p)
    assert rate == 44100 and x.shape == (1000, 6) and x.dtype == np.float32
    assert np.array_equal(x, pcm.astype(np.float32) / np.float32(32768.0))
    f = g.standard_normal((77, 1)).astype(np.float32)
    wavio.write_float32(p, f, 8000)
    y, rate = wavio.read(p)
    assert rate == 8000 and np.array_equal(y, f)


def test_partitioners():
    from ars_b200 import sharding as sh
    lengths = [100, 5, 70, 70, 30, 1, 99, 42]
    for world in (1, 2, 3, 8):
        parts = sh.partition_clips(lengths, world)
        assert sorted(sum(parts, [])) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)
    assert sh.block_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sh.frame_range_for_blocks(3, 6, 4096, 1000, 10 ** 6) == (3 * 4096 - 999, 6 * 4096)
    assert sh.frame_range_for_blocks(0, 2, 4096, 1000, 5000) == (0, 5000)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from ars_b200 import sharding as sh
    lengths = [10, 40, 20, 30, 25]
    mine = sh.partition_clips(lengths, world)[rank]
    local = {i: (-20.0 - i, -1.0 * i, -30.0 + rank) for i in mine}
    allm = sh.gather_metrics(local)
    peak = sh.reduce_peak(0.25 + rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, allm, peak))


def test_sharding_over_gloo_world_size_2():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    mine0, mine1 = got[0][1], got[1][1]
    assert sorted(mine0 + mine1) == [0, 1, 2, 3, 4] and not set(mine0) & set(mine1)
    for rank, mine, allm, peak in got:
        assert sorted(allm) == [0, 1, 2, 3, 4]            # every rank sees every clip's metrics
        assert allm[3][0] == -23.0
        assert peak == 1.25                               # max over ranks


def test_preset_schema_follows_the_reference_loader(tmp_path):
    """ars_b200/presets.py: the reference's preset JSON (rs.py:883-896, 913-927) as a batch-job description --
    defaults for missing / null / unparsable values, bool coercion, key renaming, save/load round trip."""
    from ars_b200 import presets

    s, ext = presets.preset_to_settings({})
    assert ext is False
    assert s == {"hall_type": "Room", "material": "Holz", "room_size": 100.0, "diffusion": 0.5, "air_absorption": 0.1,
                 "base_early_level": 0.8, "base_late_level": 0.6, "dry_wet": 0.5, "dry_wet_kill_start": 0.5,
                 "bass_gain": 1.0, "treble_gain": 1.0, "x_pos": 0.5, "y_pos": 0.5, "z_pos": 0.5,
                 "target_channel_layout": "5.1 (Standard)"}
    raw = {"use_external_ir": 1, "hall_type": "Cathedral", "room_size": "250", "diffusion": None, "dry_wet": "viel",
           "late_level": 1, "target_layout": "7.1 (Surround)", "_version": "4.1", "unknown": 3}
    s, ext = presets.preset_to_settings(raw)
    assert ext is True and s["hall_type"] == "Cathedral" and s["room_size"] == 250.0 and s["diffusion"] == 0.5
    assert s["dry_wet"] == 0.5 and s["base_late_level"] == 1.0 and isinstance(s["base_late_level"], float)
    assert s["target_channel_layout"] == "7.1 (Surround)" and "unknown" not in s
    back = presets.settings_to_preset(s, use_external_ir=ext, name="Dom")
    assert list(back)[:16] == list(presets.PRESET_KEYS) and back["_source_name"] == "Dom"
    p = tmp_path / "Dom_v4.json"
    presets.save_preset(str(p), back)
    s2, ext2 = presets.preset_to_settings(presets.load_preset(str(p)))
    assert (s2, ext2) == (s, ext)


def test_reciprocal_division_of_the_peak_guards_is_correctly_rounded():
    """tail_math.cuh: guard_div divides thousands of samples by one maximum as q = RN(v r), r = RN(1 / m), followed by two
    fused residual corrections.  Restated here in exact rational arithmetic: the result is the correctly rounded float32
    quotient (numpy's `x / max_val`, rs.py:403) for random and for adversarial divisors."""
    import math
    import random
    import struct
    from fractions import Fraction

    def f32(x):
        if x == 0:
            return Fraction(0)
        sgn, a = (-1 if x < 0 else 1), abs(x)
        e = math.floor(math.log2(float(a))) - 23
        while a / Fraction(2) ** e >= 2 ** 24:
            e += 1
        while a / Fraction(2) ** e < 2 ** 23:
            e -= 1
        q = a / Fraction(2) ** e
        n = q.numerator // q.denominator
        r = q - n
        if r > Fraction(1, 2) or (r == Fraction(1, 2) and n % 2 == 1):
            n += 1
        return sgn * n * Fraction(2) ** e

    def as_f32(v):
        return Fraction(struct.unpack("f", struct.pack("f", v))[0])

    random.seed(1)
    cases = []
    for i in range(3000):
        m = as_f32(random.uniform(1.0, 16.0) if i % 3 else random.uniform(1.0, 1.001))
        v = as_f32(random.uniform(-float(m), float(m))) * (Fraction(1) if i % 5 else Fraction(1, 2 ** random.randint(0, 40)))
        cases.append((v, m))
    for _ in range(400):           # divisors next to powers of two and with nearly-all-ones mantissas
        bits = 0x3f800000 + random.choice([0x7ffffe, 1, 2, 0x400000, 0x3fffff, 0x7ffffd]) + (random.randint(0, 3) << 23)
        m = Fraction(struct.unpack("f", struct.pack("I", bits))[0])
        cases.append((as_f32(random.uniform(-float(m), float(m))), m))
    for v, m in cases:
        r = f32(Fraction(1) / m)
        q = f32(v * r)
        q = f32(f32(v - q * m) * r + q)
        q = f32(f32(v - q * m) * r + q)
        assert q == f32(v / m), (float(v), float(m))


def test_traffic_record_reproduces_from_the_committed_capture():
    """bench.py's `roofline.traffic` comes from profiles/traffic.json; its round-2 entry must be what
    profiles/traffic_summary.py sums out of the committed ncu launch list (one whole render, caches kept)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rec = json.load(open(os.path.join(root, "profiles", "traffic.json")))["cfg3_r02"]
    out = subprocess.run([sys.executable, os.path.join(root, "profiles", "traffic_summary.py"),
                          os.path.join(root, "profiles", "r02_traffic_cfg3_caches_kept_final.csv"), "--json"],
                         capture_output=True, text=True, check=True).stdout
    got = json.loads(out)
    assert got["dram_bytes_per_render"] == rec["dram_bytes_per_render"]
    assert got["per_kernel"] == rec["per_kernel"]
    assert 20 <= got["launches"] <= 40                                   # (one render, not two)
    algo = rec["algorithmic_bytes_per_render"]
    assert algo == (8 + 2 * 8) * 14783999                               # 8 B in + 2 B per output channel per frame
    assert 2.5 < rec["dram_bytes_per_render"] / algo < 3.5
