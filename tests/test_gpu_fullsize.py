"""Full-size parity on the configurations the numbers are quoted on (BASELINE.json configs[2..4]): the CUDA path through
the C ABI against the CPU oracle on the same seeded inputs at the benchmark's own sizes.  Bars: float within 1e-5 of
full scale and >= 100 dB SNR, PCM within 1 LSB, metrics within 1e-3 dB (LUFS 5e-3 LU)."""
import os
import sys

import numpy as np
import pytest

import ars_oracle as orc

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (workload definitions, preset generator)

pytestmark = pytest.mark.gpu
TOL = 1e-5
RATE = 48000


def compare(got, want, what):
    """max error of full scale and SNR, chunked (the arrays are hundreds of MB)."""
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(1.0, float(np.max(np.abs(want))))
    err = num = den = 0.0
    for lo in range(0, got.shape[0], 1 << 20):
        a = got[lo:lo + (1 << 20)].astype(np.float64)
        b = want[lo:lo + (1 << 20)].astype(np.float64)
        d = a - b
        err = max(err, float(np.max(np.abs(d))))
        num += float(np.sum(d * d))
        den += float(np.sum(b * b))
    snr = 10 * np.log10(den / num) if num > 0 else np.inf
    assert err / scale <= TOL, f"{what}: max error {err / scale:.3e} of full scale"
    assert snr >= 100.0, f"{what}: SNR {snr:.1f} dB"
    return err / scale, snr


def pcm_check(got, want, what):
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1, f"{what}: PCM differs by {d.max()} LSB"
    frac = float(np.mean(d != 0))
    assert frac < 1e-2, f"{what}: {frac:.2e} of the PCM samples differ"
    return frac


def test_cfg3_full_size_vs_oracle_both_routes(rs):
    """BASELINE configs[2] exactly as bench.py times it: 300 s 6-channel clip, Cathedral / Stein, 8 s procedural IR,
    air 0.1, 5.1.2, metrics, PCM.  The default route (air ramp folded into the IR, big-block overlap-save) and the exact
    N-point route against the oracle's render (rs.py:338-408 + 464-571 + 674-711)."""
    from ars_b200 import _capi
    w = bench.WORKLOADS["cfg3"]
    dt, (final, pcm, met) = bench.oracle_render("cfg3", w["seconds"], 0, keep=True)
    x = bench.make_clip(w, w["seconds"], 0)
    lib = _capi.load_library()
    for name, opts in (("folded-air big-block route", {}), ("exact N-point route", {"air_fold": 0})):
        try:
            for k, v in opts.items():
                _capi.set_option(k, v)
            f0, o0 = int(lib.ars_air_fold_count()), int(lib.ars_olsb_count())
            np.random.seed(w["np_seed"])
            got = rs.render_array(x, RATE, **w["settings"])
            folded = int(lib.ars_air_fold_count()) > f0
            assert folded == (name.startswith("folded")), name
            assert (int(lib.ars_olsb_count()) > o0) == folded
        finally:
            _capi.set_option("air_fold", 1)
        assert got["final"].shape == (14783999, 8)
        compare(got["final"], final, name)
        pcm_check(got["pcm"], pcm, name)
        for k in ("true_peak_dbfs", "rms_dbfs"):
            assert abs(got["metrics"][k] - met[k]) <= 1e-3, (name, k)
        assert abs(got["metrics"]["lufs"] - met["lufs"]) <= 5e-3, name
        del got


def test_cfg4_eight_clips_at_real_length_vs_oracle(rs):
    """BASELINE configs[3]: 30 s stereo clips with the benchmark's random presets (halls, materials, rooms, air, EQ on
    half of them, 7.1 map) through the batch call, each against the oracle (np.random seeded per clip as bench.py does)."""
    presets = bench.cfg4_presets(8)
    n = 30 * RATE
    jobs, wants = [], []
    for i, st in enumerate(presets):
        st = dict(st)
        seed = st.pop("seed")
        x = (0.25 * np.random.default_rng(4000 + i).standard_normal((n, 2), dtype=np.float32)).astype(np.float32)
        jobs.append(dict(samples=x, rate=RATE, seed=seed, **st))
        kw = {bench.ORACLE_KW[k]: v for k, v in st.items()}
        np.random.seed(seed)
        wants.append(orc.render(x, RATE, **kw))
    res = rs.render_batch(jobs, want_float=True)
    for i, (r, wnt) in enumerate(zip(res, wants)):
        compare(r["final"], wnt["final"], f"cfg4 clip {i}")
        pcm_check(r["pcm"], wnt["pcm"], f"cfg4 clip {i}")
        assert abs(r["metrics"]["lufs"] - wnt["metrics"]["lufs"]) <= 5e-3
        assert abs(r["metrics"]["rms_dbfs"] - wnt["metrics"]["rms_dbfs"]) <= 1e-3


def test_cfg5_ten_minutes_by_twenty_second_dense_ir_vs_oracle(rs):
    """BASELINE configs[4] slice: 10 min stereo (x) 20 s DENSE stereo IR (960 000 taps, 2^22-point blocks, mirror form
    of the middle pass) against scipy's fftconvolve (rs.py:410-462), then the whole render's PCM and metrics."""
    from ars_b200 import _capi
    w = bench.WORKLOADS["cfg5"]
    x = bench.make_clip(w, 600, 0)
    ir = bench.make_ir(20.0)
    s = w["settings"]
    lib = _capi.load_library()
    want = orc.convolve_external(x, ir, s["dry_wet"], 1.0, 1.0, RATE, s["dry_wet_kill_start"])
    o0 = int(lib.ars_olsb_count())
    got = rs.convolve_audio_external_ir(x, ir, s["dry_wet"], 1.0, 1.0, RATE, s["dry_wet_kill_start"])
    assert int(lib.ars_olsb_count()) == o0 + 1
    assert got.shape == (600 * RATE + 20 * RATE - 1, 2)
    compare(got, want, "cfg5 600 s (x) 20 s stereo stage")
    del got
    six = orc.pan_5_1(want, s["x_pos"], s["y_pos"], s["z_pos"])
    final, _ = orc.map_layout(six, s["target_channel_layout"], RATE, s["z_pos"])
    met = orc.metrics(final, RATE)
    res = rs.render_array(x, RATE, external_ir_data=ir, want_float=False, **s)
    pcm_check(res["pcm"], orc.pcm16(final), "cfg5 render")
    assert abs(res["metrics"]["lufs"] - met["lufs"]) <= 5e-3
    assert abs(res["metrics"]["rms_dbfs"] - met["rms_dbfs"]) <= 1e-3
    assert abs(res["metrics"]["true_peak_dbfs"] - met["true_peak_dbfs"]) <= 1e-3
