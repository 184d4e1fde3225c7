"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in module) against
  (1) golden vectors produced by the UNMODIFIED reference (tests/golden, oracle/make_golden.py), and
  (2) the CPU oracle (oracle/ars_oracle.py) on fresh seeded inputs.

Bars (BASELINE.json north_star): float stages within max |err| <= 1e-5 of full scale
(full scale = max(1, peak of the reference result)); integer / element-wise float32 stages
(dry/wet mix, pan, map, PCM16) bit-exact.
"""
import numpy as np
import pytest

import ars_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-5
HALLS = ["Plate", "Room", "Cathedral", "Garage"]
MATERIALS = ["Stein", "Holz", "Teppich", "Glas", "Beton", "Vorhang (schwer)", "Gummi"]
LAYOUTS = ["Stereo", "5.1 (Standard)", "7.1 (Surround)", "5.1.2 (Atmos Light)", "9.1.6"]


def rel_err(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    return float(np.max(np.abs(got - ref)) / max(1.0, float(np.max(np.abs(ref)))))


def snr_db(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    noise = np.sum((got - ref) ** 2)
    sig = np.sum(ref ** 2)
    if noise == 0:
        return np.inf
    return 10 * np.log10(sig / noise) if sig > 0 else -np.inf


# ------------------------------------------------------------------ scalars (host) ----------
def test_scalar_prologue_matches_golden(rs, golden):
    g = golden("scalars")
    halls = [str(h) for h in g["halls"]]
    for row in g["table"]:
        hall = halls[int(row[0])]
        room, x, y, z, dif, dw, e, l = row[1:9]
        assert tuple(float(v) for v in rs.adjust_parameters_for_3d(hall, room, z)) == tuple(row[9:13])
        assert float(rs.compute_final_directionality_3d(x, y, z, hall, dif, dw)) == row[13]
        ae, al = rs.adapt_early_late_levels(dw, e, l)
        assert (float(ae), float(al)) == (row[14], row[15])


# ------------------------------------------------------------------ IR synthesis -------------
def test_ir_synth_matches_golden(rs, golden):
    g = golden("ir")
    for i, m in enumerate(g["meta"]):
        rate, mat, dif, seed = int(m[0]), MATERIALS[int(m[2])], m[7], int(m[9])
        dur, refl, mdel, split, direc = m[10], int(m[11]), m[12], m[13], m[14]
        np.random.seed(seed)
        e, l = rs.generate_impulse_response_split_3d(rate, dur, refl, mdel, mat, direc, split, dif)
        assert e.dtype == np.float32 and l.dtype == np.float32
        assert e.shape == g[f"early{i}"].shape
        # taps: scatter + normalise replay float64->float32 roundings exactly
        assert np.array_equal(e, g[f"early{i}"]), f"early taps differ in case {i}"
        # tail: float64 boxcar / std / pow with a different summation order -> at most float32 ulps
        assert rel_err(l, g[f"late{i}"]) <= 2e-7, f"late tail case {i}"
    e, l = rs.generate_impulse_response_split_3d(0, 1.0, 10, 0.05, "Holz", 0.5, 0.05, 0.5)
    assert np.array_equal(e, g["early_bad"]) and np.array_equal(l, g["late_bad"])


def test_ir_synth_rng_consumption_matches_oracle(rs):
    # after the call the global generator must be where the reference leaves it
    for seed in (1, 2):
        np.random.seed(seed)
        rs.generate_impulse_response_split_3d(48000, 1.2, 30, 0.05, "Stein", 0.4, 0.07, 0.5)
        a = np.random.uniform()
        np.random.seed(seed)
        orc.generate_ir(48000, 1.2, 30, 0.05, "Stein", 0.4, 0.07, 0.5)
        assert a == np.random.uniform()


# ------------------------------------------------------------------ spectral stages ---------
@pytest.mark.parametrize("tag", ["odd", "even", "short"])
def test_air_filter_matches_golden(rs, golden, tag):
    g = golden("air_filter")
    sig = g[f"in_{tag}"]
    for air in (0.005, 0.1, 0.65, 1.7):
        got = rs.apply_simple_lp_filter(sig, 48000, air)
        assert rel_err(got, g[f"air_{tag}_{air}"]) <= TOL
    got = rs.apply_simple_lp_filter(sig, 3000, 0.5)
    assert rel_err(got, g[f"air_{tag}_lowrate"]) <= TOL


def test_dry_wet_mix_bit_exact(rs, golden):
    g = golden("dry_wet")
    for i in range(8):
        dw, ks = g[f"par{i}"]
        got = rs.dynamic_dry_wet_mix(g[f"dry{i}"], g[f"wet{i}"], dw, ks)
        assert got.dtype == np.float32
        assert np.array_equal(got, g[f"mix{i}"]), f"case {i}"


def test_convolve_split_matches_golden(rs, golden):
    g = golden("convolve")
    keys = [str(k) for k in g["x_keys"]]
    for i, p in enumerate(g["split_par"]):
        x = g["x_" + keys[int(p[0])]]
        el, ll, dw, b, t, ks, air = p[1:8]
        got = rs.convolve_audio_split_3d(x, g["early"], g["late"], el, ll, dw, b, t, 48000, ks, air)
        ref = g[f"split{i}"]
        assert got.dtype == np.float32
        assert rel_err(got, ref) <= TOL, f"split case {i}: {rel_err(got, ref):.3e}"
        if np.any(ref):
            assert snr_db(got, ref) >= 100.0, f"split case {i}: SNR {snr_db(got, ref):.1f} dB"


def test_convolve_external_matches_golden(rs, golden):
    g = golden("convolve")
    keys = [str(k) for k in g["x_keys"]]
    for i, p in enumerate(g["ext_par"]):
        x = g["x_" + keys[int(p[0])]]
        dw, b, t, ks = p[1:5]
        got = rs.convolve_audio_external_ir(x, g["ext_ir"], dw, b, t, 48000, ks)
        ref = g[f"ext{i}"]
        assert rel_err(got, ref) <= TOL, f"ext case {i}: {rel_err(got, ref):.3e}"
        assert snr_db(got, ref) >= 100.0
    bad = rs.convolve_audio_external_ir(g["x_stereo"], g["ext_ir"][:, 0], .5, 1.0, 1.0, 48000, .5)
    assert np.array_equal(bad, g["ext_badir"])
    assert rs.convolve_audio_external_ir(np.zeros((0, 2), np.float32), g["ext_ir"], .5).shape == (0, 2)
    assert rs.convolve_audio_split_3d(None, g["early"], g["late"], .5, .5, .5).shape == (0, 2)


# ------------------------------------------------------------------ pan / map --------------
def test_pan_and_map_bit_exact(rs, golden):
    g = golden("pan_map")
    keys = [str(k) for k in g["sig_keys"]]
    for idx, p in enumerate(g["par"]):
        s = g["s_" + keys[int(p[0])]]
        x, y, z = p[1:4]
        six = rs.apply_surround_panning_3d(s, x, y, z)
        assert np.array_equal(six, g[f"pan{idx}"]), f"pan case {idx}"
        for li, lay in enumerate(LAYOUTS):
            src = g[f"pan{idx}"].copy()
            m, names = rs.map_channels(src, lay, 48000, z)
            assert np.array_equal(m, g[f"map{idx}_{li}"]), f"map case {idx} layout {lay}"
            if lay in ("5.1 (Standard)", "9.1.6"):
                assert m is src           # the reference hands back the input object
    m, _ = rs.map_channels(g["pan_441"].copy(), "7.1 (Surround)", 44100, .8)
    assert np.array_equal(m, g["map_441_71"])
    m, _ = rs.map_channels(g["pan_441"].copy(), "5.1.2 (Atmos Light)", 44100, .8)
    assert np.array_equal(m, g["map_441_512"])
    assert rs.apply_surround_panning_3d(None, .5, .5, .5).shape == (0, 6)
    assert rs.map_channels(np.zeros((5, 4), np.float32), "Stereo", 48000)[0].shape == (0, 2)


def test_apply_delay(rs):
    a = np.random.default_rng(0).standard_normal((100, 3)).astype(np.float32)
    assert np.array_equal(rs.apply_delay(a, 7), orc.delay_rows(a, 7))
    assert np.array_equal(rs.apply_delay(a, 1000), np.zeros_like(a))
    assert rs.apply_delay(a, 0) is a


# ------------------------------------------------------------------ metrics / PCM ----------
def test_metrics_peak_rms_match_golden(rs, golden):
    g = golden("metrics_pipeline")
    for i in range(4):
        m = rs.calculate_audio_metrics(g[f"m_in{i}"], 48000)
        ref = g[f"m_out{i}"]
        for got, want in ((m["true_peak_dbfs"], ref[0]), (m["rms_dbfs"], ref[1])):
            if np.isinf(want):
                assert got == want
            else:
                assert abs(got - want) <= 1e-4


def test_lufs_matches_oracle_restatement(rs):
    g = np.random.default_rng(7)
    for n, c, amp, rate in ((48000 * 3, 2, 0.2, 48000), (44100 * 2 + 123, 6, 0.05, 44100), (30000, 1, 0.5, 16000)):
        d = (amp * g.standard_normal((n, c))).astype(np.float32)
        d[: n // 3] *= 0.01          # quiet stretch so the relative gate matters
        want = orc.metrics(d, rate)["lufs"]
        got = rs.calculate_audio_metrics(d, rate)["lufs"]
        assert abs(got - want) <= 2e-3, (got, want)
    short = (0.1 * g.standard_normal((1000, 2))).astype(np.float32)
    assert rs.calculate_audio_metrics(short, 48000)["lufs"] is None        # < 400 ms
    assert rs.calculate_audio_metrics(np.zeros((48000, 2), np.float32), 48000)["lufs"] == -np.inf


def test_lufs_fused_chain_equals_stagewise_chain(rs):
    """The fused loudness chain (three passes, hop energies) against the pass-per-stage chain: same per-sample arithmetic,
    only the summation order of a gating block's squares differs.  Lengths that end inside / at a hop or block edge."""
    from ars_b200 import _capi
    g = np.random.default_rng(17)
    try:
        for n, rate in ((48000 * 7 + 311, 48000), (44100 * 5, 44100), (8192 * 12, 48000), (19200 + 4800 * 3, 48000),
                        (96000 * 3 + 7, 96000), (19200, 48000)):
            d = (0.2 * g.standard_normal((n, 2))).astype(np.float32)
            d[n // 4: n // 2] *= 0.003
            out = []
            for fused in (1, 0):
                _capi.set_option("lufs_fused", fused)
                out.append(rs.calculate_audio_metrics(d, rate)["lufs"])
            assert abs(out[0] - out[1]) <= 1e-9, (n, rate, out)
            assert abs(out[0] - orc.metrics(d, rate)["lufs"]) <= 2e-3
    finally:
        _capi.set_option("lufs_fused", 1)


def test_pcm16_bit_exact(rs):
    g = np.random.default_rng(8)
    x = (0.6 * g.standard_normal((20001, 6))).astype(np.float32)
    x[5, 2] = np.nan
    x[6, 1] = np.inf
    x[7, 0] = -np.inf
    x[8, :] = [0.5 / 32767, 1.5 / 32767, 2.5 / 32767, -0.5 / 32767, -1.5 / 32767, 0.99995]
    assert np.array_equal(rs.float_to_pcm16(x), orc.pcm16(x))


# ------------------------------------------------------------------ whole pipeline ---------
def test_pipeline_matches_golden(rs, golden):
    g = golden("metrics_pipeline")
    for i, p in enumerate(g["pipe_par"]):
        n, ch, rate, ext = int(p[0]), int(p[1]), int(p[2]), bool(p[3])
        hall, room, dif, air, e, l, dw, ks, b, t, x, y, z = (HALLS[int(p[4])],) + tuple(p[5:17])
        mat, lay, seed = MATERIALS[int(p[17])], LAYOUTS[int(p[18])], int(p[19])
        np.random.seed(seed)
        res = rs.render_array(g[f"pipe_in{i}"], rate, external_ir_data=g["pipe_ir"] if ext else None,
                              hall_type=hall, room_size=room, diffusion=dif, air_absorption=air, base_early_level=e,
                              base_late_level=l, dry_wet=dw, dry_wet_kill_start=ks, bass_gain=b, treble_gain=t,
                              x_pos=x, y_pos=y, z_pos=z, material=mat, target_channel_layout=lay)
        ref = g[f"pipe_out{i}"]                       # what the reference handed to sf.write (clipped float32)
        got = np.clip(res["final"], -0.9999, 0.9999)
        assert rel_err(got, ref) <= TOL, f"pipeline case {i}: {rel_err(got, ref):.3e}"
        # int16 frames: identical rounding rule on nearly identical floats -> at most 1 LSB apart, almost never
        ref_pcm = orc.pcm16(ref)
        diff = np.abs(res["pcm"].astype(np.int32) - ref_pcm.astype(np.int32))
        assert diff.max() <= 1 and np.mean(diff != 0) < 1e-2, (diff.max(), np.mean(diff != 0))
        # and bit-exact against the shared rule applied to our own float output
        assert np.array_equal(res["pcm"], orc.pcm16(res["final"]))
        assert rs._metrics_text(res["metrics"]) == str(g[f"pipe_text{i}"]), (rs._metrics_text(res["metrics"]),
                                                                         str(g[f"pipe_text{i}"]))


def test_file_level_entry_point(rs, tmp_path):
    from ars_b200 import wavio
    g = np.random.default_rng(9)
    x = (0.25 * g.standard_normal((24000, 2))).astype(np.float32)
    src = tmp_path / "in.wav"
    wavio.write_float32(str(src), x, 48000)
    np.random.seed(5)
    p1, p2, text = rs.apply_raytrace_convolution_3d(str(src), None, False, "Room", 120, .5, .1, .8, .6, .5, .5, 1.2, .9,
                                                    .4, .5, .6, "Holz", "7.1 (Surround)")
    assert p1 is not None and p1 == p2 and text.startswith("LUFS: ")
    pcm, rate = wavio.read(p1)
    np.random.seed(5)
    ref = orc.render(x, 48000, hall="Room", room_size=120, diffusion=.5, air=.1, early=.8, late=.6, dry_wet_amount=.5,
                     kill_start=.5, bass=1.2, treble=.9, x=.4, y=.5, z=.6, material="Holz", layout="7.1 (Surround)")
    assert rate == 48000 and pcm.shape == ref["pcm"].shape
    d = np.abs(np.round(pcm * 32768).astype(np.int32) - ref["pcm"].astype(np.int32))
    assert d.max() <= 1
    bad = rs.apply_raytrace_convolution_3d(str(tmp_path / "missing.wav"), None, False, "Room", 120, .5, .1, .8, .6, .5,
                                           .5, 1., 1., .5, .5, .5, "Holz", "Stereo")
    assert bad[0] is None and bad[1] is None and bad[2].startswith("Fehler beim Laden")


# ------------------------------------------------------------------ larger sizes ------------
def test_cfg1_external_ir_vs_oracle(rs):
    """BASELINE configs[0]: 10 s 48 kHz stereo sweep (x) 2 s synthetic stereo IR, stereo out."""
    rate = 48000
    t = np.arange(10 * rate) / rate
    l = (0.5 * np.sin(2 * np.pi * (20 * t + (19980 / 20) * t * t))).astype(np.float32)
    x = np.stack((l, l[::-1]), axis=1)
    g = np.random.default_rng(1)
    ir = (g.standard_normal((2 * rate, 2)) * np.exp(-np.arange(2 * rate) / (0.3 * rate))[:, None]).astype(np.float32)
    ir /= np.max(np.abs(ir))
    for bass, treble in ((1.0, 1.0), (1.2, 0.9)):
        want = orc.render(x, rate, external_ir=ir, dry_wet_amount=.5, kill_start=.5, bass=bass, treble=treble,
                          layout="Stereo", with_lufs=False)
        got = rs.render_array(x, rate, external_ir_data=ir, want_stereo=True, dry_wet=.5, dry_wet_kill_start=.5,
                              bass_gain=bass, treble_gain=treble, target_channel_layout="Stereo")
        assert rel_err(got["stereo"], want["stereo"]) <= TOL
        assert rel_err(got["final"], want["final"]) <= TOL
        assert snr_db(got["final"], want["final"]) >= 100.0


def test_cfg2_like_room_render_vs_oracle(rs):
    """BASELINE configs[1] at 1/6 length: mono -> 5.1, Room / Holz / 200 m^3, air 0.1, EQ 1.5 / 0.8."""
    rate = 48000
    x = (0.3 * np.random.default_rng(0).standard_normal(10 * rate)).astype(np.float32)
    kw = dict(hall="Room", room_size=200., diffusion=.5, air=.1, early=.8, late=.6, dry_wet_amount=.6, kill_start=.5,
              bass=1.5, treble=.8, x=.3, y=.4, z=.6, material="Holz", layout="5.1 (Standard)")
    np.random.seed(11)
    want = orc.render(x, rate, **kw)
    np.random.seed(11)
    got = rs.render_array(x, rate, want_stereo=True, hall_type="Room", room_size=200., diffusion=.5, air_absorption=.1,
                          base_early_level=.8, base_late_level=.6, dry_wet=.6, dry_wet_kill_start=.5, bass_gain=1.5,
                          treble_gain=.8, x_pos=.3, y_pos=.4, z_pos=.6, material="Holz",
                          target_channel_layout="5.1 (Standard)")
    assert got["final"].shape == want["final"].shape == (10 * rate + 90504 - 1, 6)
    assert rel_err(got["stereo"], want["stereo"]) <= TOL
    assert rel_err(got["final"], want["final"]) <= TOL
    assert snr_db(got["final"], want["final"]) >= 100.0
    d = np.abs(got["pcm"].astype(np.int32) - want["pcm"].astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) < 1e-2
    for k in ("true_peak_dbfs", "rms_dbfs"):
        assert abs(got["metrics"][k] - want["metrics"][k]) <= 1e-3
    assert abs(got["metrics"]["lufs"] - want["metrics"]["lufs"]) <= 5e-3


def test_linearity_and_shift_at_full_size(rs):
    """Size-independent properties at a BASELINE-scale length (60 s): with EQ/air on but every
    normaliser idle the render is linear, so render(a*x) == a*render(x)."""
    rate = 48000
    g = np.random.default_rng(3)
    x = (0.05 * g.standard_normal((60 * rate, 2))).astype(np.float32)
    kw = dict(hall_type="Cathedral", room_size=817., diffusion=.3, air_absorption=.2, dry_wet=.4, bass_gain=1.3,
              treble_gain=.7, target_channel_layout="5.1.2 (Atmos Light)", want_metrics=False, want_pcm=False)
    np.random.seed(21)
    a = rs.render_array(x, rate, **kw)["final"]
    np.random.seed(21)
    b = rs.render_array((2.0 * x).astype(np.float32), rate, **kw)["final"]
    assert np.max(np.abs(a)) < 0.5
    assert rel_err(b, 2.0 * a) <= 2e-6


def test_batch_pipeline_equals_single_renders(rs):
    """ars_render_batch (copy/compute overlap, two buffer slots) must reproduce ars_render bit for bit."""
    g = np.random.default_rng(12)
    rate = 48000
    jobs = []
    for i, (n, lay, hall) in enumerate([(30000, "7.1 (Surround)", "Room"), (52001, "Stereo", "Plate"),
                                        (41000, "5.1.2 (Atmos Light)", "Cathedral"), (30000, "5.1 (Standard)", "Room"),
                                        (64000, "7.1 (Surround)", "Plate")]):
        jobs.append(dict(samples=(0.3 * g.standard_normal((n, 2))).astype(np.float32), rate=rate, seed=100 + i,
                         hall_type=hall, room_size=50. + 40 * i, air_absorption=0.1 * i, bass_gain=1.0 + 0.2 * i,
                         treble_gain=1.0, dry_wet=0.3 + 0.1 * i, x_pos=.2 * i, target_channel_layout=lay))
    ir = (g.standard_normal((3000, 2)) * np.exp(-np.arange(3000) / 600.0)[:, None]).astype(np.float32)
    jobs.append(dict(samples=(0.3 * g.standard_normal((20000, 2))).astype(np.float32), rate=rate,
                     external_ir_data=ir, bass_gain=1.3, target_channel_layout="5.1 (Standard)"))
    # clips with more than two channels (the reference reads channels 0-1 only, rs.py:345)
    jobs.append(dict(samples=(0.3 * g.standard_normal((150001, 6))).astype(np.float32), rate=rate, seed=300,
                     hall_type="Cathedral", air_absorption=0.1, target_channel_layout="5.1.2 (Atmos Light)"))
    jobs.append(dict(samples=(0.3 * g.standard_normal((90000, 3))).astype(np.float32), rate=rate, seed=301,
                     hall_type="Room", bass_gain=1.2, target_channel_layout="Stereo"))
    jobs.append(dict(samples=(0.3 * g.standard_normal((140000, 6))).astype(np.float32), rate=rate, seed=302,
                     hall_type="Plate", air_absorption=0.0, target_channel_layout="7.1 (Surround)"))
    got = rs.render_batch(jobs, want_float=True)
    for job, b in zip(jobs, got):
        job = dict(job)
        seed = job.pop("seed", None)
        if seed is not None:
            np.random.seed(seed)
        a = rs.render_array(job.pop("samples"), job.pop("rate"), **job)
        assert np.array_equal(a["final"], b["final"])
        assert np.array_equal(a["pcm"], b["pcm"])
        assert a["metrics"] == b["metrics"]


@pytest.mark.parametrize("logf", [12, 13])
def test_overlap_save_path_matches_oracle_and_n_point_path(rs, golden, logf):
    """Mask-free renders run through the partitioned overlap-save convolution (K2-K4).  Same bars as the N-point
    path, checked against the golden vectors, the oracle, and against the N-point path itself."""
    from ars_b200 import _capi
    g = golden("convolve")
    keys = [str(k) for k in g["x_keys"]]
    try:
        _capi.set_option("upols_logf", logf)
        for use in (1, 0):
            _capi.set_option("upols", use)
            for i, p in enumerate(g["split_par"]):
                el, ll, dw, b, t, ks, air = p[1:8]
                if air > 0.01 or b != 1.0 or t != 1.0:
                    continue                                    # masked cases never take the overlap-save path
                got = rs.convolve_audio_split_3d(g["x_" + keys[int(p[0])]], g["early"], g["late"], el, ll, dw, b, t, 48000,
                                                 ks, air)
                assert rel_err(got, g[f"split{i}"]) <= TOL, (use, i, rel_err(got, g[f"split{i}"]))
            for i, p in enumerate(g["ext_par"]):
                dw, b, t, ks = p[1:5]
                if b != 1.0 or t != 1.0:
                    continue
                got = rs.convolve_audio_external_ir(g["x_" + keys[int(p[0])]], g["ext_ir"], dw, b, t, 48000, ks)
                assert rel_err(got, g[f"ext{i}"]) <= TOL, (use, i)
                assert snr_db(got, g[f"ext{i}"]) >= 100.0
        # a longer dense stereo IR (many partitions, one of them silent) against scipy through the oracle
        rg = np.random.default_rng(31)
        x = (0.2 * rg.standard_normal((150001, 2))).astype(np.float32)
        ir = (rg.standard_normal((40000, 2)) * np.exp(-np.arange(40000) / 9000.0)[:, None]).astype(np.float32)
        ir[8192:12288] = 0
        ir /= np.max(np.abs(ir)) * 40
        want = orc.convolve_external(x, ir, .7, 1.0, 1.0, 48000, .5)
        _capi.set_option("upols", 1)
        a = rs.convolve_audio_external_ir(x, ir, .7, 1.0, 1.0, 48000, .5)
        _capi.set_option("upols", 0)
        b = rs.convolve_audio_external_ir(x, ir, .7, 1.0, 1.0, 48000, .5)
        assert rel_err(a, want) <= TOL and rel_err(b, want) <= TOL and rel_err(a, b) <= TOL
        assert snr_db(a, want) >= 100.0
        # whole render, procedural IR, mono input, 7.1
        kw = dict(hall_type="Plate", room_size=300., air_absorption=0.0, dry_wet=.45, bass_gain=1.0, treble_gain=1.0,
                  x_pos=.6, y_pos=.3, z_pos=.8, material="Glas", target_channel_layout="7.1 (Surround)")
        xm = (0.3 * rg.standard_normal(96000)).astype(np.float32)
        res = []
        for use in (1, 0):
            _capi.set_option("upols", use)
            np.random.seed(77)
            res.append(rs.render_array(xm, 48000, want_stereo=True, **kw))
        np.random.seed(77)
        want = orc.render(xm, 48000, hall="Plate", room_size=300., air=0.0, dry_wet_amount=.45, x=.6, y=.3, z=.8,
                          material="Glas", layout="7.1 (Surround)")
        for r in res:
            assert rel_err(r["final"], want["final"]) <= TOL
            assert abs(r["metrics"]["lufs"] - want["metrics"]["lufs"]) <= 5e-3
    finally:
        _capi.set_option("upols", 1)
        _capi.set_option("upols_logf", 13)


def _emulated_sharded_render(sh, x, rate, ir, settings, world):
    """`world` ranks of a block-sharded long render one after another on the one GPU; the collectives between the phases
    are done with torch ops on the ranks' state blocks.  -> (pcm, metrics)"""
    import torch
    d_ir = torch.from_numpy(ir).cuda()
    ranks = [sh.LongRenderRank(x, rate, d_ir, settings, r, world) for r in range(world)]

    def reduce(view_of, op):
        vals = torch.stack([view_of(r).clone() for r in ranks])
        red = vals.max(dim=0).values if op == "max" else vals.sum(dim=0)
        for r in ranks:
            view_of(r).copy_(red)

    with torch.cuda.stream(sh.lib_stream()):
        for r in ranks: r.convolve()
        reduce(lambda r: r.words()[0:4], "max")
        tails = [r.y_tail() for r in ranks]
        for k in range(1, world): ranks[k].set_halo(tails[k - 1])
        for r in ranks: r.pan_max()
        reduce(lambda r: r.words()[4:5], "max")
        for r in ranks: r.map_max()
        reduce(lambda r: r.words()[5:6], "max")
        for r in ranks: r.final()
        for r in ranks: r.loudness_hops()
        reduce(lambda r: r.d_hops, "sum")
        reduce(lambda r: r.words()[8:10], "max")
        reduce(lambda r: r.sumsq(), "sum")
        status = ranks[0].loudness_gate()
        pcm = torch.cat([r.d_pcm[:r.frames()] for r in ranks], dim=0)
        m = ranks[0].metrics(status)
    return pcm.cpu().numpy(), m


@pytest.mark.parametrize("layout", ["5.1 (Standard)", "5.1.2 (Atmos Light)", "Stereo"])
@pytest.mark.parametrize("route", ["big-block", "partitioned"])
def test_block_sharded_long_render_is_bit_identical(rs, layout, route):
    """splitting a mask-free render by overlap-save block ranges must not change a bit --
    on either convolution route, with the peak guards firing, with a layout delay reaching into the previous rank's
    frames, and with more ranks than blocks."""
    from ars_b200 import _capi, sharding as sh
    g = np.random.default_rng(41)
    rate = 48000
    n = 1400011 if route == "big-block" else 230011
    x = (0.9 * g.standard_normal((n, 2))).astype(np.float32)          # loud: the stereo guard and the pan guard fire
    ir = (g.standard_normal((21000, 2)) * np.exp(-np.arange(21000) / 5000.0)[:, None]).astype(np.float32)
    ir /= np.max(np.abs(ir)) * 8
    settings = dict(dry_wet=.6, dry_wet_kill_start=.5, bass_gain=1.0, treble_gain=1.0, x_pos=.3, y_pos=.6, z_pos=.7,
                    target_channel_layout=layout)
    try:
        _capi.set_option("olsb", 1 if route == "big-block" else 0)
        c0 = int(_capi.load_library().ars_olsb_count())
        whole = rs.render_array(x, rate, external_ir_data=ir, **settings)
        assert (int(_capi.load_library().ars_olsb_count()) > c0) == (route == "big-block")
        for world in (1, 3, 8):
            pcm, m = _emulated_sharded_render(sh, x, rate, ir, settings, world)
            assert pcm.shape == whole["pcm"].shape
            assert np.array_equal(pcm, whole["pcm"]), f"world {world}: PCM differs from the single-GPU render"
            assert m["true_peak_dbfs"] == whole["metrics"]["true_peak_dbfs"]
            assert abs(m["rms_dbfs"] - whole["metrics"]["rms_dbfs"]) < 1e-6
            assert abs(m["lufs"] - whole["metrics"]["lufs"]) < 1e-9, (world, m["lufs"], whole["metrics"]["lufs"])
    finally:
        _capi.set_option("olsb", 1)


def test_block_sharded_long_render_on_real_ranks():
    """The same over torch.distributed / NCCL with one process per GPU (needs >= 2 GPUs; `gpurun --gpus 2`): the PCM
    gathered on rank 0 must equal the single-GPU render bit for bit."""
    import os
    import subprocess
    import sys
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "examples", "long_render_sharded.py"), "--check", "--seconds", "90",
           "--ir-seconds", "1.5"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BIT-IDENTICAL" in r.stdout, r.stdout[-2000:]


def test_random_presets_vs_oracle(rs):
    """Presets drawn from the UI ranges, plus the edge values the reference special-cases:
    dw 0 / 1, kill 1, air just below / above its gate, gains inside np.isclose(1), odd / even N, unknown names."""
    g = np.random.default_rng(2024)
    halls = ["Plate", "Room", "Cathedral", "Aula"]
    mats = ["Stein", "Holz", "Teppich", "Glas", "Beton", "Vorhang (schwer)", "Filz"]
    lays = ["Stereo", "5.1 (Standard)", "7.1 (Surround)", "5.1.2 (Atmos Light)", "22.2"]
    edge = [dict(dry_wet=0.0), dict(dry_wet=1.0), dict(dry_wet_kill_start=1.0, dry_wet=.9), dict(air_absorption=0.0099),
            dict(air_absorption=0.0101), dict(bass_gain=1.000009, treble_gain=0.999995), dict(bass_gain=1.00002),
            dict(base_early_level=0.0), dict(base_late_level=1e-7), dict(x_pos=0.0, y_pos=1.0, z_pos=0.0),
            dict(x_pos=1.7, y_pos=-0.3, z_pos=2.0), dict(dry_wet_kill_start=0.0, dry_wet=.5)]
    worst = 0.0
    for i in range(28):
        n = int(g.integers(9000, 30000))
        cin = int(g.choice([1, 2, 2, 3]))
        x = (g.uniform(0.05, 0.9) * g.standard_normal((n, cin))).astype(np.float32)
        if cin == 1 and i % 2:
            x = x[:, 0]
        kw = dict(hall_type=str(g.choice(halls)), material=str(g.choice(mats)), room_size=float(10 * g.integers(1, 101)),
                  diffusion=float(g.uniform()), air_absorption=float(g.uniform()), dry_wet=float(g.uniform()),
                  dry_wet_kill_start=float(g.uniform()), x_pos=float(g.uniform()), y_pos=float(g.uniform()),
                  z_pos=float(g.uniform()), base_early_level=float(g.uniform(0, 2)), base_late_level=float(g.uniform(0, 2)),
                  bass_gain=1.0 if g.uniform() < .5 else float(g.uniform(.1, 5)),
                  treble_gain=1.0 if g.uniform() < .5 else float(g.uniform(.1, 5)),
                  target_channel_layout=str(g.choice(lays)))
        if i < len(edge):
            kw.update(edge[i])
        rate = int(g.choice([16000, 44100, 48000]))
        np.random.seed(500 + i)
        got = rs.render_array(x, rate, want_stereo=True, **kw)
        np.random.seed(500 + i)
        want = orc.render(x, rate, hall=kw["hall_type"], room_size=kw["room_size"], diffusion=kw["diffusion"],
                          air=kw["air_absorption"], early=kw["base_early_level"], late=kw["base_late_level"],
                          dry_wet_amount=kw["dry_wet"], kill_start=kw["dry_wet_kill_start"], bass=kw["bass_gain"],
                          treble=kw["treble_gain"], x=kw["x_pos"], y=kw["y_pos"], z=kw["z_pos"], material=kw["material"],
                          layout=kw["target_channel_layout"])
        assert got["final"].shape == want["final"].shape, (i, kw)
        e = max(rel_err(got["stereo"], want["stereo"]), rel_err(got["final"], want["final"]))
        worst = max(worst, e)
        assert e <= TOL, (i, e, kw)
        d = np.abs(got["pcm"].astype(np.int32) - want["pcm"].astype(np.int32))
        assert d.max() <= 1, (i, kw)
        for k in ("true_peak_dbfs", "rms_dbfs"):
            a, b = got["metrics"][k], want["metrics"][k]
            assert (a == b) if np.isinf(b) else abs(a - b) <= 2e-3, (i, k, a, b)
        a, b = got["metrics"]["lufs"], want["metrics"]["lufs"]
        assert (a is None and b is None) or (np.isinf(b) and a == b) or abs(a - b) <= 1e-2, (i, a, b)
    print("worst relative error over the preset sweep:", worst)


def test_silent_and_tiny_inputs(rs):
    z = np.zeros((20000, 2), np.float32)
    np.random.seed(1)
    r = rs.render_array(z, 48000, hall_type="Room", bass_gain=2.0, target_channel_layout="7.1 (Surround)")
    assert not r["final"].any() and not r["pcm"].any()
    assert r["metrics"]["lufs"] == -np.inf and r["metrics"]["true_peak_dbfs"] == -np.inf
    t = (1e-12 * np.random.default_rng(0).standard_normal((15000, 2))).astype(np.float32)
    np.random.seed(1)
    got = rs.render_array(t, 48000, hall_type="Plate", air_absorption=0.0, want_stereo=True)
    np.random.seed(1)
    want = orc.render(t, 48000, hall="Plate", air=0.0)
    assert np.array_equal(got["stereo"] != 0, want["stereo"] != 0) or rel_err(got["stereo"], want["stereo"]) <= TOL
    assert rel_err(got["final"], want["final"]) <= TOL


def test_ir_resampling_matches_scipy(rs, tmp_path):
    """scipy.signal.resample (rs.py:1039) on the GPU: up / down, odd / even lengths, then the file-level path with
    an external IR at another sample rate."""
    from scipy.signal import resample
    from ars_b200 import wavio
    g = np.random.default_rng(55)
    for n, num in ((4800, 5292), (4801, 4410), (6000, 3000), (3000, 6000), (4410, 4800), (5001, 7777), (64, 65), (9, 4)):
        ir = (g.standard_normal((n, 2)) * np.exp(-np.arange(n) / (0.3 * n))[:, None]).astype(np.float32)
        want = resample(ir, num, axis=0)
        got = rs.resample_ir(ir, num)
        assert got.shape == want.shape and got.dtype == np.float32
        assert rel_err(got, want) <= TOL, (n, num, rel_err(got, want))
    x = (0.25 * g.standard_normal((30000, 2))).astype(np.float32)
    ir = (g.standard_normal((8000, 2)) * np.exp(-np.arange(8000) / 1500.0)[:, None]).astype(np.float32)
    ir /= np.max(np.abs(ir)) * 4
    wavio.write_float32(str(tmp_path / "x.wav"), x, 48000)
    wavio.write_float32(str(tmp_path / "ir.wav"), ir, 44100)
    p1, _, text = rs.apply_raytrace_convolution_3d(str(tmp_path / "x.wav"), str(tmp_path / "ir.wav"), True, "Room", 100, .5,
                                                   .1, .8, .6, .5, .5, 1.2, .9, .5, .5, .5, "Holz", "5.1 (Standard)")
    assert p1 is not None, text
    pcm, rate = wavio.read(p1)
    ir_rs = resample(ir, int(8000 * 48000 / 44100), axis=0).astype(np.float32)
    want = orc.render(x, 48000, external_ir=ir_rs, dry_wet_amount=.5, kill_start=.5, bass=1.2, treble=.9,
                      layout="5.1 (Standard)")
    d = np.abs(np.round(pcm * 32768).astype(np.int32) - want["pcm"].astype(np.int32))
    assert pcm.shape == want["pcm"].shape and d.max() <= 1


def test_sparse_ir_spectrum_route_matches_full_transform(rs, golden):
    """The IR spectrum of a procedural IR is taken through the overlap-save route (cached chirp delay line);
    it must agree with the two-M-point-transform route and with the goldens."""
    from ars_b200 import _capi
    g = golden("convolve")
    keys = [str(k) for k in g["x_keys"]]
    x = (0.3 * np.random.default_rng(3).standard_normal((50000, 2))).astype(np.float32)
    kw = dict(hall_type="Cathedral", room_size=600., air_absorption=.3, bass_gain=1.4, treble_gain=.8, dry_wet=.55,
              target_channel_layout="5.1 (Standard)")
    out = []
    try:
        for use in (1, 0):
            _capi.set_option("sparse_ir", use)
            for i, p in enumerate(g["split_par"]):
                el, ll, dw, b, t, ks, air = p[1:8]
                got = rs.convolve_audio_split_3d(g["x_" + keys[int(p[0])]], g["early"], g["late"], el, ll, dw, b, t, 48000,
                                                 ks, air)
                assert rel_err(got, g[f"split{i}"]) <= TOL, (use, i, rel_err(got, g[f"split{i}"]))
            np.random.seed(8)
            out.append(rs.render_array(x, 48000, want_stereo=True, **kw))
    finally:
        _capi.set_option("sparse_ir", 1)
    np.random.seed(8)
    want = orc.render(x, 48000, hall="Cathedral", room_size=600., air=.3, bass=1.4, treble=.8, dry_wet_amount=.55,
                      layout="5.1 (Standard)")
    for r in out:
        assert rel_err(r["stereo"], want["stereo"]) <= TOL and rel_err(r["final"], want["final"]) <= TOL
    assert rel_err(out[0]["final"], out[1]["final"]) <= 2e-6


def test_cfg2_full_size_vs_oracle(rs):
    """BASELINE configs[1] at full size: 60 s mono -> 5.1, Room / Holz / 200 m^3, air 0.1, bass 1.5 / treble 0.8;
    N = 2 970 503 is prime (the reference's FFTs fall back to Bluestein there, and so do ours)."""
    rate = 48000
    x = (0.3 * np.random.default_rng(0).standard_normal(60 * rate)).astype(np.float32)
    np.random.seed(11)
    want = orc.render(x, rate, hall="Room", room_size=200., diffusion=.5, air=.1, early=.8, late=.6, dry_wet_amount=.6,
                      kill_start=.5, bass=1.5, treble=.8, x=.3, y=.4, z=.6, material="Holz", layout="5.1 (Standard)")
    np.random.seed(11)
    got = rs.render_array(x, rate, hall_type="Room", room_size=200., diffusion=.5, air_absorption=.1, base_early_level=.8,
                          base_late_level=.6, dry_wet=.6, dry_wet_kill_start=.5, bass_gain=1.5, treble_gain=.8, x_pos=.3,
                          y_pos=.4, z_pos=.6, material="Holz", target_channel_layout="5.1 (Standard)")
    assert got["final"].shape == (2970503, 6)
    assert rel_err(got["final"], want["final"]) <= TOL
    assert snr_db(got["final"], want["final"]) >= 100.0
    d = np.abs(got["pcm"].astype(np.int32) - want["pcm"].astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) < 1e-2
    assert abs(got["metrics"]["lufs"] - want["metrics"]["lufs"]) <= 5e-3


def test_cfg3_full_size_properties(rs):
    """BASELINE configs[2] at full size (300 s, 8 s procedural IR, air 0.1, 5.1.2): N = 14 783 999, 2^25-point FFTs.
    Size-independent properties: linearity (every guard idle), and agreement of the two IR-spectrum routes."""
    from ars_b200 import _capi
    rate = 48000
    x = (0.02 * np.random.default_rng(2).standard_normal((300 * rate, 2), dtype=np.float32)).astype(np.float32)
    kw = dict(hall_type="Cathedral", room_size=20000., ir_duration=8.0, material="Stein", air_absorption=.1, dry_wet=.5,
              target_channel_layout="5.1.2 (Atmos Light)", want_metrics=False, want_pcm=False)
    np.random.seed(3)
    a = rs.render_array(x, rate, **kw)["final"]
    assert a.shape == (14783999, 8) and np.max(np.abs(a)) < 0.9
    np.random.seed(3)
    b = rs.render_array((2.0 * x).astype(np.float32), rate, **kw)["final"]
    assert rel_err(b, 2.0 * a) <= 2e-6
    del b
    try:
        _capi.set_option("sparse_ir", 0)
        np.random.seed(3)
        c = rs.render_array(x, rate, **kw)["final"]
    finally:
        _capi.set_option("sparse_ir", 1)
    assert rel_err(c, a) <= 2e-6
    del c
    # ... and of the two convolution routes: `a` took the folded-air overlap-save route, this one the N-point filter
    try:
        _capi.set_option("air_fold", 0)
        np.random.seed(3)
        e = rs.render_array(x, rate, **kw)["final"]
    finally:
        _capi.set_option("air_fold", 1)
    assert rel_err(e, a) <= 3e-6, rel_err(e, a)
    del e
    # the 18 ms height pair is the rear pair delayed and scaled by 0.6 * z (rs.py:549-554)
    assert np.array_equal(a[864:, 6], (a[:-864, 4].astype(np.float64) * (0.5 * 0.6)).astype(np.float32))


def test_channel_levels_match_numpy(rs):
    """rs.py:769-798: per-channel RMS dBFS and the side-signal RMS ("stereo width")."""
    g = np.random.default_rng(77)
    for n, ch in ((48000, 6), (12345, 2), (999, 1), (30000, 8)):
        d = (0.3 * g.standard_normal((n, ch)) * g.uniform(0.01, 1.0, ch)).astype(np.float32)
        levels, side = rs.channel_levels(d)
        for c in range(ch):
            want = 20 * np.log10(float(np.sqrt(np.mean(d[:, c] ** 2))))
            assert abs(levels[c] - want) <= 1e-4
        want_side = float(np.sqrt(np.mean(((d[:, 0] - d[:, 1]) * 0.5) ** 2))) if ch >= 2 else 0.0
        assert abs(side - want_side) <= 1e-6 * max(1.0, want_side)


def test_device_pointer_render_equals_host_render(rs):
    """ars_render_dev (device buffers, what bench.py times) must produce exactly what ars_render does."""
    import torch
    from ars_b200 import _capi
    lib = _capi.init()
    g = np.random.default_rng(5)
    rate = 48000
    x = (0.4 * g.standard_normal((70001, 2))).astype(np.float32)
    kw = dict(hall_type="Room", room_size=250., air_absorption=.2, bass_gain=1.3, treble_gain=.9, dry_wet=.5,
              target_channel_layout="7.1 (Surround)")
    np.random.seed(17)
    host = rs.render_array(x, rate, **kw)
    p, refl = rs.make_render_params(rate, want_lufs=True, **kw)
    np.random.seed(17)
    taps, bases, noise = rs.draw_ir_randoms(rate, p.ir_duration, refl, p.ir_max_delay, p.ir_split_time)
    N = int(lib.ars_render_out_len(p, x.shape[0], 0))
    d_x = torch.from_numpy(x).cuda()
    d_noise = torch.from_numpy(noise).cuda()
    d_pcm = torch.empty((N, 8), dtype=torch.int16, device="cuda")
    d_f32 = torch.empty((N, 8), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    keep = []
    draws = _capi.make_draws(taps, bases, int(d_noise.data_ptr()), keep)
    draws.noise_len = int(noise.size)
    m = _capi.ArsMetrics()
    _capi.check(lib.ars_render_dev(p, d_x.data_ptr(), x.shape[0], 2, None, 0, draws, None, d_f32.data_ptr(),
                                   d_pcm.data_ptr(), m), "ars_render_dev")
    _capi.check(lib.ars_sync(), "ars_sync")
    assert np.array_equal(d_pcm.cpu().numpy(), host["pcm"])
    assert np.array_equal(d_f32.cpu().numpy(), host["final"])
    assert rs._metrics_dict(m) == host["metrics"]


def test_async_device_renders_deliver_the_same_metrics_later(rs):
    """ars_render_dev_async enqueues and returns; the metrics of every render enqueued so far are filled in by the next
    ars_sync() -- same values as the synchronous call, clip after clip, PCM included."""
    import torch
    from ars_b200 import _capi
    lib = _capi.init()
    g = np.random.default_rng(9)
    rate = 48000
    kw = dict(hall_type="Room", room_size=180., air_absorption=.1, dry_wet=.4, target_channel_layout="5.1.2 (Atmos Light)")
    p, refl = rs.make_render_params(rate, want_lufs=True, **kw)
    clips, keep = [], []
    for k in range(5):
        x = ((0.2 + 0.2 * k) * g.standard_normal((40001 + 7000 * k, 2))).astype(np.float32)
        np.random.seed(100 + k)
        taps, bases, noise = rs.draw_ir_randoms(rate, p.ir_duration, refl, p.ir_max_delay, p.ir_split_time)
        d_x, d_noise = torch.from_numpy(x).cuda(), torch.from_numpy(noise).cuda()
        N = int(lib.ars_render_out_len(p, x.shape[0], 0))
        draws = _capi.make_draws(taps, bases, int(d_noise.data_ptr()), keep)
        draws.noise_len = int(noise.size)
        clips.append((x, d_x, d_noise, draws, N))
    torch.cuda.synchronize()
    want = []
    for x, d_x, d_noise, draws, N in clips:
        d_pcm = torch.empty((N, 8), dtype=torch.int16, device="cuda")
        m = _capi.ArsMetrics()
        _capi.check(lib.ars_render_dev(p, d_x.data_ptr(), x.shape[0], 2, None, 0, draws, None, None, d_pcm.data_ptr(), m), "dev")
        want.append((rs._metrics_dict(m), d_pcm.cpu().numpy()))
    later, pcms = [], []
    for x, d_x, d_noise, draws, N in clips:
        d_pcm = torch.empty((N, 8), dtype=torch.int16, device="cuda")
        m = _capi.ArsMetrics()
        _capi.check(lib.ars_render_dev_async(p, d_x.data_ptr(), x.shape[0], 2, None, 0, draws, None, None, d_pcm.data_ptr(), m), "async")
        later.append(m)
        pcms.append(d_pcm)
    _capi.check(lib.ars_sync(), "ars_sync")
    for (wm, wp), m, d_pcm in zip(want, later, pcms):
        assert rs._metrics_dict(m) == wm
        assert np.array_equal(d_pcm.cpu().numpy(), wp)
    assert len({wm["lufs"] for wm, _ in want}) == 5            # (five different clips, five different readings)


@pytest.mark.parametrize("case", ["procedural 5.1.2", "procedural 5.1.2 in four lanes", "external IR 7.1", "EQ on (exact-N route)"])
def test_head_start_of_async_renders_changes_nothing(rs, case):
    """Between two asynchronous renders of the same geometry the IR chain and the first passes of the second one run next
    to the tail of the first (ars_render_dev_async, option head_start).  Different clips and different random draws of
    one geometry, enqueued back to back: every PCM frame and every metric as from the synchronous calls, and the head
    start must actually have been taken (except on the exact-N route, which has no side chain and no lanes)."""
    import torch
    from ars_b200 import _capi
    lib = _capi.init()
    g = np.random.default_rng(13)
    rate = 48000
    n = 700001
    ext = None
    striped = case.endswith("lanes")
    if case.startswith("procedural"):
        kw = dict(hall_type="Cathedral", room_size=300., air_absorption=.1, dry_wet=.5, target_channel_layout="5.1.2 (Atmos Light)")
        if striped:
            n = 1500001                # (several transforms; one per stripe below, so the stripes go to the four lanes as in a long render)
    elif case.startswith("external"):
        kw = dict(dry_wet=.6, bass_gain=1.0, treble_gain=1.0, target_channel_layout="7.1 (Surround)")
        ext = (g.standard_normal((30000, 2)) * np.exp(-np.arange(30000) / 6000.0)[:, None] / 40).astype(np.float32)
    else:
        kw = dict(hall_type="Room", room_size=150., air_absorption=.2, bass_gain=1.4, treble_gain=.8, dry_wet=.5,
                  target_channel_layout="5.1 (Standard)")
        n = 90001
    p, refl = rs.make_render_params(rate, want_lufs=True, external_ir=ext is not None, **kw)
    C = 8 if "5.1 (" not in kw["target_channel_layout"] else 6
    d_ext = torch.from_numpy(ext).cuda() if ext is not None else None
    L = 0 if ext is None else ext.shape[0]
    clips, keep = [], []
    for k in range(6):
        x = ((0.1 + 0.25 * k) * g.standard_normal((n, 2))).astype(np.float32)     # (the later ones trip the stereo guard)
        d_x = torch.from_numpy(x).cuda()
        draws = None
        if ext is None:
            np.random.seed(300 + k)
            taps, bases, noise = rs.draw_ir_randoms(rate, p.ir_duration, refl, p.ir_max_delay, p.ir_split_time)
            d_noise = torch.from_numpy(noise).cuda()
            keep.append(d_noise)
            draws = _capi.make_draws(taps, bases, int(d_noise.data_ptr()), keep)
            draws.noise_len = int(noise.size)
        clips.append((d_x, draws))
    N = int(lib.ars_render_out_len(p, n, L))
    torch.cuda.synchronize()

    def run(fn, sync_each):
        out = []
        for d_x, draws in clips:
            d_pcm = torch.empty((N, C), dtype=torch.int16, device="cuda")
            m = _capi.ArsMetrics()
            _capi.check(fn(p, d_x.data_ptr(), n, 2, d_ext.data_ptr() if d_ext is not None else None, L, draws, None, None,
                           d_pcm.data_ptr(), m), "render")
            out.append((m, d_pcm))
        _capi.check(lib.ars_sync(), "ars_sync")
        return [(rs._metrics_dict(m), q.cpu().numpy()) for m, q in out]

    try:
        if striped:
            _capi.set_option("olsb_stripe", 1)
        want = run(lib.ars_render_dev, True)
        _capi.set_option("loud_stream", 0)
        _capi.set_option("tail_overlap", 0)
        h0, l0, t0 = int(lib.ars_head_start_count()), int(lib.ars_meter_stream_count()), int(lib.ars_tail_overlap_count())
        got = run(lib.ars_render_dev_async, False)
        taken = int(lib.ars_head_start_count()) - h0
        assert taken == len(clips) - 1, taken
        assert int(lib.ars_meter_stream_count()) == l0 and int(lib.ars_tail_overlap_count()) == t0     # (both off)
        # option loud_stream: the meters run on the meter stream, off the chain of last and final passes; with tail_overlap
        # the last passes wait for their slot's meter only, from the second or third render on
        _capi.set_option("loud_stream", 1)
        one_y = run(lib.ars_render_dev_async, False)
        metered = int(lib.ars_meter_stream_count()) - l0
        assert metered == len(clips), metered
        assert int(lib.ars_tail_overlap_count()) == t0
        _capi.set_option("tail_overlap", 1)
        two_y = run(lib.ars_render_dev_async, False)
        overlapped = int(lib.ars_tail_overlap_count()) - t0
        assert len(clips) - 2 <= overlapped <= len(clips) - 1, overlapped
        again = run(lib.ars_render_dev_async, False)          # (both slots warm now)
        _capi.set_option("head_start", 0)
        plain = run(lib.ars_render_dev_async, False)
        assert int(lib.ars_head_start_count()) - h0 == 4 * taken
        assert int(lib.ars_meter_stream_count()) - l0 == 4 * metered
    finally:
        _capi.set_option("head_start", 1)
        _capi.set_option("loud_stream", 1)          # (the defaults)
        _capi.set_option("tail_overlap", 0)
        _capi.set_option("olsb_stripe", 0)
    for k, (wm, wp) in enumerate(want):
        for other in (got, one_y, two_y, again, plain):
            om, op = other[k]
            assert om == wm
            assert np.array_equal(op, wp)
    assert len({wm["lufs"] for wm, _ in want}) == len(clips)


@pytest.mark.parametrize("layout", ["Stereo", "5.1 (Standard)", "7.1 (Surround)", "5.1.2 (Atmos Light)"])
@pytest.mark.parametrize("level", [0.05, 0.9, 3.0])
def test_final_pass_inside_the_meter_is_bit_identical_to_the_two_kernels(rs, layout, level):
    """final_with_loudness (metrics.cu): the final pass carried in the one-pass meter's feed must leave exactly what
    tail_final + the meter leave -- PCM, float frames, peak, RMS and LUFS bit for bit -- with every guard idle, with the
    stereo guard dividing, with the pan guard firing (general frame math), for the Stereo down-mix, with tiny and zero
    stretches, for a clip that ends inside a block and for one shorter than a gating block (falls back)."""
    from ars_b200 import _capi
    g = np.random.default_rng(78)
    rate = 48000
    ir = (g.standard_normal((3000, 2)) * np.exp(-np.arange(3000) / 700.0)[:, None]).astype(np.float32)
    ir[0] = 1.0
    for n in (200, 30011, 150001):
        x = (level * g.standard_normal((n, 2))).astype(np.float32)
        x[n // 3: n // 3 + 4000] *= np.float32(1e-30)
        x[n // 2: n // 2 + 4000] = 0.0
        kw = dict(external_ir_data=ir, dry_wet=.5, bass_gain=1.0, treble_gain=1.0, x_pos=.2, y_pos=.1, z_pos=.9,
                  target_channel_layout=layout)
        res = {}
        try:
            for mode in (0, 1):
                _capi.set_option("final_in_meter", mode)
                res[mode] = (rs.render_array(x, rate, **kw), rs.render_array(x, rate, want_float=False, **kw))
        finally:
            _capi.set_option("final_in_meter", 0)
        for k in (0, 1):
            assert np.array_equal(res[1][k]["pcm"], res[0][k]["pcm"]), (n, k)
            assert res[1][k]["metrics"] == res[0][k]["metrics"], (n, k, res[1][k]["metrics"], res[0][k]["metrics"])
        assert np.array_equal(res[1][0]["final"].view(np.uint32), res[0][0]["final"].view(np.uint32)), n
        if n > 48000 * 0.4:
            assert res[0][0]["metrics"]["lufs"] is not None


@pytest.mark.parametrize("layout", ["5.1 (Standard)", "7.1 (Surround)", "5.1.2 (Atmos Light)"])
@pytest.mark.parametrize("loud", [False, True])
def test_final_pass_lean_loop_is_bit_identical_to_the_general_loop(rs, layout, loud):
    """The lean frame loop of the final pass (packed guard division behind one range test, 32-bit offsets, loop split at
    the layout delay) must reproduce the general loop bit for bit: PCM, float frames (compared as bit patterns), peak,
    RMS and loudness -- with the stereo guard idle and dividing, with values below the division's plain range (their
    frames take the element-wise form), exact zeros, and a clip shorter than the layout delay."""
    from ars_b200 import _capi
    g = np.random.default_rng(77)
    rate = 48000
    ir = (g.standard_normal((3000, 2)) * np.exp(-np.arange(3000) / 700.0)[:, None]).astype(np.float32)
    ir[0] = 1.0
    for n in (200, 150001):
        x = ((0.9 if loud else 0.05) * g.standard_normal((n, 2))).astype(np.float32)
        x[n // 3: n // 3 + 4000] *= np.float32(1e-30)            # tiny but non-zero stage output
        x[n // 2: n // 2 + 4000] = 0.0                            # (the IR's ring-out keeps these frames non-zero)
        x[-3500:] = 0.0
        kw = dict(external_ir_data=ir, dry_wet=.5, bass_gain=1.0, treble_gain=1.0, x_pos=.2, y_pos=.4, z_pos=.8,
                  target_channel_layout=layout)
        res, pcm_only = {}, {}
        try:
            for mode in (0, 1, 2):
                _capi.set_option("final_lean", mode)
                res[mode] = rs.render_array(x, rate, **kw)
                pcm_only[mode] = rs.render_array(x, rate, want_float=False, **kw)      # (its own instantiation of the loop)
        finally:
            _capi.set_option("final_lean", 2)
        for mode in (1, 2):
            assert np.array_equal(res[mode]["pcm"], res[0]["pcm"]), (n, mode)
            assert np.array_equal(res[mode]["final"].view(np.uint32), res[0]["final"].view(np.uint32)), (n, mode)
            assert res[mode]["metrics"] == res[0]["metrics"], (n, mode, res[mode]["metrics"], res[0]["metrics"])
            assert np.array_equal(pcm_only[mode]["pcm"], res[0]["pcm"]), (n, mode)
            assert pcm_only[mode]["metrics"] == res[0]["metrics"], (n, mode)
        if loud and n > 1000:
            assert np.max(np.abs(res[0]["final"])) <= 1.0 and res[0]["metrics"]["true_peak_dbfs"] > -12.0


def test_large_transform_2pow28_properties(rs):
    """A 25-minute clip with the EQ mask on: N = 72 000 000 + L - 1, Bluestein length 2^28 (three passes of 2^8 /
    2^8 / 2^12 points, two-level twiddle tables, 64-bit indices).  Linearity plus a direct time-domain spot check of
    the convolution at a handful of output frames."""
    rate = 48000
    n = 25 * 60 * rate
    g = np.random.default_rng(9)
    x = (0.05 * g.standard_normal((n, 2), dtype=np.float32)).astype(np.float32)
    ir = (g.standard_normal((4000, 2)) * np.exp(-np.arange(4000) / 800.0)[:, None]).astype(np.float32)
    ir /= np.max(np.abs(ir)) * 10
    kw = dict(external_ir_data=ir, dry_wet=.5, dry_wet_kill_start=1.0, bass_gain=1.0, treble_gain=1.00002,
              target_channel_layout="Stereo", want_metrics=False, want_pcm=False, want_stereo=True)
    a = rs.render_array(x, rate, **kw)                      # treble 1.00002 is outside np.isclose -> N-point path
    assert a["stereo"].shape == (n + 3999, 2) and np.max(np.abs(a["stereo"])) < 1.0
    # spot check: with gains this close to 1 the EQ changes the result by < 2e-5 * |y|, far below the tolerance used
    for f in (0, 1234, 3999, n // 2 + 17, n - 1, n + 3000):
        lo = max(0, f - 3999)
        wl = sum(float(ir[f - m, 0]) * float(x[m, 0]) for m in range(lo, min(n, f + 1)))
        wr = sum(float(ir[f - m, 1]) * float(x[m, 1]) for m in range(lo, min(n, f + 1)))
        dl = float(x[f, 0]) if f < n else 0.0
        dr = float(x[f, 1]) if f < n else 0.0
        assert abs(a["stereo"][f, 0] - (0.5 * dl + 0.5 * wl)) <= 3e-5
        assert abs(a["stereo"][f, 1] - (0.5 * dr + 0.5 * wr)) <= 3e-5
    b = rs.render_array((2.0 * x).astype(np.float32), rate, **kw)
    assert rel_err(b["stereo"], 2.0 * a["stereo"]) <= 2e-6


def test_concurrent_callers_are_serialised(rs):
    """The reference is called from Gradio worker threads: concurrent calls into the library must
    give the same results as sequential ones (one mutex, one stream)."""
    import threading
    g = np.random.default_rng(66)
    ir = (g.standard_normal((3000, 2)) * np.exp(-np.arange(3000) / 700.0)[:, None]).astype(np.float32)
    ir /= np.max(np.abs(ir)) * 5
    jobs = [((0.3 * g.standard_normal((20000 + 1111 * i, 2))).astype(np.float32),
             dict(external_ir_data=ir, dry_wet=.3 + .1 * i, bass_gain=1.0 + .1 * (i % 3), treble_gain=1.0,
                  target_channel_layout=["Stereo", "5.1 (Standard)", "7.1 (Surround)", "5.1.2 (Atmos Light)"][i % 4]))
            for i in range(8)]
    want = [rs.render_array(x, 48000, **kw) for x, kw in jobs]
    got = [None] * len(jobs)

    def work(i):
        x, kw = jobs[i]
        got[i] = rs.render_array(x, 48000, **kw)
        rs.apply_surround_panning_3d(x, .3, .4, .5)          # interleave other entry points too
        rs.calculate_audio_metrics(x, 48000)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for a, b in zip(want, got):
        assert np.array_equal(a["final"], b["final"]) and np.array_equal(a["pcm"], b["pcm"])
        assert a["metrics"] == b["metrics"]


def test_batch_with_many_distinct_lengths_recycles_plans_correctly(rs):
    """More distinct output lengths than the plan cache holds (16): evicted plans hand their buffers to new ones.
    Every clip must still equal its single-call render bit for bit."""
    g = np.random.default_rng(91)
    rate = 16000
    halls = ["Plate", "Room", "Cathedral"]
    jobs = []
    for i in range(22):
        n = int(g.integers(4000, 40000))
        jobs.append(dict(samples=(0.3 * g.standard_normal((n, 2))).astype(np.float32), rate=rate, seed=300 + i,
                         hall_type=halls[i % 3], room_size=float(10 * g.integers(1, 101)),
                         air_absorption=float(g.uniform()), bass_gain=float(g.uniform(.5, 2)), dry_wet=float(g.uniform()),
                         target_channel_layout=["Stereo", "7.1 (Surround)", "5.1.2 (Atmos Light)"][i % 3]))
    for rep in range(2):                      # second round: every plan of round one has been evicted in between
        got = rs.render_batch(jobs, want_float=True)
        for job, b in zip(jobs, got):
            job = dict(job)
            np.random.seed(job.pop("seed"))
            a = rs.render_array(job.pop("samples"), job.pop("rate"), **job)
            assert np.array_equal(a["final"], b["final"]) and np.array_equal(a["pcm"], b["pcm"])
            assert a["metrics"] == b["metrics"]


# ------------------------------------------------------------------ air absorption folded into the IR ----------
def _fold_opts(**kw):
    from ars_b200 import _capi
    for k, v in kw.items():
        _capi.set_option(k, v)


def _fold_defaults():
    _fold_opts(air_fold=1, air_fold_eps_e9=2000, air_fold_max_taps=131072, mac_tiled_min=4, ols_r2=0)


def _fold_count():
    from ars_b200 import _capi
    return int(_capi.init().ars_air_fold_count())


def test_air_fold_stage_matches_exact_path_and_oracle(rs):
    """A render whose only spectral mask is the air ramp (air > 0.01, EQ flat) folds the ramp into the IR and runs as
    one overlap-save convolution (upols.cuh).  Against the exact N-point route and against the oracle, both MAC forms."""
    rate = 48000
    rng = np.random.default_rng(21)
    np.random.seed(5)
    early, late = rs.generate_impulse_response_split_3d(rate, 1.5, 30, 0.06, "Holz", 0.5, 0.08, 0.5)
    cases = [((rng.standard_normal((70001, 2)) * 0.3).astype(np.float32), 0.1),
             ((rng.standard_normal(40000) * 0.3).astype(np.float32), 0.15),
             ((rng.standard_normal((52000, 3)) * 0.3).astype(np.float32), 0.05)]
    try:
        for x, air in cases:
            want = orc.convolve_split(x, early, late, 0.7, 0.9, 0.6, 1.0, 1.0, rate, 0.5, air)
            outs = {}
            for name, opts in (("exact", dict(air_fold=0)), ("fold", dict(air_fold=1, mac_tiled_min=4)),
                               ("fold_prologue", dict(air_fold=1, mac_tiled_min=1000)),
                               ("fold_r2", dict(air_fold=1, mac_tiled_min=4, ols_r2=1)),
                               ("fold_r2_prologue", dict(air_fold=1, mac_tiled_min=1000, ols_r2=1))):
                _fold_opts(**opts)
                c0 = _fold_count()
                outs[name] = rs.convolve_audio_split_3d(x, early, late, 0.7, 0.9, 0.6, 1.0, 1.0, rate, 0.5, air)
                assert (_fold_count() - c0 == 1) == (name != "exact"), name
                assert rel_err(outs[name], want) <= TOL, (name, air, rel_err(outs[name], want))
                assert snr_db(outs[name], want) >= 100.0, (name, air, snr_db(outs[name], want))
            assert rel_err(outs["fold"], outs["exact"]) <= 3e-6, rel_err(outs["fold"], outs["exact"])
            assert rel_err(outs["fold_prologue"], outs["fold"]) <= 1e-6
            assert rel_err(outs["fold_r2"], outs["fold"]) <= 1e-6 and rel_err(outs["fold_r2_prologue"], outs["fold"]) <= 1e-6
    finally:
        _fold_defaults()


def test_air_fold_wraps_around_the_period_like_the_n_point_filter(rs):
    """The reference's ramp acts on the N-point rfft, i.e. circularly: with an IR much shorter than the kept air
    kernel and energy at both ends of the signal the pre- and post-ring wrap around.  Pure tones at the two kinks of
    the ramp (2 kHz, Nyquist) are the worst case of the truncation bound."""
    rate = 48000
    n = 90000
    t = np.arange(n)
    late = np.zeros(600, np.float32)
    late[100:600] = (np.random.default_rng(2).standard_normal(500) * np.exp(-np.arange(500) / 80.0)).astype(np.float32)
    late *= 0.7 / np.max(np.abs(late))
    early = np.zeros(600, np.float32)
    early[[3, 40, 77]] = [0.9, -0.4, 0.2]
    sigs = {"noise": (np.random.default_rng(3).standard_normal((n, 2)) * 0.3).astype(np.float32),
            "tone2k": np.stack([0.5 * np.sin(2 * np.pi * 2000.0 * t / rate), 0.5 * np.cos(np.pi * t)], 1).astype(np.float32)}
    try:
        for name, x in sigs.items():
            want = orc.convolve_split(x, early, late, 0.8, 1.0, 0.7, 1.0, 1.0, rate, 0.5, 0.12)
            _fold_opts(air_fold=1)
            c0 = _fold_count()
            got = rs.convolve_audio_split_3d(x, early, late, 0.8, 1.0, 0.7, 1.0, 1.0, rate, 0.5, 0.12)
            assert _fold_count() == c0 + 1
            assert rel_err(got, want) <= TOL, (name, rel_err(got, want))
    finally:
        _fold_defaults()


def test_air_fold_error_bound_option_and_fallback(rs):
    """air_fold_eps_e9 / air_fold_max_taps pick the kept kernel length; a ramp too deep for the bound falls back to the
    exact N-point filter."""
    rate = 48000
    x = (np.random.default_rng(8).standard_normal((160000, 2)) * 0.3).astype(np.float32)
    np.random.seed(6)
    early, late = rs.generate_impulse_response_split_3d(rate, 1.0, 20, 0.05, "Teppich", 0.4, 0.06, 0.3)
    try:
        want = orc.convolve_split(x, early, late, 0.8, 0.6, 0.5, 1.0, 1.0, rate, 0.5, 0.6)
        _fold_defaults()
        c0 = _fold_count()
        got = rs.convolve_audio_split_3d(x, early, late, 0.8, 0.6, 0.5, 1.0, 1.0, rate, 0.5, 0.6)
        assert _fold_count() == c0, "air 0.6 needs 53 k taps at 2e-6, too much fold work for a 4 s clip: must take the exact route"
        assert rel_err(got, want) <= TOL
        _fold_opts(air_fold_eps_e9=4000, air_fold_max_taps=65536)
        got = rs.convolve_audio_split_3d(x, early, late, 0.8, 0.6, 0.5, 1.0, 1.0, rate, 0.5, 0.6)
        assert _fold_count() == c0 + 1
        assert rel_err(got, want) <= TOL, rel_err(got, want)
        # beyond 32768 taps the fold is taken when the clip is long enough for it to pay (late span x K <= 128 N)
        _fold_defaults()
        xl = (np.random.default_rng(9).standard_normal((80 * rate, 2)) * 0.3).astype(np.float32)
        want_l = orc.convolve_split(xl, early, late, 0.8, 0.6, 0.5, 1.0, 1.0, rate, 0.5, 0.5)
        c1 = _fold_count()
        got_l = rs.convolve_audio_split_3d(xl, early, late, 0.8, 0.6, 0.5, 1.0, 1.0, rate, 0.5, 0.5)
        assert _fold_count() == c1 + 1, "air 0.5 (44 k taps) on an 80 s clip folds"
        assert rel_err(got_l, want_l) <= TOL, rel_err(got_l, want_l)
        assert snr_db(got_l, want_l) >= 100.0
        del xl, want_l, got_l
        c0 = _fold_count() - 1
        # EQ on: the brick-wall mask cannot be folded
        got = rs.convolve_audio_split_3d(x, early, late, 0.8, 0.6, 0.5, 1.3, 0.8, rate, 0.5, 0.1)
        assert _fold_count() == c0 + 1
    finally:
        _fold_defaults()


def test_air_fold_whole_render_vs_oracle(rs):
    """cfg3-like render (Cathedral / Stein, 8 s IR, air 0.1, EQ flat, 5.1.2, metrics + PCM), 20 s clip."""
    rate = 48000
    x = (0.2 * np.random.default_rng(2).standard_normal((20 * rate, 6))).astype(np.float32)
    np.random.seed(3)
    want = orc.render(x, rate, hall="Cathedral", room_size=20000., ir_duration=8.0, material="Stein", air=.1,
                      dry_wet_amount=.5, layout="5.1.2 (Atmos Light)")
    c0 = _fold_count()
    np.random.seed(3)
    got = rs.render_array(x, rate, hall_type="Cathedral", room_size=20000., ir_duration=8.0, material="Stein",
                          air_absorption=.1, dry_wet=.5, target_channel_layout="5.1.2 (Atmos Light)")
    assert _fold_count() == c0 + 1
    assert rel_err(got["final"], want["final"]) <= TOL
    assert snr_db(got["final"], want["final"]) >= 100.0
    d = np.abs(got["pcm"].astype(np.int32) - want["pcm"].astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) < 1e-2
    assert abs(got["metrics"]["lufs"] - want["metrics"]["lufs"]) <= 5e-3
    assert abs(got["metrics"]["rms_dbfs"] - want["metrics"]["rms_dbfs"]) <= 1e-3


def test_preset_manifest_batch_equals_single_renders(rs, tmp_path):
    """The reference's preset JSON as a batch-job description (ars_b200/presets.py): a manifest of WAV + preset pairs
    rendered in one pipelined batch writes the same PCM_16 payload as rendering each clip on its own."""
    from ars_b200 import presets, wavio
    rate = 48000
    g = np.random.default_rng(31)
    a = (0.3 * g.standard_normal((30000, 2))).astype(np.float32)
    b = (0.3 * g.standard_normal(41000)).astype(np.float32)
    ir = (g.standard_normal((6000, 2)) * np.exp(-np.arange(6000) / 900.0)[:, None]).astype(np.float32)
    ir /= np.max(np.abs(ir))
    wavio.write_float32(str(tmp_path / "a.wav"), a, rate)
    wavio.write_float32(str(tmp_path / "b.wav"), b, rate)
    wavio.write_float32(str(tmp_path / "ir.wav"), ir, rate)
    p1 = {"hall_type": "Cathedral", "material": "Stein", "room_size": 900, "air_absorption": 0.1, "dry_wet": 0.6,
          "target_layout": "5.1.2 (Atmos Light)", "z_pos": 0.8}
    presets.save_preset(str(tmp_path / "dom_v4.json"), presets.settings_to_preset(presets.preset_to_settings(p1)[0]))
    manifest = {"jobs": [{"audio": "a.wav", "preset": "dom_v4.json", "seed": 5, "out": "a_out.wav"},
                         {"audio": "b.wav", "preset": {"hall_type": "Plate", "bass_gain": 1.4, "target_layout": "Stereo"},
                          "seed": 6, "out": "b_out.wav"},
                         {"audio": "a.wav", "preset": {"use_external_ir": True, "dry_wet": 0.4}, "external_ir": "ir.wav"}]}
    (tmp_path / "jobs.json").write_text(__import__("json").dumps(manifest))
    res = presets.render_manifest(str(tmp_path / "jobs.json"))
    assert [r["pcm"].shape[1] for r in res] == [8, 2, 6] and res[2]["out"] is None
    np.random.seed(5)
    one = rs.render_array(a, rate, **presets.preset_to_settings(p1)[0])
    assert np.array_equal(one["pcm"], res[0]["pcm"])
    back, r2 = wavio.read(str(tmp_path / "a_out.wav"))
    assert r2 == rate and np.array_equal((back * 32768.0).astype(np.int16), one["pcm"])
    np.random.seed(6)
    two = rs.render_array(b, rate, hall_type="Plate", bass_gain=1.4, target_channel_layout="Stereo")
    assert np.array_equal(two["pcm"], res[1]["pcm"])
    three = rs.render_array(a, rate, external_ir_data=ir, dry_wet=0.4)
    assert np.array_equal(three["pcm"], res[2]["pcm"])
    assert abs(three["metrics"]["lufs"] - res[2]["metrics"]["lufs"]) <= 1e-9


def test_spectrogram_matches_scipy(rs):
    """Numerics of the visualiser's spectrogram (rs.py:621-634) against scipy.signal.spectrogram with the reference's
    arguments: segment length by duration, 50 % overlap, Hann, constant detrend, one-sided density."""
    from scipy.signal import spectrogram
    g = np.random.default_rng(41)
    for seconds, ch, rate in ((2.0, 1, 48000), (7.5, 2, 44100), (31.0, 6, 48000)):
        n = int(seconds * rate)
        t = np.arange(n) / rate
        x = (0.2 * g.standard_normal((n, ch)) + 0.5 * np.sin(2 * np.pi * 1000.0 * t)[:, None] + 0.1).astype(np.float32)
        nperseg = 4096 if seconds > 30 else 2048 if seconds > 5 else 1024
        f0, t0, s0 = spectrogram(x[:, 0].astype(np.float64), fs=rate, nperseg=nperseg, noverlap=nperseg // 2, window="hann")
        f1, t1, s1 = rs.spectrogram(x if ch > 1 else x[:, 0], rate)
        assert s1.shape == s0.shape and s1.dtype == np.float32
        assert np.allclose(f1, f0) and np.allclose(t1, t0)
        assert np.max(np.abs(s1 - s0)) <= 2e-5 * np.max(s0)
        db0, db1 = 10 * np.log10(np.maximum(s0, 1e-10)), 10 * np.log10(np.maximum(s1, 1e-10))
        assert abs(np.median(db1) - np.median(db0)) <= 1e-2 and abs(db1.max() - db0.max()) <= 1e-3
    with pytest.raises(ValueError):
        rs.spectrogram(np.zeros(700, np.float32), 48000)


def _true_peak_4x_numpy(x):
    """BS.1770-4 Annex 2 interpolator on the CPU: 48-tap FIR as four 12-tap phases, peak over phases, samples (tail
    included) and channels, in dBTP."""
    ph0 = [0.0017089843750, 0.0109863281250, -0.0196533203125, 0.0332031250000, -0.0594482421875, 0.1373291015625,
           0.9721679687500, -0.1022949218750, 0.0476074218750, -0.0266113281250, 0.0148925781250, -0.0083007812500]
    ph1 = [-0.0291748046875, 0.0292968750000, -0.0517578125000, 0.0891113281250, -0.1665039062500, 0.4650878906250,
           0.7797851562500, -0.2003173828125, 0.1015625000000, -0.0582275390625, 0.0330810546875, -0.0189208984375]
    phases = [ph0, ph1, ph1[::-1], ph0[::-1]]
    x = np.asarray(x, np.float64)
    if x.ndim == 1:
        x = x[:, None]
    pk = 0.0
    for c in range(x.shape[1]):
        for h in phases:
            pk = max(pk, float(np.max(np.abs(np.convolve(x[:, c], h)))))
    return 20 * np.log10(pk) if pk > 1e-15 else -np.inf


def test_true_peak_4x_add_on(rs):
    """The 4x-oversampled true peak (north star, subsystem 4; the reference itself reports the sample peak): array entry
    point and the render's figure against the same interpolator in numpy on the render's own float output; the classic
    fs/4 tone at 45 degrees reads 3 dB above its sample peak."""
    rate = 48000
    t = np.arange(rate)
    tone = (0.5 * np.sin(2 * np.pi * (rate / 4) * t / rate + np.pi / 4)).astype(np.float32)
    got = rs.true_peak_4x(tone)
    assert abs(got - _true_peak_4x_numpy(tone)) <= 1e-4
    sample_peak = 20 * np.log10(np.max(np.abs(tone)))
    assert 2.5 <= got - sample_peak <= 3.1, (got, sample_peak)
    g = np.random.default_rng(77)
    x = (0.4 * g.standard_normal((3 * rate + 17, 5))).astype(np.float32)
    assert abs(rs.true_peak_4x(x) - _true_peak_4x_numpy(x)) <= 1e-4
    clip = (0.3 * g.standard_normal((6 * rate, 2))).astype(np.float32)
    for layout in ("5.1.2 (Atmos Light)", "7.1 (Surround)", "Stereo", "5.1 (Standard)"):
        for amp in (1.0, 4.0):                       # amp 4: the peak guards fire
            np.random.seed(8)
            res = rs.render_array((amp * clip).astype(np.float32), rate, hall_type="Room", room_size=150., air_absorption=0.0,
                                  x_pos=.3, y_pos=.7, z_pos=.8, target_channel_layout=layout, want_true_peak_4x=True)
            want = _true_peak_4x_numpy(res["final"])
            assert abs(res["metrics"]["true_peak_4x_dbfs"] - want) <= 2e-3, (layout, amp, res["metrics"], want)
            # (phase 0 of the standard's interpolator is not the identity -- its centre tap is 0.972 -- so on noise the
            # oversampled figure may read a few tenths of a dB under the sample peak)
            assert res["metrics"]["true_peak_4x_dbfs"] >= res["metrics"]["true_peak_dbfs"] - 0.5
    np.random.seed(8)
    plain = rs.render_array(clip, rate, hall_type="Room")
    assert "true_peak_4x_dbfs" not in plain["metrics"]


def test_file_level_entry_point_never_raises(rs, tmp_path, monkeypatch):
    """rs.py:1096-1109: apply_raytrace_convolution_3d reports every failure as (None, None, message).  A failure of the GPU
    library itself (here: an invalid layout id forced through the C ABI) must come back the same way; the stage-level
    functions keep raising ArsError (there is no CPU path to fall back to)."""
    from ars_b200 import _capi, wavio
    rate = 48000
    x = (0.1 * np.random.default_rng(1).standard_normal((rate, 2))).astype(np.float32)
    wav = str(tmp_path / "in.wav")
    wavio.write_pcm16(wav, rs.float_to_pcm16(x), rate)
    args = (wav, None, False, "Room", 100.0, 0.5, 0.1, 0.8, 0.6, 0.5, 0.5, 1.0, 1.0, 0.5, 0.5, 0.5, "Holz", "5.1 (Standard)")
    ok = rs.apply_raytrace_convolution_3d(*args)
    assert ok[0] is not None and ok[2].startswith("LUFS:")
    monkeypatch.setitem(_capi.LAYOUT_IDS, "5.1 (Standard)", 99)
    bad = rs.apply_raytrace_convolution_3d(*args)
    assert bad[0] is None and bad[1] is None and isinstance(bad[2], str) and "Fehler" in bad[2], bad
    with pytest.raises(_capi.ArsError):
        rs.map_channels(np.zeros((10, 6), np.float32), "5.1 (Standard)", rate)
    # a write failure removes its temporary file (rs.py:1088-1091)
    monkeypatch.undo()
    import glob
    import tempfile
    before = set(glob.glob(tempfile.gettempdir() + "/processed_*.wav"))
    monkeypatch.setattr(wavio, "write_pcm16", lambda *a, **k: (_ for _ in ()).throw(OSError("disk full")))
    res = rs.apply_raytrace_convolution_3d(*args)
    assert res[0] is None and "Schreiben" in res[2]
    assert set(glob.glob(tempfile.gettempdir() + "/processed_*.wav")) == before


def test_metrics_of_integer_input_keep_peak_and_rms(rs):
    """rs.py:685-698: pyloudnorm refuses non-floating data inside the reference's inner try -- lufs None, peak and RMS
    still reported."""
    x = (1000 * np.random.default_rng(3).standard_normal((48000, 2))).astype(np.int16)
    m = rs.calculate_audio_metrics(x, 48000)
    assert m["lufs"] is None
    ref = orc.metrics(x.astype(np.float32), 48000, with_lufs=False)
    assert abs(m["true_peak_dbfs"] - ref["true_peak_dbfs"]) < 1e-4 and abs(m["rms_dbfs"] - ref["rms_dbfs"]) < 1e-4
