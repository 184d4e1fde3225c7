"""Pins oracle/ars_oracle.py against golden vectors produced by the unmodified
reference (oracle/make_golden.py).  Bit-exact unless a tolerance is stated."""
import numpy as np
import pytest

import ars_oracle as O

HALLS = ["Plate", "Room", "Cathedral", "Garage"]
MATERIALS = ["Stein", "Holz", "Teppich", "Glas", "Beton", "Vorhang (schwer)", "Gummi"]
LAYOUTS = ["Stereo", "5.1 (Standard)", "7.1 (Surround)", "5.1.2 (Atmos Light)", "9.1.6"]


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b, equal_nan=True), float(np.max(np.abs(a.astype(np.float64) - b)))


def test_scalar_prologue(golden):
    t = golden("scalars")["table"]
    for row in t:
        hall = HALLS[int(row[0])]
        room, x, y, z, dif, dw, e, l = row[1:9]
        dur, refl, mdel, split = O.shape_params(hall, room, z)
        assert (dur, refl, mdel, split) == (row[9], row[10], row[11], row[12])
        assert O.directionality(x, y, z, hall, dif, dw) == row[13]
        assert O.adapt_levels(dw, e, l) == (row[14], row[15])


def test_ir_synthesis_bit_exact(golden):
    g = golden("ir")
    for i, m in enumerate(g["meta"]):
        rate, hall, mat = int(m[0]), HALLS[int(m[1])], MATERIALS[int(m[2])]
        dif, seed = m[7], int(m[9])
        dur, refl, mdel, split, direc = m[10], int(m[11]), m[12], m[13], m[14]
        np.random.seed(seed)
        e, l = O.generate_ir(rate, dur, refl, mdel, mat, direc, split, dif)
        same(e, g[f"early{i}"])
        same(l, g[f"late{i}"])
    e, l = O.generate_ir(0, 1.0, 10, 0.05, "Holz", 0.5, 0.05, 0.5)
    same(e, g["early_bad"])
    same(l, g["late_bad"])


def test_air_filter(golden):
    g = golden("air_filter")
    for tag in ("odd", "even", "short"):
        sig = g[f"in_{tag}"]
        for air in (0.005, 0.1, 0.65, 1.7):
            same(O.air_filter(sig, 48000, air), g[f"air_{tag}_{air}"])
        same(O.air_filter(sig, 3000, 0.5), g[f"air_{tag}_lowrate"])


def test_dry_wet(golden):
    g = golden("dry_wet")
    for i in range(8):
        dw, ks = g[f"par{i}"]
        same(O.dry_wet(g[f"dry{i}"], g[f"wet{i}"], dw, ks), g[f"mix{i}"])


def test_convolve_split_and_external(golden):
    g = golden("convolve")
    keys = [str(k) for k in g["x_keys"]]
    for i, p in enumerate(g["split_par"]):
        x = g["x_" + keys[int(p[0])]]
        el, ll = (np.float64(p[1]), np.float64(p[2])) if p[8] else (float(p[1]), float(p[2]))
        y = O.convolve_split(x, g["early"], g["late"], el, ll, p[3], p[4], p[5], 48000, p[6], p[7])
        same(y, g[f"split{i}"])
    for i, p in enumerate(g["ext_par"]):
        x = g["x_" + keys[int(p[0])]]
        same(O.convolve_external(x, g["ext_ir"], p[1], p[2], p[3], 48000, p[4]), g[f"ext{i}"])
    same(O.convolve_external(g["x_stereo"], g["ext_ir"][:, 0], .5, 1.0, 1.0, 48000, .5), g["ext_badir"])


def test_pan_and_map(golden):
    g = golden("pan_map")
    keys = [str(k) for k in g["sig_keys"]]
    for idx, p in enumerate(g["par"]):
        s = g["s_" + keys[int(p[0])]]
        six = O.pan_5_1(s, p[1], p[2], p[3])
        same(six, g[f"pan{idx}"])
        for li, lay in enumerate(LAYOUTS):
            m, names = O.map_layout(six.copy(), lay, 48000, p[3])
            same(m, g[f"map{idx}_{li}"])
    six = O.pan_5_1(g["s_norm"], .4, .6, .8)
    same(six, g["pan_441"])
    same(O.map_layout(six.copy(), "7.1 (Surround)", 44100, .8)[0], g["map_441_71"])
    same(O.map_layout(six.copy(), "5.1.2 (Atmos Light)", 44100, .8)[0], g["map_441_512"])


def test_map_5_1_returns_same_object():
    six = np.full((10, 6), 2.0, np.float32)
    out, names = O.map_layout(six, "5.1 (Standard)", 48000)
    assert out is six and float(six.max()) == 1.0          # rs.py:538,559: scaled in place
    assert names == ["FL", "FR", "C", "LFE", "RL", "RR"]


def test_peak_rms(golden):
    g = golden("metrics_pipeline")
    for i in range(4):
        m = O.metrics(g[f"m_in{i}"], 48000, with_lufs=False)
        assert [m["true_peak_dbfs"], m["rms_dbfs"]] == list(g[f"m_out{i}"])


def test_full_pipeline_function(golden):
    """Reference `apply_raytrace_convolution_3d` (soundfile faked) vs oracle render():
    the float array handed to sf.write must match bit for bit, and so must the metrics text."""
    g = golden("metrics_pipeline")
    for i, p in enumerate(g["pipe_par"]):
        rate, ext = int(p[2]), bool(p[3])
        hall, mat, lay, seed = HALLS[int(p[4])], MATERIALS[int(p[17])], LAYOUTS[int(p[18])], int(p[19])
        ir = g["pipe_ir"] if ext else None
        np.random.seed(seed)
        r = O.render(g[f"pipe_in{i}"], rate, external_ir=ir, hall=hall, room_size=p[5], diffusion=p[6],
                     air=p[7], early=p[8], late=p[9], dry_wet_amount=p[10], kill_start=p[11], bass=p[12],
                     treble=p[13], x=p[14], y=p[15], z=p[16], material=mat, layout=lay)
        clipped = np.clip(r["final"], -0.9999, 0.9999)
        same(clipped, g[f"pipe_out{i}"])
        m = r["metrics"]
        lufs = f"{m['lufs']:.2f}" if m["lufs"] is not None and not np.isinf(m["lufs"]) else "N/A"
        text = f"LUFS: {lufs} | Peak: {m['true_peak_dbfs']:.1f} dBFS | RMS: {m['rms_dbfs']:.1f} dBFS"
        assert text == str(g[f"pipe_text{i}"])


def test_pcm16_rule():
    x = np.array([[0.0, 1.5, -1.5, np.nan, 0.5 / 32767, 1.5 / 32767, 2.5 / 32767, -0.9999]], np.float32)
    out = O.pcm16(x)
    assert out.dtype == np.int16
    assert out.tolist() == [[0, 32764, -32764, 0, 0, 2, 2, -32764]]     # half-to-even, clip at 0.9999


def test_loudness_known_answers():
    """Sanity pins for the (unpinned) BS.1770 restatement: a 997 Hz sine at -20 dBFS peak,
    mono, reads about -23.0 LUFS; silence gates to -inf; < 400 ms raises."""
    rate = 48000
    t = np.arange(rate * 5) / rate
    x = (0.1 * np.sin(2 * np.pi * 997 * t)).astype(np.float32)
    assert abs(O.integrated_loudness(x, rate) - (-23.0)) < 0.1
    with pytest.raises(ValueError):
        O.integrated_loudness(x[:1000], rate)
    blocks = O.loudness_blocks(rate * 5, rate)
    assert blocks[0] == (0, 19200) and len(blocks) == 47


def test_loudness_restatement_agrees_with_torchaudio():
    """pyloudnorm is not installable here (PARITY UNPINNED for LUFS).  Secondary cross-check: torchaudio's
    independent implementation of ITU-R BS.1770-4 (functional.loudness: RBJ high-shelf + 38 Hz high-pass, 400 ms / 75 %
    blocks, -70 LUFS absolute and -10 LU relative gates) against the oracle's restatement of pyloudnorm, on stationary and
    on level-switching noise and on a tone.  torchaudio filters in float32, hence the 1e-3 LU bar."""
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    rate = 48000
    g = np.random.default_rng(0)
    sigs = []
    for secs, amp in ((3.0, 0.1), (10.0, 0.05), (1.0, 0.08)):        # (torchaudio clamps its filters' output to +-1: keep clear)
        x = (amp * g.standard_normal(int(secs * rate))).astype(np.float32)
        sigs.append(x.copy())
        x[len(x) // 2:] *= 0.2                      # a level drop exercises the relative gate
        sigs.append(x)
    t = np.arange(4 * rate) / rate
    sigs.append((0.25 * np.sin(2 * np.pi * 997.0 * t)).astype(np.float32))
    for x in sigs:
        ours = O.integrated_loudness(x, rate)
        theirs = float(ta.functional.loudness(torch.from_numpy(x)[None, :], rate))
        assert abs(ours - theirs) <= 1e-3, (ours, theirs)
