"""Generate tests/golden/*.npz from the UNMODIFIED reference.

Runs only in the authoring container (needs /root/reference).  It imports
`/root/reference/raytracer_studio.py` with the six absent third-party modules
stubbed, feeds seeded inputs through the reference's own
functions and stores inputs + outputs.  The committed .npz files are what
travels; nothing at test / smoke / bench time reads /root/reference.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import tempfile
import unittest.mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/raytracer_studio.py"


def load_reference():
    """exec the reference module under MagicMock stand-ins for gradio, matplotlib,
    soundfile and pyloudnorm.  Returns the module object."""
    for name in ("gradio", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "soundfile", "pyloudnorm"):
        sys.modules.setdefault(name, unittest.mock.MagicMock(name=name))
    spec = importlib.util.spec_from_file_location("raytracer_studio_ref", REF)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def save(name, **arrays):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


HALLS = ["Plate", "Room", "Cathedral", "Garage"]           # last one exercises the fallback
MATERIALS = ["Stein", "Holz", "Teppich", "Glas", "Beton", "Vorhang (schwer)", "Gummi"]
LAYOUTS = ["Stereo", "5.1 (Standard)", "7.1 (Surround)", "5.1.2 (Atmos Light)", "9.1.6"]


def gold_scalars(rs):
    rows = []
    g = np.random.default_rng(100)
    for hall in HALLS:
        for _ in range(12):
            room = float(g.choice([10, 50, 100, 200, 430, 817, 1000, 20000]))
            x, y, z, dif, dw, e, l = g.uniform(0, 1, 7)
            e *= 2
            l *= 2
            dur, refl, mdel, split = quiet(rs.adjust_parameters_for_3d, hall, room, z)
            direc = quiet(rs.compute_final_directionality_3d, x, y, z, hall, dif, dw)
            ae, al = quiet(rs.adapt_early_late_levels, dw, e, l)
            rows.append([HALLS.index(hall), room, x, y, z, dif, dw, e, l,
                         dur, refl, mdel, split, direc, ae, al])
    save("scalars", table=np.asarray(rows, np.float64), halls=np.asarray(HALLS))


def gold_ir(rs):
    cases = [  # rate, hall, material, room, x,y,z, diffusion, dw, seed
        (48000, "Plate", "Beton", 10.0, .5, .5, .5, .5, .5, 1),
        (48000, "Room", "Holz", 200.0, .3, .4, .6, .5, .6, 11),
        (16000, "Cathedral", "Stein", 817.0, .2, .9, .1, .0, .9, 3),
        (8000, "Room", "Gummi", 50.0, .7, .1, .9, 1.0, .2, 7),
        (44100, "Plate", "Vorhang (schwer)", 1000.0, .5, .5, .5, .25, 1.0, 5),
        (100, "Plate", "Holz", 10.0, .5, .5, .5, .5, .5, 9),       # tiny IR (length 40)
    ]
    out = {}
    meta = []
    for i, (rate, hall, mat, room, x, y, z, dif, dw, seed) in enumerate(cases):
        dur, refl, mdel, split = quiet(rs.adjust_parameters_for_3d, hall, room, z)
        direc = quiet(rs.compute_final_directionality_3d, x, y, z, hall, dif, dw)
        np.random.seed(seed)
        e, l = quiet(rs.generate_impulse_response_split_3d, rate, dur, refl, mdel, mat, direc, split, dif)
        out[f"early{i}"] = e
        out[f"late{i}"] = l
        meta.append([rate, HALLS.index(hall), MATERIALS.index(mat), room, x, y, z, dif, dw, seed,
                     dur, refl, mdel, split, direc])
    # degenerate arguments (rs.py:246)
    e, l = quiet(rs.generate_impulse_response_split_3d, 0, 1.0, 10, 0.05, "Holz", 0.5, 0.05, 0.5)
    out["early_bad"], out["late_bad"] = e, l
    save("ir", meta=np.asarray(meta, np.float64), **out)


def gold_spectral(rs):
    g = np.random.default_rng(200)
    out = {}
    for tag, n in (("odd", 4099), ("even", 6000), ("short", 37)):
        sig = (0.4 * g.standard_normal((n, 2))).astype(np.float32)
        out[f"in_{tag}"] = sig
        for air in (0.005, 0.1, 0.65, 1.7):
            out[f"air_{tag}_{air}"] = quiet(rs.apply_simple_lp_filter, sig, 48000, air)
        out[f"air_{tag}_lowrate"] = quiet(rs.apply_simple_lp_filter, sig, 3000, 0.5)   # no bin >= 2 kHz
    save("air_filter", **out)

    out = {}
    for i, (nd, nw, dw, ks) in enumerate([(500, 500, .5, .5), (500, 620, .8, .5), (620, 500, .3, .5),
                                           (400, 400, 0., .5), (400, 400, 1., .5), (400, 400, .9, 1.0),
                                           (400, 400, .7, .9999999), (400, 400, .2, .0)]):
        d = g.standard_normal((nd, 2)).astype(np.float32)
        w = g.standard_normal((nw, 2)).astype(np.float32)
        out[f"dry{i}"], out[f"wet{i}"] = d, w
        out[f"par{i}"] = np.array([dw, ks])
        out[f"mix{i}"] = quiet(rs.dynamic_dry_wet_mix, d, w, dw, ks)
    save("dry_wet", **out)


def gold_convolve(rs):
    g = np.random.default_rng(300)
    rate = 48000
    out = {}
    # a short synthetic early/late pair with the reference's structure
    L = 1500
    early = np.zeros(L, np.float32)
    early[g.integers(1, 400, 12)] = g.uniform(0.1, 0.9, 12).astype(np.float32)
    late = np.zeros(L, np.float32)
    late[400:] = (g.uniform(-1, 1, L - 400) * 0.7 * 0.995 ** np.arange(L - 400)).astype(np.float32)
    out["early"], out["late"] = early, late
    xs = {"stereo": (0.3 * g.standard_normal((5003, 2))).astype(np.float32),
          "mono1d": (0.3 * g.standard_normal(4000)).astype(np.float32),
          "mono2d": (0.3 * g.standard_normal((4096, 1))).astype(np.float32),
          "six": (0.2 * g.standard_normal((3001, 6))).astype(np.float64),
          "loud": (1.5 * g.standard_normal((2500, 2))).astype(np.float32),
          "silent": np.zeros((1200, 2), np.float32),
          "tiny": (1e-12 * g.standard_normal((900, 2))).astype(np.float32)}
    for k, v in xs.items():
        out["x_" + k] = v
    cases = [  # input key, early lvl, late lvl, dw, bass, treble, kill, air, levels as np.float64?
        ("stereo", .6, .7, .5, 1.0, 1.0, .5, 0.0, True),       # pure convolution
        ("stereo", .6, .7, .5, 1.0, 1.0, .5, 0.1, True),       # air only
        ("stereo", .6, .7, .6, 1.5, 0.8, .5, 0.1, True),       # air + EQ
        ("stereo", .6, .7, .6, 1.5, 0.8, .5, 0.0, False),      # EQ only, python-float levels
        ("mono1d", .8, .6, .3, 1.0, 2.5, .5, 0.3, True),
        ("mono2d", .8, .6, 1.0, 0.4, 1.0, .5, 0.05, True),     # dw=1 -> dry killed
        ("six", .8, .6, 0.0, 1.0, 1.0, .5, 0.1, True),         # dw=0 -> dry only
        ("loud", 1.2, 1.4, .7, 3.0, 3.0, 1.0, 0.9, True),      # normalisation fires
        ("silent", .6, .7, .5, 1.3, 1.0, .5, 0.1, True),
        ("tiny", .6, .7, .5, 1.0, 1.0, .5, 0.0, True),         # < 1e-9 flush
        ("stereo", 1e-7, .7, .5, 1.0, 1.0, .5, 0.2, True),     # early branch gated off
        ("stereo", .6, 0.0, .5, 1.0, 1.0, .5, 0.2, True),      # late branch gated off
    ]
    par = []
    for i, (k, el, ll, dw, b, t, ks, air, as64) in enumerate(cases):
        lv = (np.float64(el), np.float64(ll)) if as64 else (el, ll)
        out[f"split{i}"] = quiet(rs.convolve_audio_split_3d, xs[k], early, late, lv[0], lv[1], dw, b, t, rate, ks, air)
        par.append([list(xs).index(k), el, ll, dw, b, t, ks, air, float(as64)])
    out["split_par"] = np.asarray(par)
    out["x_keys"] = np.asarray(list(xs))
    # external stereo IR
    ir = (g.standard_normal((1800, 2)) * np.exp(-np.arange(1800) / 300.0)[:, None]).astype(np.float32)
    ir /= np.max(np.abs(ir))
    out["ext_ir"] = ir
    ecases = [("stereo", .5, 1.0, 1.0, .5), ("stereo", .5, 1.2, 0.9, .5), ("mono1d", .9, 1.0, 1.0, .5),
              ("loud", .4, 0.3, 4.0, .2), ("six", 1.0, 1.0, 1.0, 1.0)]
    par = []
    for i, (k, dw, b, t, ks) in enumerate(ecases):
        out[f"ext{i}"] = quiet(rs.convolve_audio_external_ir, xs[k], ir, dw, b, t, rate, ks)
        par.append([list(xs).index(k), dw, b, t, ks])
    out["ext_par"] = np.asarray(par)
    out["ext_badir"] = quiet(rs.convolve_audio_external_ir, xs["stereo"], ir[:, 0], .5, 1.0, 1.0, rate, .5)
    save("convolve", **out)


def gold_panmap(rs):
    g = np.random.default_rng(400)
    out = {}
    sigs = {"norm": (0.35 * g.standard_normal((1400, 2))).astype(np.float32),
            "loud": (1.1 * g.standard_normal((1000, 2))).astype(np.float32),
            "tiny": (1e-11 * g.standard_normal((700, 2))).astype(np.float32),
            "short": (0.5 * g.standard_normal((300, 2))).astype(np.float32)}   # shorter than the 12/18 ms delays
    pos = [(.5, .5, .5), (.3, .4, .6), (0., 0., 0.), (1., 1., 1.), (.9, .05, .95)]
    par = []
    idx = 0
    for k, s in sigs.items():
        out["s_" + k] = s
        for (x, y, z) in pos:
            six = quiet(rs.apply_surround_panning_3d, s, x, y, z)
            out[f"pan{idx}"] = six
            for li, lay in enumerate(LAYOUTS):
                m, names = quiet(rs.map_channels, six.copy(), lay, 48000, z)
                out[f"map{idx}_{li}"] = m
            par.append([list(sigs).index(k), x, y, z])
            idx += 1
    out["par"] = np.asarray(par)
    out["sig_keys"] = np.asarray(list(sigs))
    out["layouts"] = np.asarray(LAYOUTS)
    # 44.1 kHz delay lengths
    six = quiet(rs.apply_surround_panning_3d, sigs["norm"], .4, .6, .8)
    out["pan_441"] = six
    out["map_441_71"] = quiet(rs.map_channels, six.copy(), "7.1 (Surround)", 44100, .8)[0]
    out["map_441_512"] = quiet(rs.map_channels, six.copy(), "5.1.2 (Atmos Light)", 44100, .8)[0]
    save("pan_map", **out)


class _Meter:
    """Stand-in for pyloudnorm.Meter backed by the oracle's restatement, so that the
    reference pipeline function can run end to end (LUFS itself stays 'parity unpinned')."""

    def __init__(self, rate):
        self.rate = rate

    def integrated_loudness(self, x):
        sys.path.insert(0, HERE)
        import ars_oracle
        return ars_oracle.integrated_loudness(x, self.rate)


def gold_metrics_pipeline(rs):
    g = np.random.default_rng(500)
    out = {}
    # peak / RMS half of calculate_audio_metrics on its own (LUFS mocked away)
    for i, (n, c, amp) in enumerate([(5000, 6, .3), (777, 2, 1.4), (1000, 8, 1e-20), (64, 1, .5)]):
        d = (amp * g.standard_normal((n, c))).astype(np.float32)
        m = quiet(rs.calculate_audio_metrics, d, 48000)
        out[f"m_in{i}"] = d
        out[f"m_out{i}"] = np.array([m["true_peak_dbfs"], m["rms_dbfs"]], np.float64)

    # full pipeline function with soundfile faked and pyloudnorm replaced by the restatement
    sf = sys.modules["soundfile"]
    sys.modules["pyloudnorm"].Meter = _Meter
    rs.pyln.Meter = _Meter
    store = {}

    def fake_read(path, dtype="float32", always_2d=True):
        return store[path]

    written = {}

    def fake_write(path, data, rate, subtype=None, format=None):
        written["data"], written["rate"], written["subtype"] = np.array(data), rate, subtype

    sf.read, sf.write = fake_read, fake_write
    tmpdir = tempfile.mkdtemp()
    irpath = os.path.join(tmpdir, "ir.wav")
    open(irpath, "wb").close()
    ir = (g.standard_normal((2400, 2)) * np.exp(-np.arange(2400) / 500.0)[:, None]).astype(np.float32)
    ir /= np.max(np.abs(ir))
    store[irpath] = (ir, 16000)
    out["pipe_ir"] = ir
    cases = [  # n, ch, rate, use_ext, hall, room, dif, air, early, late, dw, kill, bass, treble, x,y,z, material, layout, seed
        (9000, 1, 16000, False, "Room", 200., .5, .1, .8, .6, .6, .5, 1.5, .8, .3, .4, .6, "Holz", "5.1 (Standard)", 11),
        (7011, 2, 16000, False, "Plate", 10., .2, .0, .8, .6, .5, .5, 1., 1., .5, .5, .5, "Beton", "7.1 (Surround)", 2),
        (10000, 6, 16000, False, "Cathedral", 30., .7, .1, 1.0, 1.2, .5, .5, 1., 1., .8, .2, .9, "Stein", "5.1.2 (Atmos Light)", 3),
        (8000, 2, 16000, True, "Room", 100., .5, .1, .8, .6, .5, .5, 1.2, .9, .5, .5, .5, "Holz", "Stereo", 4),
        (7000, 2, 16000, True, "Room", 100., .5, .1, .8, .6, .7, .5, 1., 1., .25, .75, .5, "Holz", "5.1 (Standard)", 5),
        (5000, 2, 16000, False, "Room", 100., .5, .1, .8, .6, .5, .5, 1., 1., .5, .5, .5, "Glas", "Stereo", 6),  # < 0.4 s: LUFS N/A
    ]
    par = []
    for i, c in enumerate(cases):
        (n, ch, rate, ext, hall, room, dif, air, e, l, dw, ks, b, t, x, y, z, mat, lay, seed) = c
        audio = (0.25 * g.standard_normal((n, ch))).astype(np.float32)
        apath = os.path.join(tmpdir, f"a{i}.wav")
        store[apath] = (audio, rate)
        np.random.seed(seed)
        written.clear()
        p1, p2, text = quiet(rs.apply_raytrace_convolution_3d, apath, irpath if ext else None, ext, hall,
                             room, dif, air, e, l, dw, ks, b, t, x, y, z, mat, lay)
        assert p1 is not None, text
        if p1 and os.path.exists(p1):
            os.remove(p1)
        out[f"pipe_in{i}"] = audio
        out[f"pipe_out{i}"] = written["data"]
        out[f"pipe_text{i}"] = np.asarray(text)
        par.append([n, ch, rate, float(ext), HALLS.index(hall), room, dif, air, e, l, dw, ks, b, t, x, y, z,
                    MATERIALS.index(mat), LAYOUTS.index(lay), seed])
    out["pipe_par"] = np.asarray(par, np.float64)
    save("metrics_pipeline", **out)


def main():
    rs = load_reference()
    gold_scalars(rs)
    gold_ir(rs)
    gold_spectral(rs)
    gold_convolve(rs)
    gold_panmap(rs)
    gold_metrics_pipeline(rs)
    import scipy
    with open(os.path.join(GOLD, "VERSIONS.txt"), "w") as f:
        f.write(f"numpy {np.__version__}\nscipy {scipy.__version__}\npython {sys.version.split()[0]}\n"
                "generated by oracle/make_golden.py from /root/reference/raytracer_studio.py (unmodified)\n")


if __name__ == "__main__":
    main()
