"""Drop-in for the render hot path of `raytracer_studio.py` (CipherCorePro/Audio-Raytracing-Studio).

Same function names, positional order, defaults, numpy-in / numpy-out behaviour and
"print and fall back, never raise" convention as the reference (`rs.py` = the reference
file; line numbers cited per function) -- but every array operation runs in hand-written
sm_100a CUDA kernels behind the C ABI of libars_b200 (include/ars_b200.h).  The scalar
prologue (hall presets, parameter shaping, the replay of the reference's random draws)
stays on the host, as SURVEY.md section 2.1 scopes it.

There is NO CPU implementation of the array work here: without the shared library and a
B200-class GPU the compute functions fail loudly (`ArsError`), they do not fall back.

Out of scope: Gradio UI, presets, plots, A/B report, analyser.py.
"""
from __future__ import annotations

import math
import os
import tempfile
import traceback

import numpy as np

from . import _capi
from ._capi import ArsError, ArsIrDraws, ArsMetrics, ArsRenderParams
from . import wavio

# ---- tables (rs.py:29-43) ---------------------------------------------------------------
material_absorption = {
    "Stein": 0.15, "Holz": 0.35, "Teppich": 0.7, "Glas": 0.2,
    "Beton": 0.1, "Vorhang (schwer)": 0.8,
}
DEFAULT_MATERIAL = "Holz"
DEFAULT_HALL_TYPE = "Room"
CHANNEL_LAYOUTS = {
    "Stereo": {"channels": 2, "names": ["FL", "FR"]},
    "5.1 (Standard)": {"channels": 6, "names": ["FL", "FR", "C", "LFE", "RL", "RR"]},
    "7.1 (Surround)": {"channels": 8, "names": ["FL", "FR", "C", "LFE", "RL", "RR", "SL", "SR"]},
    "5.1.2 (Atmos Light)": {"channels": 8, "names": ["FL", "FR", "C", "LFE", "RL", "RR", "TFL", "TFR"]},
}
DEFAULT_CHANNEL_LAYOUT = "5.1 (Standard)"

_F32 = np.float32


def _lib():
    return _capi.init()


# ---- scalar prologue (host) -------------------------------------------------------------
def adjust_reverb_parameters_by_hall(hall_type: str):
    """rs.py:157-166 -> (duration s, reflection count, max early delay s, early/late split s)."""
    table = {"Plate": (0.8, 25, 0.025, 0.03), "Room": (1.5, 35, 0.06, 0.08), "Cathedral": (4.0, 20, 0.10, 0.12)}
    if hall_type not in table:
        print(f"Warnung: Unbekannter Hall-Typ '{hall_type}', verwende Standardwerte für 'Room'.")
        return table["Room"]
    return table[hall_type]


def adapt_early_late_levels(dry_wet: float, base_early: float = 0.8, base_late: float = 0.6):
    """rs.py:168-182: early quieter / late louder as the mix gets wetter; returns np.float64 scalars."""
    try:
        dw = np.clip(float(dry_wet), 0.0, 1.0)
        base_early = float(base_early)
        base_late = float(base_late)
        bend = dw ** 1.5
        return (np.clip(base_early * (1.0 - (bend * 0.7)), 0.0, 2.0),
                np.clip(base_late * (1.0 + (bend * 0.6)), 0.0, 2.0))
    except Exception as e:
        print(f"Fehler in adapt_early_late_levels: {e}")
        return base_early, base_late


def compute_final_directionality_3d(x_pos, y_pos, z_pos, hall_type: str, diffusion_grade, dry_wet=0.5):
    """rs.py:184-209."""
    try:
        x, y, z = (np.clip(float(v), 0.0, 1.0) for v in (x_pos, y_pos, z_pos))
        diffusion = np.clip(float(diffusion_grade), 0.0, 1.0)
        dw = np.clip(float(dry_wet), 0.0, 1.0)
        off_axis = np.sqrt(((x - 0.5) * 2) ** 2 + ((z - 0.5) * 1.0) ** 2) / np.sqrt(1 ** 2 + 0.5 ** 2)
        off_depth = abs(y - 0.5) * 2
        position = np.clip((1.0 - off_axis * 0.3) * (1.0 - off_depth * 0.2), 0.5, 1.0)
        hall = {"Plate": 0.95, "Room": 0.65, "Cathedral": 0.25}.get(hall_type, 0.65)
        core = hall * position * (1.0 - (diffusion * 0.8))
        lift = max(0.0, (dw - 0.6) * 0.4)
        return np.clip(core + lift, 0.05, 0.95)
    except Exception as e:
        print(f"Fehler in compute_final_directionality_3d: {e}")
        traceback.print_exc()
        return 0.5


def adjust_parameters_for_3d(hall_type: str, room_size, z_pos):
    """rs.py:211-236 -> (duration, reflection count, max delay, split time)."""
    try:
        room_size = float(room_size)
        z_pos = float(z_pos)
        dur0, refl0, delay0, split0 = adjust_reverb_parameters_by_hall(hall_type)
        rel = room_size / 100.0
        k_dur = np.clip(rel ** 0.33, 0.5, 2.5)
        k_delay = np.clip(rel ** 0.25, 0.7, 1.8)
        k_refl = np.clip(1 + (room_size - 100) / 500.0, 0.8, 1.5)
        duration = np.clip(dur0 * k_dur, 0.1, 10.0)
        refl = np.clip(int(refl0 * k_refl), 5, 80)
        k_z = 1.0 + ((z_pos - 0.5) * 0.1)
        max_delay = np.clip(delay0 * k_delay * k_z, 0.01, 0.3)
        split = np.clip(split0 * k_delay, 0.02, 0.2)
        return duration, refl, max_delay, split
    except Exception as e:
        print(f"Fehler in adjust_parameters_for_3d: {e}")
        traceback.print_exc()
        return adjust_reverb_parameters_by_hall(DEFAULT_HALL_TYPE)


# ---- IR synthesis (a5) ------------------------------------------------------------------
def _ir_geometry(rate: int, duration: float, max_delay: float, split_time: float):
    """Integer geometry, rs.py:249,254-255,259,271-272 -> (length, split, tap_hi, late_len)."""
    length = max(1, int(duration * rate))
    split = max(1, min(int(split_time * rate), length - 1))
    tap_hi = min(max(2, int(max_delay * rate)), split)
    return length, split, tap_hi, length - split


def draw_ir_randoms(rate, ir_duration, reflection_count, max_delay, early_late_split, rng=np.random):
    """Replay of the reference's random call sequence (rs.py:262,264,285) on `rng` -- by default numpy's
    global legacy generator, exactly what the reference consumes, so `np.random.seed(s)` makes both
    sides draw identical taps and noise.  -> (tap delays int64[k], base strengths f64[k], noise f64[late_len])"""
    length, split, tap_hi, late_len = _ir_geometry(int(rate), float(ir_duration), float(max_delay),
                                                  float(early_late_split))
    delays, bases = [], []
    if int(reflection_count) > 0 and split > 1 and tap_hi > 1:
        for _ in range(int(reflection_count)):
            d = rng.randint(1, max(2, tap_hi))
            if 0 < d < split:
                delays.append(d)
                bases.append(rng.uniform(0.3, 0.8))
    noise = rng.uniform(-1, 1, size=late_len) if late_len > 0 else np.zeros(0)
    return np.asarray(delays, np.int64), np.asarray(bases, np.float64), np.asarray(noise, np.float64)


def generate_impulse_response_split_3d(rate, ir_duration, reflection_count, max_delay, material, directionality,
                                       early_late_split, diffusion_grade):
    """rs.py:238-308 -> (early_ir, late_ir), float32[int(duration*rate)] each.  The random draws are made
    here on the host from numpy's global generator (same calls, same order as the reference); the tap
    scatter, tail smoothing / envelope and both peak normalisations run on the GPU (K1)."""
    try:
        rate = int(rate)
        ir_duration = float(ir_duration)
        reflection_count = int(reflection_count)
        max_delay = float(max_delay)
        directionality = float(directionality)
        split_time = float(early_late_split)
        diffusion = float(diffusion_grade)
        if rate <= 0 or ir_duration <= 0:
            return np.array([1.0], dtype=_F32), np.zeros(1, dtype=_F32)
        absorption = material_absorption.get(material, material_absorption.get(DEFAULT_MATERIAL, 0.35))
        length = _ir_geometry(rate, ir_duration, max_delay, split_time)[0]
        taps, bases, noise = draw_ir_randoms(rate, ir_duration, reflection_count, max_delay, split_time)
        keep: list = []
        draws = _capi.make_draws(taps, bases, noise, keep)
        early = np.empty(length, _F32)
        late = np.empty(length, _F32)
        lib = _lib()
        _capi.check(lib.ars_ir_synth(float(rate), ir_duration, max_delay, float(absorption), directionality,
                                     split_time, diffusion, draws, _capi.ptr(early), _capi.ptr(late), length),
                    "ars_ir_synth")
        return early, late
    except ArsError:
        raise
    except Exception as e:
        print(f"Fehler in generate_impulse_response_split_3d: {e}")
        traceback.print_exc()
        return np.array([1.0], dtype=_F32), np.zeros(1, dtype=_F32)


# ---- spectral stages ----------------------------------------------------------------------
def apply_simple_lp_filter(signal, rate, air_absorption_factor):
    """rs.py:310-336: whole-signal air-absorption ramp above 2 kHz on the exact N-point DFT."""
    if air_absorption_factor < 0.01 or not isinstance(signal, np.ndarray) or signal.ndim != 2 or signal.size == 0:
        return signal
    try:
        n, ch = signal.shape
        if n < 2:
            return signal
        lib = _lib()
        out = np.empty((n, ch), _F32)
        # the library filters (n, 2) frames; other widths go through it two columns at a time
        for c0 in range(0, ch, 2):
            pair = np.zeros((n, 2), _F32)
            w = min(2, ch - c0)
            pair[:, :w] = signal[:, c0:c0 + w]
            res = np.empty((n, 2), _F32)
            _capi.check(lib.ars_air_filter(_capi.ptr(pair), n, float(rate), float(air_absorption_factor),
                                           _capi.ptr(res)), "ars_air_filter")
            out[:, c0:c0 + w] = res[:, :w]
        return out
    except ArsError:
        raise
    except Exception as e:
        print(f"Fehler im Luftabsorptionsfilter: {e}")
        traceback.print_exc()
        return signal


def resample_ir(ir, num):
    """scipy.signal.resample(ir, num, axis=0) for a stereo IR (rs.py:1039), on the GPU."""
    a = np.ascontiguousarray(ir, dtype=_F32)
    if a.ndim != 2 or a.shape[1] != 2 or a.shape[0] == 0 or num <= 0:
        raise ValueError("resample_ir: needs an (n, 2) array and num > 0")
    out = np.empty((int(num), 2), _F32)
    _capi.check(_lib().ars_resample(_capi.ptr(a), a.shape[0], int(num), _capi.ptr(out)), "ars_resample")
    return out


def dynamic_dry_wet_mix(dry_signal, wet_signal, dry_wet, kill_start=0.5):
    """rs.py:84-144."""
    try:
        dry = np.ascontiguousarray(np.asarray(dry_signal, dtype=_F32))
        wet = np.ascontiguousarray(np.asarray(wet_signal, dtype=_F32))
        dry_wet = float(dry_wet)
        kill_start = float(kill_start)
        if dry.ndim != wet.ndim or dry.shape[1:] != wet.shape[1:]:
            raise ValueError(f"operands could not be broadcast together with shapes {dry.shape} {wet.shape}")
        total = max(dry.shape[0], wet.shape[0])
        ch = int(np.prod(dry.shape[1:])) if dry.ndim > 1 else 1
        out = np.empty((total,) + dry.shape[1:], _F32)
        if total == 0 or ch == 0:
            return out
        _capi.check(_lib().ars_dry_wet_mix(_capi.ptr(dry), dry.shape[0], _capi.ptr(wet), wet.shape[0], ch, dry_wet,
                                           kill_start, _capi.ptr(out)), "ars_dry_wet_mix")
        return out
    except ArsError:
        raise
    except Exception as e:
        print(f"Fehler in dynamic_dry_wet_mix: {e}")
        traceback.print_exc()
        return np.array([], dtype=_F32)


def _as_frames(data):
    """float32 C-contiguous (n, cin) view of the input; 1-D becomes (n, 1) (the kernels duplicate mono and
    read only the first two channels, rs.py:343-346)."""
    a = np.asarray(data)
    if a.ndim == 1:
        a = a[:, None]
    return np.ascontiguousarray(a, dtype=_F32)


def convolve_audio_split_3d(data, early_ir, late_ir, early_level, late_level, dry_wet, bass_gain=1.0,
                            treble_gain=1.0, rate=44100, kill_start_dw=0.5, air_absorption_factor=0.0):
    """rs.py:338-408: early + late convolution, air absorption on the late part, dry/wet with dry-kill,
    brick-wall bass/treble EQ, peak guard -> (n + L - 1, 2) float32.  One fused N-point spectral filter on
    the GPU (K2-K5)."""
    if data is None or np.size(data) == 0:
        return np.zeros((0, 2), dtype=_F32)
    x = _as_frames(data)
    e = np.ascontiguousarray(np.asarray(early_ir, dtype=_F32).flatten()) if early_ir is not None else None
    l = np.ascontiguousarray(np.asarray(late_ir, dtype=_F32).flatten()) if late_ir is not None else None
    lib = _lib()
    n, cin = x.shape
    le = int(e.size) if e is not None else 0
    ll = int(l.size) if l is not None else 0
    n_out = lib.ars_convolve_out_len(n, le, ll)
    out = np.empty((n_out, 2), _F32)
    _capi.check(lib.ars_convolve_split(_capi.ptr(x), n, cin, _capi.ptr(e) if le else None, le,
                                       _capi.ptr(l) if ll else None, ll, float(early_level), float(late_level),
                                       float(dry_wet), float(bass_gain), float(treble_gain), float(rate),
                                       float(kill_start_dw), float(air_absorption_factor), _capi.ptr(out)),
                "ars_convolve_split")
    return out


def convolve_audio_external_ir(data, external_ir_data, dry_wet, bass_gain=1.0, treble_gain=1.0, rate=44100,
                               kill_start_dw=0.5):
    """rs.py:410-462: per-channel convolution with a stereo IR, dry/wet, EQ, peak guard."""
    if data is None or np.size(data) == 0:
        return np.zeros((0, 2), dtype=_F32)
    if (external_ir_data is None or not isinstance(external_ir_data, np.ndarray) or external_ir_data.ndim != 2
            or external_ir_data.shape[1] != 2):
        print("FEHLER (conv_ext): Ungültige externe IR Daten.")
        return data.astype(_F32)
    x = _as_frames(data)
    ir = np.ascontiguousarray(external_ir_data, dtype=_F32)
    n, cin = x.shape
    L = ir.shape[0]
    if L == 0:
        print("FEHLER (conv_ext): Ungültige externe IR Daten.")
        return data.astype(_F32)
    out = np.empty((n + L - 1, 2), _F32)
    _capi.check(_lib().ars_convolve_external(_capi.ptr(x), n, cin, _capi.ptr(ir), L, float(dry_wet), float(bass_gain),
                                             float(treble_gain), float(rate), float(kill_start_dw), _capi.ptr(out)),
                "ars_convolve_external")
    return out


# ---- panner / mapper ------------------------------------------------------------------------
def _as_stereo(audio):
    a = np.asarray(audio)
    if a.ndim == 1:
        a = np.stack((a, a), axis=1)
    elif a.shape[1] == 1:
        a = np.repeat(a, 2, axis=1)
    elif a.shape[1] != 2:
        a = a[:, :2]
    return np.ascontiguousarray(a, dtype=_F32)


def apply_surround_panning_3d(audio_data, x_pos, y_pos, z_pos):
    """rs.py:464-505: stereo -> FL FR C LFE RL RR by square-root gains of (x, y, z); peak guard."""
    if audio_data is None or np.size(audio_data) == 0:
        return np.zeros((0, 6), dtype=_F32)
    try:
        s = _as_stereo(audio_data)
        out = np.empty((s.shape[0], 6), _F32)
        _capi.check(_lib().ars_pan(_capi.ptr(s), s.shape[0], float(x_pos), float(y_pos), float(z_pos), _capi.ptr(out)),
                    "ars_pan")
        return out
    except ArsError:
        raise
    except Exception as e:
        print(f"Fehler in apply_surround_panning_3d: {e}")
        traceback.print_exc()
        n = audio_data.shape[0] if audio_data is not None and getattr(audio_data, "ndim", 0) == 2 else 0
        return np.zeros((n, 6), dtype=_F32)


def apply_delay(signal, delay_samples):
    """rs.py:507-515: prepend `delay_samples` zero frames, trim to the original length."""
    if not isinstance(signal, np.ndarray) or signal.ndim != 2:
        return signal
    delay_samples = int(delay_samples)
    if delay_samples <= 0:
        return signal
    n, ch = signal.shape
    if n == 0 or ch == 0:
        return signal
    src = np.ascontiguousarray(signal, dtype=_F32)
    out = np.empty((n, ch), _F32)
    _capi.check(_lib().ars_delay(_capi.ptr(src), n, ch, delay_samples, _capi.ptr(out)), "ars_delay")
    return out.astype(signal.dtype, copy=False)


def map_channels(data_5_1, target_layout_name, rate, z_pos=0.5):
    """rs.py:517-571: 6 channels -> Stereo / 5.1 / 7.1 (delayed side pair) / 5.1.2 (delayed height pair);
    peak guard.  As in the reference, "5.1 (Standard)" hands back the input object (scaled in place if it
    peaks above 1)."""
    if target_layout_name not in CHANNEL_LAYOUTS:
        print(f"Warnung: Unbekanntes Ziel-Layout '{target_layout_name}'. Nutze 5.1 Standard.")
        target_layout_name = DEFAULT_CHANNEL_LAYOUT
    info = CHANNEL_LAYOUTS[target_layout_name]
    ch, names = info["channels"], info["names"]
    if data_5_1 is None or not isinstance(data_5_1, np.ndarray) or data_5_1.ndim != 2 or data_5_1.shape[1] != 6:
        print("Fehler (map_channels): Eingangsdaten sind kein gültiges 6-Kanal-Audio.")
        return np.zeros((0, ch), dtype=_F32), names
    n = data_5_1.shape[0]
    if n == 0:
        return (data_5_1 if target_layout_name == "5.1 (Standard)" else np.zeros((0, ch), data_5_1.dtype)), names
    src = np.ascontiguousarray(data_5_1, dtype=_F32)
    out = np.empty((n, ch), _F32)
    _capi.check(_lib().ars_map_channels(_capi.ptr(src), n, _capi.LAYOUT_IDS[target_layout_name], float(int(rate)),
                                        float(z_pos), _capi.ptr(out)), "ars_map_channels")
    if target_layout_name == "5.1 (Standard)":
        data_5_1[...] = out          # same object, normalised in place (rs.py:538,559)
        return data_5_1, names
    return out, names


# ---- metrics ----------------------------------------------------------------------------------
def _metrics_dict(m: ArsMetrics):
    lufs = None
    if m.lufs_status == _capi.LUFS_OK:
        lufs = float(m.lufs)
    out = {"lufs": lufs, "true_peak_dbfs": float(m.true_peak_dbfs), "rms_dbfs": float(m.rms_dbfs)}
    if m.true_peak_4x_status == 0:      # add-on, only when asked for (want_true_peak_4x): BS.1770-4 oversampled peak
        out["true_peak_4x_dbfs"] = float(m.true_peak_4x_dbfs)
    return out


def calculate_audio_metrics(data, rate):
    """rs.py:674-711: integrated loudness of mean(ch0, ch1) (K-weighted, gated), sample peak and RMS in
    dBFS over all channels.  `lufs` is None when the clip is shorter than one 400 ms block."""
    metrics = {"lufs": None, "true_peak_dbfs": None, "rms_dbfs": None}
    if data is None or not isinstance(data, np.ndarray) or data.size == 0 or rate <= 0:
        return metrics
    if data.ndim != 2:
        if data.ndim == 1:
            data = data[:, np.newaxis]
        else:
            print(f"Warnung (Metriken): Ungültige Datenform {data.shape}.")
            return metrics
    n, ch = data.shape
    if ch == 0:
        return metrics
    try:
        # pyloudnorm refuses non-floating input; the reference catches that inside its own try (rs.py:685-693): lufs stays
        # None, peak and RMS are still computed (rs.py:694-698)
        want_lufs = 1 if np.issubdtype(data.dtype, np.floating) else 0
        if not want_lufs:
            print("Fehler bei LUFS-Berechnung: Data must be floating point.")
        x = np.ascontiguousarray(data, dtype=_F32)
        m = ArsMetrics()
        _capi.check(_lib().ars_metrics(_capi.ptr(x), n, ch, float(rate), want_lufs, m), "ars_metrics")
        out = _metrics_dict(m)
        if not want_lufs:          # (the reference's silence test comes before the meter call, rs.py:689)
            feed = data[:, 0] if ch == 1 else np.mean(data[:, :2], axis=1)
            if np.max(np.abs(feed)) < 1e-6:
                out["lufs"] = -np.inf
        return out
    except ArsError:
        raise
    except Exception as e:
        print(f"Fehler bei Metrikberechnung: {e}")
        traceback.print_exc()
        return metrics


def true_peak_4x(data):
    """4x-oversampled true peak in dBTP over all channels (ITU-R BS.1770-4 Annex 2) -- an add-on: the reference's
    `true_peak_dbfs` is the sample peak (rs.py:695-697) and `calculate_audio_metrics` keeps returning that."""
    x = np.ascontiguousarray(data, dtype=_F32)
    if x.ndim == 1:
        x = x[:, None]
    n, ch = x.shape
    if n == 0 or ch == 0:
        return -np.inf
    out = _capi.C.c_double(0)
    _capi.check(_lib().ars_true_peak_4x(_capi.ptr(x), n, ch, _capi.C.byref(out)), "ars_true_peak_4x")
    return float(out.value)


def channel_levels(data):
    """Numerics of the reference's A/B report (rs.py:769-798): -> (per-channel RMS dBFS list, side-signal RMS of the
    first two channels -- the report's "stereo width" metric; 0.0 for mono)."""
    x = np.ascontiguousarray(data, dtype=_F32)
    if x.ndim == 1:
        x = x[:, None]
    n, ch = x.shape
    if n == 0 or ch == 0:
        return [], 0.0
    rms = np.empty(ch, _F32)
    side = _capi.C.c_float(0)
    _capi.check(_lib().ars_channel_rms(_capi.ptr(x), n, ch, _capi.ptr(rms), _capi.C.byref(side)), "ars_channel_rms")
    return [20 * math.log10(float(v)) if v > 1e-15 else -np.inf for v in rms], float(side.value)


def spectrogram(data, rate):
    """Numerics of the visualiser's spectrogram (rs.py:621-634): first channel, nperseg 4096 / 2048 / 1024 by duration
    (> 30 s / > 5 s / else), 50 % overlap, Hann, scipy's defaults -> (f, t, Sxx) as scipy.signal.spectrogram returns them.
    The dB scaling and colour limits (rs.py:636-639) are plotting and stay with the reference."""
    x = np.ascontiguousarray(data, dtype=_F32)
    if x.ndim == 1:
        x = x[:, None]
    n, ch = x.shape
    duration = n / rate if rate > 0 else 0
    nperseg = 4096 if duration > 30 else 2048 if duration > 5 else 1024
    nperseg = min(nperseg, n)
    if nperseg < 2:
        raise ValueError("Signal zu kurz für Spektrogramm.")
    if nperseg & (nperseg - 1):
        raise ValueError("Spektrogramm: Segmentlänge muss eine Zweierpotenz sein (Signal kürzer als 1024 Samples).")
    nseg = int(_lib().ars_spectrogram_segments(n, nperseg))
    sxx = np.empty((nperseg // 2 + 1, nseg), _F32)
    _capi.check(_lib().ars_spectrogram(_capi.ptr(x), n, ch, float(rate), nperseg, _capi.ptr(sxx)), "ars_spectrogram")
    f = np.fft.rfftfreq(nperseg, 1.0 / rate)
    t = (np.arange(nseg) * (nperseg - nperseg // 2) + nperseg / 2) / float(rate)
    return f, t, sxx


def float_to_pcm16(data):
    """clip +-0.9999, scrub non-finite, float -> int16 (rs.py:1082-1084 + libsndfile's rule)."""
    x = np.ascontiguousarray(data, dtype=_F32)
    out = np.empty(x.shape, np.int16)
    if x.size:
        _capi.check(_lib().ars_pcm16(_capi.ptr(x), x.size, _capi.ptr(out)), "ars_pcm16")
    return out


# ---- whole render on arrays ---------------------------------------------------------------------
def make_render_params(rate, *, external_ir=False, hall_type="Room", room_size=100.0, diffusion=0.5,
                       air_absorption=0.1, base_early_level=0.8, base_late_level=0.6, dry_wet=0.5,
                       dry_wet_kill_start=0.5, bass_gain=1.0, treble_gain=1.0, x_pos=0.5, y_pos=0.5, z_pos=0.5,
                       material=DEFAULT_MATERIAL, target_channel_layout=DEFAULT_CHANNEL_LAYOUT, want_lufs=True,
                       ir_duration=None, want_true_peak_4x=False):
    """Host prologue of rs.py:1049-1056 -> (ArsRenderParams, reflection count).  `ir_duration` overrides the
    duration derived from hall/room (used for the '8 s IR' benchmark variant)."""
    p = ArsRenderParams()
    p.rate = float(int(rate))
    p.external_ir = 1 if external_ir else 0
    layout = target_channel_layout if target_channel_layout in CHANNEL_LAYOUTS else DEFAULT_CHANNEL_LAYOUT
    p.layout = _capi.LAYOUT_IDS[layout]
    refl = 0
    if not external_ir:
        dur, refl, mdel, split = adjust_parameters_for_3d(hall_type, room_size, z_pos)
        if ir_duration is not None:
            dur = float(ir_duration)
        p.ir_duration, p.ir_max_delay, p.ir_split_time = float(dur), float(mdel), float(split)
        p.directionality = float(compute_final_directionality_3d(x_pos, y_pos, z_pos, hall_type, diffusion, dry_wet))
        p.absorption = float(material_absorption.get(material, material_absorption.get(DEFAULT_MATERIAL, 0.35)))
        p.diffusion = float(diffusion)
        e_lvl, l_lvl = adapt_early_late_levels(dry_wet, base_early_level, base_late_level)
        p.early_level, p.late_level = float(e_lvl), float(l_lvl)
        p.air_absorption = float(air_absorption)
    p.dry_wet, p.kill_start = float(dry_wet), float(dry_wet_kill_start)
    p.bass_gain, p.treble_gain = float(bass_gain), float(treble_gain)
    p.x, p.y, p.z = float(x_pos), float(y_pos), float(z_pos)
    p.want_lufs = (1 if want_lufs else 0) | (2 if want_true_peak_4x else 0)
    return p, int(refl)


def render_array(samples, rate, *, external_ir_data=None, want_stereo=False, want_float=True, want_pcm=True,
                 want_metrics=True, **settings):
    """Compute part of apply_raytrace_convolution_3d (rs.py:1020-1084) on arrays, one fused GPU pipeline:
    IR synthesis (or external stereo IR) -> spectral filter -> pan -> map -> metrics -> int16.
    -> dict(stereo, final, names, metrics, pcm)."""
    x = _as_frames(samples)
    n, cin = x.shape
    if n == 0:
        raise ValueError("Audiodatei ist leer.")
    ext = external_ir_data is not None
    p, refl = make_render_params(rate, external_ir=ext, want_lufs=want_metrics, **settings)
    lib = _lib()
    keep: list = []
    draws = None
    ir = None
    L = 0
    if ext:
        ir = np.ascontiguousarray(external_ir_data, dtype=_F32)
        if ir.ndim != 2 or ir.shape[1] != 2 or ir.shape[0] == 0:
            raise ValueError("Externe IR muss Stereo sein.")
        L = ir.shape[0]
    else:
        taps, bases, noise = draw_ir_randoms(int(rate), p.ir_duration, refl, p.ir_max_delay, p.ir_split_time)
        draws = _capi.make_draws(taps, bases, noise, keep)
    N = lib.ars_render_out_len(p, n, L)
    layout = settings.get("target_channel_layout", DEFAULT_CHANNEL_LAYOUT)
    if layout not in CHANNEL_LAYOUTS:
        layout = DEFAULT_CHANNEL_LAYOUT
    C = CHANNEL_LAYOUTS[layout]["channels"]
    stereo = _capi.result_empty((N, 2), _F32) if want_stereo else None
    final = _capi.result_empty((N, C), _F32) if want_float else None
    pcm = _capi.result_empty((N, C), np.int16) if want_pcm else None
    m = ArsMetrics() if want_metrics else None
    _capi.check(lib.ars_render(p, _capi.ptr(x), n, cin, _capi.ptr(ir), L, draws, _capi.ptr(stereo), _capi.ptr(final),
                               _capi.ptr(pcm), m), "ars_render")
    return {"stereo": stereo, "final": final, "names": CHANNEL_LAYOUTS[layout]["names"],
            "metrics": _metrics_dict(m) if m is not None else None, "pcm": pcm}


def render_batch(jobs, *, want_float=False, want_pcm=True, want_metrics=True):
    """Render a list of independent clips with the copy/compute pipeline of `ars_render_batch` (the batch form
    of apply_raytrace_convolution_3d's compute part).  Each job is a dict: samples, rate, optional
    external_ir_data, optional seed (np.random.seed before that clip's draws, as a caller of the reference would
    do), and the keyword settings of `render_array`.  -> list of dict(final, names, metrics, pcm)."""
    lib = _lib()
    keep: list = []
    clips = (_capi.ArsClip * len(jobs))()
    outs = []
    for i, job in enumerate(jobs):
        job = dict(job)
        x = _as_frames(job.pop("samples"))
        rate = job.pop("rate")
        ir = job.pop("external_ir_data", None)
        seed = job.pop("seed", None)
        n, cin = x.shape
        if n == 0:
            raise ValueError("Audiodatei ist leer.")
        p, refl = make_render_params(rate, external_ir=ir is not None, want_lufs=want_metrics, **job)
        L = 0
        draws = None
        if ir is not None:
            ir = np.ascontiguousarray(ir, dtype=_F32)
            if ir.ndim != 2 or ir.shape[1] != 2 or ir.shape[0] == 0:
                raise ValueError("Externe IR muss Stereo sein.")
            L = ir.shape[0]
        else:
            if seed is not None:
                np.random.seed(seed)
            taps, bases, noise = draw_ir_randoms(int(rate), p.ir_duration, refl, p.ir_max_delay, p.ir_split_time)
            draws = _capi.make_draws(taps, bases, noise, keep)
        N = lib.ars_render_out_len(p, n, L)
        layout = job.get("target_channel_layout", DEFAULT_CHANNEL_LAYOUT)
        if layout not in CHANNEL_LAYOUTS:
            layout = DEFAULT_CHANNEL_LAYOUT
        C = CHANNEL_LAYOUTS[layout]["channels"]
        final = _capi.result_empty((N, C), _F32) if want_float else None
        pcm = _capi.result_empty((N, C), np.int16) if want_pcm else None
        m = ArsMetrics() if want_metrics else None
        keep += [x, ir, p, draws, m]
        k = clips[i]
        k.params = _capi.C.pointer(p)
        k.in_ = _capi.ptr(x)
        k.n, k.cin = n, cin
        k.ext_ir, k.ext_ir_len = _capi.ptr(ir), L
        k.draws = _capi.C.pointer(draws) if draws is not None else None
        k.out_f32, k.out_pcm = _capi.ptr(final), _capi.ptr(pcm)
        k.metrics = _capi.C.pointer(m) if m is not None else None
        outs.append((final, pcm, m, CHANNEL_LAYOUTS[layout]["names"]))
    _capi.check(lib.ars_render_batch(clips, len(jobs)), "ars_render_batch")
    return [{"final": f, "pcm": q, "names": nm, "metrics": _metrics_dict(m) if m is not None else None}
            for (f, q, m, nm) in outs]


def _metrics_text(mt):
    """rs.py:1071-1075."""
    lufs, peak, rms = mt.get("lufs"), mt.get("true_peak_dbfs"), mt.get("rms_dbfs")
    lufs_s = f"{lufs:.2f}" if lufs is not None and not np.isinf(lufs) else "N/A"
    peak_s = f"{peak:.1f}" if peak is not None and not np.isinf(peak) else "-inf"
    rms_s = f"{rms:.1f}" if rms is not None and not np.isinf(rms) else "-inf"
    return f"LUFS: {lufs_s} | Peak: {peak_s} dBFS | RMS: {rms_s} dBFS"


def _read_audio(path):
    """sf.read(path, dtype='float32', always_2d=True) of the reference (rs.py:1013, 1034).  RIFF/WAVE files go through the
    native reader (wavio.py); anything else (FLAC, OGG, AIFF ...) needs libsndfile and is read through `soundfile` when
    that package is importable -- file decoding is outside the render path."""
    try:
        return wavio.read(path)
    except Exception as wav_error:
        try:
            import soundfile as sf
        except ImportError:
            raise wav_error
        data, rate = sf.read(path, dtype="float32", always_2d=True)
        return np.ascontiguousarray(data, dtype=_F32), int(rate)


def apply_raytrace_convolution_3d(audio_file_path, external_ir_path, use_external_ir_cb, hall_type_val, room_size_val,
                                  diffusion_val, air_absorption_val, base_early_level, base_late_level, dry_wet,
                                  dry_wet_kill_start, bass_gain, treble_gain, x_pos, y_pos, z_pos, material,
                                  target_channel_layout):
    """rs.py:991-1125: WAV in -> (wav_path, wav_path, "LUFS: .. | Peak: .. dBFS | RMS: .. dBFS"), or
    (None, None, message) on any failure.  File I/O is the native WAV codec in wavio.py (the reference
    uses libsndfile); the render itself is one `ars_render` call."""
    try:
        try:
            use_ext = bool(use_external_ir_cb)
            room = float(room_size_val)
            diffusion = float(diffusion_val)
            air = float(air_absorption_val)
            early, late = float(base_early_level), float(base_late_level)
            dw, kill = float(dry_wet), float(dry_wet_kill_start)
            bass, treble = float(bass_gain), float(treble_gain)
            x, y, z = float(x_pos), float(y_pos), float(z_pos)
            if not isinstance(hall_type_val, str) or not isinstance(material, str) or \
                    not isinstance(target_channel_layout, str):
                raise ValueError("Ungültiger String-Inputtyp.")
        except (ValueError, TypeError, AttributeError) as e:
            msg = f"Fehlerhafte Eingabeparameter: {e}"
            print(f"ERROR: {msg}")
            return None, None, msg
        try:
            samples, rate = _read_audio(audio_file_path)
            if samples.size == 0:
                raise ValueError("Audiodatei ist leer.")
        except Exception as e:
            msg = f"Fehler beim Laden: {e}"
            print(f"ERROR: {msg}")
            return None, None, msg
        ir = None
        if use_ext:
            ir_path = getattr(external_ir_path, "name", external_ir_path)
            if not ir_path or not os.path.exists(ir_path):
                msg = "Externe IR gewählt, aber keine Datei gefunden."
                print(f"WARNUNG: {msg}")
                return None, None, msg
            try:
                ir, ir_rate = _read_audio(ir_path)
                if ir.size == 0:
                    raise ValueError("Externe IR-Datei ist leer.")
                if ir.ndim != 2 or ir.shape[1] != 2:
                    msg = "Externe IR muss Stereo sein."
                    print(f"ERROR: {msg}")
                    return None, None, msg
                if ir_rate != rate:            # rs.py:1037-1040
                    num = int(ir.shape[0] * rate / ir_rate)
                    if num <= 0:
                        raise ValueError("Resampling würde IR-Länge Null ergeben.")
                    ir = resample_ir(ir, num)
            except Exception as e:
                msg = f"Fehler Laden/Resample IR: {e}"
                print(f"ERROR: {msg}")
                return None, None, msg
        res = render_array(samples, rate, external_ir_data=ir, want_float=False, want_pcm=True, want_metrics=True,
                           hall_type=hall_type_val, room_size=room, diffusion=diffusion, air_absorption=air,
                           base_early_level=early, base_late_level=late, dry_wet=dw, dry_wet_kill_start=kill,
                           bass_gain=bass, treble_gain=treble, x_pos=x, y_pos=y, z_pos=z, material=material,
                           target_channel_layout=target_channel_layout)
        text = _metrics_text(res["metrics"])
        path = None
        try:
            with tempfile.NamedTemporaryFile(delete=False, suffix=".wav", prefix="processed_") as f:
                path = f.name
            wavio.write_pcm16(path, res["pcm"], rate)
            return path, path, text
        except Exception as e:
            msg = f"Fehler beim Schreiben der WAV-Datei: {e}"
            print(f"ERROR: {msg}")
            if path and os.path.exists(path):         # rs.py:1088-1091: no half-written temp file is left behind
                try:
                    os.remove(path)
                except OSError:
                    pass
            return None, None, msg
    except Exception as e:
        # rs.py:1096-1109: this entry point never raises -- a failure of the GPU library (ArsError) included; the
        # stage-level functions keep raising, there is no CPU path to fall back to
        msg = f"Unerwarteter Fehler: {e}"
        print(f"ERROR: {msg}")
        traceback.print_exc()
        return None, None, msg
