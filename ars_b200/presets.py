"""The reference's preset JSON (rs.py:883-896 save, 913-927 load) as the description of a batch job.

The reference stores the sixteen controls of its UI under the keys below and, on load, falls back to a default per
missing / null / unparsable key.  `preset_to_settings` applies exactly those rules and renames the keys to the keyword
arguments of `render_array` / `render_batch`, so a folder of presets the Gradio app has written can drive a batch of
renders (BASELINE configs[3]: many clips, one preset each) without the UI:

    manifest = {"jobs": [{"audio": "a.wav", "preset": "presets/Kathedrale_v4.json", "out": "a_out.wav", "seed": 7},
                         {"audio": "b.wav", "preset": {"hall_type": "Plate", "dry_wet": 0.3}}]}
    results = render_manifest(manifest)          # one ars_render_batch call; PCM_16 WAVs written where "out" is given

Host logic only: file I/O is wavio.py, every array operation happens inside `ars_render_batch`.
"""
from __future__ import annotations

import json
import os

from . import wavio

# rs.py:883-887 / 918-920: the order of the UI controls
PRESET_KEYS = ("use_external_ir", "hall_type", "material", "room_size", "diffusion", "air_absorption", "early_level",
               "late_level", "dry_wet", "dry_wet_kill_start", "bass_gain", "treble_gain", "x_pos", "y_pos", "z_pos",
               "target_layout")
# rs.py:913-916
PRESET_DEFAULTS = {"use_external_ir": False, "hall_type": "Room", "material": "Holz", "room_size": 100.0, "diffusion": 0.5,
                   "air_absorption": 0.1, "early_level": 0.8, "late_level": 0.6, "dry_wet": 0.5,
                   "dry_wet_kill_start": 0.5, "bass_gain": 1.0, "treble_gain": 1.0, "x_pos": 0.5, "y_pos": 0.5,
                   "z_pos": 0.5, "target_layout": "5.1 (Standard)"}
_FLOAT_KEYS = ("room_size", "diffusion", "air_absorption", "early_level", "late_level", "dry_wet", "dry_wet_kill_start",
               "bass_gain", "treble_gain", "x_pos", "y_pos", "z_pos")
# preset key -> keyword of render_array (= parameter of apply_raytrace_convolution_3d, rs.py:991)
_RENAME = {"hall_type": "hall_type", "material": "material", "room_size": "room_size", "diffusion": "diffusion",
           "air_absorption": "air_absorption", "early_level": "base_early_level", "late_level": "base_late_level",
           "dry_wet": "dry_wet", "dry_wet_kill_start": "dry_wet_kill_start", "bass_gain": "bass_gain",
           "treble_gain": "treble_gain", "x_pos": "x_pos", "y_pos": "y_pos", "z_pos": "z_pos",
           "target_layout": "target_channel_layout"}


def normalize_preset(data: dict) -> dict:
    """The sixteen control values the reference's loader would put into the UI (rs.py:921-928)."""
    out = {}
    for key in PRESET_KEYS:
        value = data.get(key, PRESET_DEFAULTS[key]) if isinstance(data, dict) else PRESET_DEFAULTS[key]
        if value is None:
            value = PRESET_DEFAULTS[key]
        if key == "use_external_ir":
            value = bool(value)
        elif key in _FLOAT_KEYS:
            try:
                value = float(value)
            except (ValueError, TypeError):
                print(f"Warnung: Konnte Preset-Wert für '{key}' nicht in Float konvertieren.")
                value = PRESET_DEFAULTS[key]
        out[key] = value
    return out


def preset_to_settings(data: dict) -> tuple[dict, bool]:
    """-> (keyword settings for render_array / render_batch, use_external_ir)."""
    p = normalize_preset(data)
    return {_RENAME[k]: p[k] for k in PRESET_KEYS if k != "use_external_ir"}, p["use_external_ir"]


def settings_to_preset(settings: dict, use_external_ir: bool = False, name: str | None = None, version: str = "ars_b200") -> dict:
    """The inverse mapping, in the layout save_current_preset_v4 writes (rs.py:888-893)."""
    inv = {v: k for k, v in _RENAME.items()}
    data = {"use_external_ir": bool(use_external_ir)}
    for k in PRESET_KEYS[1:]:
        data[k] = PRESET_DEFAULTS[k]
    for kw, value in settings.items():
        if kw in inv:
            data[inv[kw]] = value
    data["_source_name"] = name
    data["_version"] = version
    return data


def load_preset(path: str) -> dict:
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def save_preset(path: str, data: dict):
    with open(path, "w", encoding="utf-8") as f:
        json.dump(data, f, indent=4, ensure_ascii=False)


def jobs_from_manifest(manifest, base_dir: str | None = None):
    """manifest: dict with "jobs" (or the list itself); each job has "audio" (WAV path), "preset" (path or dict),
    optional "external_ir" (stereo WAV path, used when the preset says use_external_ir), "seed", "out".
    -> (jobs for render_batch, per-job (rate, out path))."""
    from . import raytracer_studio as rs

    if isinstance(manifest, str):
        base_dir = base_dir or os.path.dirname(os.path.abspath(manifest))
        with open(manifest, "r", encoding="utf-8") as f:
            manifest = json.load(f)
    entries = manifest["jobs"] if isinstance(manifest, dict) else manifest

    def resolve(p):
        return p if base_dir is None or os.path.isabs(p) else os.path.join(base_dir, p)

    jobs, meta = [], []
    for e in entries:
        preset = e.get("preset", {})
        if isinstance(preset, str):
            preset = load_preset(resolve(preset))
        settings, use_ext = preset_to_settings(preset)
        samples, rate = wavio.read(resolve(e["audio"]))
        job = dict(samples=samples, rate=rate, **settings)
        if use_ext:
            if not e.get("external_ir"):
                raise ValueError("Externe IR gewählt, aber keine Datei gefunden.")          # rs.py:1029
            ir, ir_rate = wavio.read(resolve(e["external_ir"]))
            if ir.ndim != 2 or ir.shape[1] != 2:
                raise ValueError("Externe IR muss Stereo sein.")                            # rs.py:1036
            if ir_rate != rate:                                                              # rs.py:1037-1040
                ir = rs.resample_ir(ir, int(ir.shape[0] * rate / ir_rate))
            job["external_ir_data"] = ir
        if "seed" in e:
            job["seed"] = int(e["seed"])
        jobs.append(job)
        meta.append((rate, resolve(e["out"]) if e.get("out") else None))
    return jobs, meta


def render_manifest(manifest, base_dir: str | None = None, want_metrics: bool = True):
    """Render every job of a manifest in one pipelined batch; writes PCM_16 WAVs for the jobs that name an "out".
    -> list of dict(pcm, names, metrics, out)."""
    from . import raytracer_studio as rs

    jobs, meta = jobs_from_manifest(manifest, base_dir)
    results = rs.render_batch(jobs, want_float=False, want_pcm=True, want_metrics=want_metrics)
    for r, (rate, out) in zip(results, meta):
        r["out"] = out
        if out:
            wavio.write_pcm16(out, r["pcm"], rate)
    return results
