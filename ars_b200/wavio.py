"""Minimal WAV container codec for the pipeline edge (rs.py:1013, 1034, 1084).

The reference reads with `soundfile.read(dtype='float32', always_2d=True)` and writes with
`soundfile.write(subtype='PCM_16', format='WAV')` (libsndfile).  This module covers the same two
calls for RIFF/WAVE files: PCM 8/16/24/32-bit, IEEE float 32/64, plain and EXTENSIBLE headers.
Integer samples are scaled by 1 / 2^(bits-1) on read, as libsndfile does.  The float -> int16
conversion itself is NOT done here: the GPU epilogue produces the int16 frames (ars_pcm16 /
ars_render); this module only wraps them in a header.
"""
from __future__ import annotations

import struct

import numpy as np

_PCM, _FLOAT, _EXT = 1, 3, 0xFFFE


def read(path):
    """-> (float32 array of shape (frames, channels), sample rate)."""
    with open(path, "rb") as f:
        blob = f.read()
    if len(blob) < 12 or blob[:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos = 12
    fmt = None
    data = None
    while pos + 8 <= len(blob):
        cid, size = blob[pos:pos + 4], struct.unpack_from("<I", blob, pos + 4)[0]
        body = blob[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            data = body
            break
        pos += 8 + size + (size & 1)
    if fmt is None or data is None or len(fmt) < 16:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, rate, _, align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if tag == _EXT and len(fmt) >= 26:
        tag = struct.unpack_from("<H", fmt, 24)[0]
    if ch == 0 or align == 0:
        raise ValueError(f"{path}: bad fmt chunk")
    frames = len(data) // align
    raw = np.frombuffer(data, dtype=np.uint8, count=frames * align)
    if tag == _FLOAT and bits == 32:
        out = raw.view("<f4").astype(np.float32)
    elif tag == _FLOAT and bits == 64:
        out = raw.view("<f8").astype(np.float32)
    elif tag == _PCM and bits == 16:
        out = raw.view("<i2").astype(np.float32) / np.float32(32768.0)
    elif tag == _PCM and bits == 8:
        out = (raw.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    elif tag == _PCM and bits == 24:
        b = raw.reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        out = v.astype(np.float32) / np.float32(8388608.0)
    elif tag == _PCM and bits == 32:
        out = (raw.view("<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAV encoding (tag {tag}, {bits} bit)")
    return np.ascontiguousarray(out.reshape(frames, ch)), int(rate)


def write_pcm16(path, frames_i16: np.ndarray, rate: int):
    """Write interleaved int16 frames (frames, channels) as a WAVE_FORMAT_PCM file."""
    a = np.ascontiguousarray(frames_i16, dtype="<i2")
    if a.ndim == 1:
        a = a[:, None]
    ch = a.shape[1]
    payload = a.tobytes()
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + len(payload), b"WAVE", b"fmt ", 16, _PCM, ch, int(rate),
                      int(rate) * ch * 2, ch * 2, 16, b"data", len(payload))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(payload)
        if len(payload) & 1:
            f.write(b"\0")


def write_float32(path, frames_f32: np.ndarray, rate: int):
    """Write float32 frames (test fixtures / inputs)."""
    a = np.ascontiguousarray(frames_f32, dtype="<f4")
    if a.ndim == 1:
        a = a[:, None]
    ch = a.shape[1]
    payload = a.tobytes()
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + len(payload), b"WAVE", b"fmt ", 16, _FLOAT, ch, int(rate),
                      int(rate) * ch * 4, ch * 4, 32, b"data", len(payload))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(payload)
