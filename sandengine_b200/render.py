"""Headless read-out of the reference's `output_color`: the colour buffer that the compute shader
writes (operations.glsl:100-108,170) is shaded on demand by kernel `se_shade` and read back as packed RGBA8
(`Simulation.download_color(rgba8=True)`); this module writes it as a PNG with nothing but zlib.

The window path of the reference (sandengine-core/src/renderer.rs:212-246, shaders/fragment140.glsl: blur-based
occlusion, a background image, sky light) is presentation and stays out of scope; what is dumped here is the texture
that path samples, so a frame can be inspected or diffed without a display.
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png(path, rgba8: np.ndarray, level: int = 6) -> None:
    """rgba8: (H, W) uint32 packed R | G << 8 | B << 16 | A << 24 (what `se_sim_download_color` returns), or
    (H, W, 4) uint8.  8-bit RGBA, non-interlaced, filter type 0 on every row."""
    a = np.asarray(rgba8)
    if a.dtype == np.uint32 and a.ndim == 2:
        a = np.ascontiguousarray(a).view(np.uint8).reshape(a.shape[0], a.shape[1], 4)   # little-endian: bytes are R, G, B, A
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("expected (H, W) uint32 packed RGBA8 or (H, W, 4) uint8")
    h, w = a.shape[:2]
    raw = np.empty((h, 1 + 4 * w), np.uint8)
    raw[:, 0] = 0
    raw[:, 1:] = a.reshape(h, 4 * w)
    png = (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
           + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b""))
    Path(path).write_bytes(png)


def read_png_rgba8(path) -> np.ndarray:
    """Inverse of write_png for the files it writes (8-bit RGBA, filter 0): -> (H, W, 4) uint8.  For tests."""
    b = Path(path).read_bytes()
    if b[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG")
    pos, idat, w = 8, b"", None
    while pos < len(b):
        (n,), tag = struct.unpack(">I", b[pos:pos + 4]), b[pos + 4:pos + 8]
        data = b[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])
        if crc != (zlib.crc32(tag + data) & 0xFFFFFFFF):
            raise ValueError("bad CRC")
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", data)
            if (depth, ctype, interlace) != (8, 6, 0):
                raise ValueError("only 8-bit RGBA, non-interlaced")
        elif tag == b"IDAT":
            idat += data
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    if raw[:, 0].any():
        raise ValueError("only filter type 0")
    return raw[:, 1:].reshape(h, w, 4).copy()
