"""render.write_png: the headless read-out format, checked without a GPU."""
import subprocess
import sys

import numpy as np
import pytest

from sandengine_b200.render import read_png_rgba8, write_png


def test_png_roundtrip_packed_and_bytes(tmp_path):
    rng = np.random.default_rng(3)
    for (h, w) in [(1, 1), (7, 13), (64, 48)]:
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        packed = (rgba[..., 0].astype(np.uint32) | (rgba[..., 1].astype(np.uint32) << 8)
                  | (rgba[..., 2].astype(np.uint32) << 16) | (rgba[..., 3].astype(np.uint32) << 24))
        write_png(tmp_path / "a.png", packed)
        write_png(tmp_path / "b.png", rgba)
        assert (tmp_path / "a.png").read_bytes() == (tmp_path / "b.png").read_bytes()
        assert np.array_equal(read_png_rgba8(tmp_path / "a.png"), rgba)


def test_png_is_a_valid_png_for_an_independent_reader(tmp_path):
    """`file` / Python's own tooling is not available everywhere; check the structure the spec requires instead:
    signature, IHDR first with the right geometry, IEND last, CRCs valid (read_png_rgba8 verifies every CRC)."""
    img = np.zeros((5, 9, 4), np.uint8)
    img[..., 0] = 255; img[..., 3] = 255
    write_png(tmp_path / "red.png", img)
    b = (tmp_path / "red.png").read_bytes()
    assert b[:8] == b"\x89PNG\r\n\x1a\n" and b[12:16] == b"IHDR" and b[-8:-4] == b"IEND"
    assert int.from_bytes(b[16:20], "big") == 9 and int.from_bytes(b[20:24], "big") == 5
    assert np.array_equal(read_png_rgba8(tmp_path / "red.png"), img)


def test_rejects_wrong_shapes(tmp_path):
    with pytest.raises(ValueError):
        write_png(tmp_path / "x.png", np.zeros((4, 4), np.float32))
    with pytest.raises(ValueError):
        write_png(tmp_path / "x.png", np.zeros((4, 4, 3), np.uint8))
