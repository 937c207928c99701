"""Synthetic initial grids: a counter-based generator so that C/Python/CUDA agree
without an RNG library.  cell(x, y) is chosen by thresholds on hashi(x*461 + y*2131 + seed*2131^2) >> 8,
hashi = the shader's lowbias32 (math.glsl:17-25).  The simulation itself has no seed: "fixed
seed" only ever means the seed of this generator plus the starting frame.
"""
from __future__ import annotations

import numpy as np

# (material name, fraction) -- default mix of SURVEY.md 8d config 1
DEFAULT_MIX = (("EMPTY", 0.45), ("sand", 0.15), ("water", 0.15), ("dirt", 0.08), ("rock", 0.07), ("smoke", 0.05),
               ("toxic_sludge", 0.03), ("radioactive", 0.02))
# ids of those names in data/materials.yaml (EMPTY 0, sand 3, rock 4, water 5, radioactive 6, smoke 7, toxic_sludge 8, dirt 10)
DEFAULT_IDS = {"EMPTY": 0, "sand": 3, "rock": 4, "water": 5, "radioactive": 6, "smoke": 7, "toxic_sludge": 8, "vine": 9, "dirt": 10}


def hashi(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def thresholds(mix, ids):
    """Cumulative 24-bit thresholds and the id of each bucket; the last bucket absorbs rounding."""
    cum, acc = [], 0.0
    for _, frac in mix:
        acc += frac
        cum.append(min(int(round(acc * (1 << 24))), 1 << 24))
    cum[-1] = 1 << 24
    return np.array(cum, dtype=np.uint32), np.array([ids[name] for name, _ in mix], dtype=np.uint32)


def synthetic_grid(width: int, height: int, seed: int, mix=DEFAULT_MIX, ids=DEFAULT_IDS, row_begin: int = 0, row_end: int | None = None,
                   out: np.ndarray | None = None) -> np.ndarray:
    """Rows [row_begin, row_end) of the width x height grid for `seed` as packed uint32 material ids."""
    row_end = height if row_end is None else row_end
    rows = row_end - row_begin
    if out is None:
        out = np.empty((rows, width), np.uint32)
    cum, bucket_ids = thresholds(mix, ids)
    xs = (np.arange(width, dtype=np.uint64) * 461).astype(np.uint32)
    sterm = np.uint32((seed * 2131 * 2131) & 0xFFFFFFFF)
    chunk = max(1, (1 << 22) // max(width, 1))
    with np.errstate(over="ignore"):
        for r0 in range(0, rows, chunk):
            r1 = min(rows, r0 + chunk)
            ys = ((np.arange(row_begin + r0, row_begin + r1, dtype=np.uint64) * 2131) & 0xFFFFFFFF).astype(np.uint32)
            h = hashi(xs[None, :] + ys[:, None] + sterm) >> np.uint32(8)
            out[r0:r1] = bucket_ids[np.searchsorted(cum, h, side="right")]
    return out


def kat_grid(n: int) -> np.ndarray:
    """The survey's state known-answer grid: g[y][x] = pal[(7x + 13y + (x*y mod 5)) mod 8]."""
    pal = np.array([0, 3, 5, 7, 4, 10, 0, 0], dtype=np.uint32)
    y, x = np.mgrid[0:n, 0:n]
    return pal[(7 * x + 13 * y + ((x * y) % 5)) % 8]


def grid_checksum(cells: np.ndarray, width: int | None = None, row_begin: int = 0) -> int:
    """numpy twin of se_sim_checksum (csrc/static_kernels.cu): sum over cells of mix(global index, id) mod 2^64."""
    cells = np.asarray(cells, dtype=np.uint32)
    w = int(width if width is not None else cells.shape[1])
    total = 0
    with np.errstate(over="ignore"):
        for r0 in range(0, cells.shape[0], 1024):            # bounded temporaries
            blk = cells[r0:r0 + 1024].astype(np.uint64).ravel()
            idx = np.uint64((row_begin + r0) * w) + np.arange(blk.size, dtype=np.uint64)
            h = idx * np.uint64(0x9E3779B97F4A7C15) + blk * np.uint64(0xD6E8FEB86659FD93)
            h ^= h >> np.uint64(32)
            h *= np.uint64(0xD6E8FEB86659FD93)
            h ^= h >> np.uint64(32)
            total = (total + int(h.sum(dtype=np.uint64))) & 0xFFFFFFFFFFFFFFFF
    return total
