"""Stroke rasteriser standing in for utils/vis.py:5-36 (matplotlib is not
installed): cumulative sum of the offsets, split at pen lifts (p rounded, i.e.
p > 0.5), polylines drawn into an 8-bit image, written as PNG with zlib."""
import struct
import zlib

import numpy as np


def strokes_to_polylines(strokes):
    pos = np.cumsum(strokes[:, :2], axis=0)
    lifts = np.round(strokes[:, 2])
    lines, prev = [], 0
    for i, up in enumerate(lifts):
        if up:
            lines.append(pos[prev:i])  # the move INTO point i is pen-up (vis.py:24-32)
            prev = i
    return [l for l in lines if len(l) > 0]


def save_strokes_png(strokes, path, height=128):
    lines = strokes_to_polylines(np.asarray(strokes, dtype=np.float64))
    pts = np.concatenate(lines) if lines else np.zeros((1, 2))
    lo, hi = pts.min(0), pts.max(0)
    span = np.maximum(hi - lo, 1e-6)
    scale = (height - 8) / span[1]
    width = int(min(max(span[0] * scale + 8, 16), 8192))
    img = np.full((height, width), 255, np.uint8)
    for l in lines:
        p = (l - lo) * scale + 4
        for a, b in zip(p[:-1], p[1:]):
            n = int(max(abs(b - a).max(), 1)) + 1
            xs = np.clip(np.linspace(a[0], b[0], n).round().astype(int), 0, width - 1)
            ys = np.clip(np.linspace(a[1], b[1], n).round().astype(int), 0, height - 1)
            img[ys, xs] = 0

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    raw = b"".join(b"\x00" + img[r].tobytes() for r in range(height))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, 8, 0, 0, 0, 0))
    png += chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)
    return path
