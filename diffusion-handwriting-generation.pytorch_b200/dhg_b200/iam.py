"""The on-disk formats either side of training: IAM-OnDB stroke XML and transcription files in, stroke
arrays out -- host-side numpy, same results as the reference's

    utils/io.py:11-66            parse_strokes_xml      (point deltas, pen-lift channel shifted by one, 3 x simplification)
    utils/io.py:69-95            parse_lines_txt        (the text under 'CSR:' of an ascii/*.txt file)
    utils/io.py:118-147          combine_strokes        (merge the n most collinear neighbouring offsets)
    utils/preprocessing.py:4-26  pad_stroke_seq         (pad to max_seq_len with (0, 0, 1); reject long / wild lines)
    utils/preprocessing.py:29-44 pad_img                (white padding on the right)

These run once per dataset line on the CPU (a few hundred points); nothing here is on the sampling path.  Parity is
pinned by tests/golden/iam_pipeline.npz, produced by the reference's own functions (tests/golden/make_golden.py).
"""
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np


def combine_strokes(x, n):
    """Sum the `n` pairs (2i, 2i+1) of consecutive offsets that deviate least from a straight line (smallest
    |a| + |b| - |a + b|); the merged offset keeps a pen lift if either part had one.  Re-normalises to unit std."""
    pairs = len(x) // 2
    a, b = x[0:2 * pairs:2, :2], x[1:2 * pairs:2, :2]
    detour = np.linalg.norm(a, axis=1) + np.linalg.norm(b, axis=1) - np.linalg.norm(a + b, axis=1)
    first = np.argsort(detour)[:n] * 2
    x[first] += x[first + 1]
    x[first, 2] = x[first, 2] > 0
    x = np.delete(x, first + 1, axis=0)
    x[:, :2] /= np.std(x[:, :2])
    return x


def parse_strokes_xml(xml_path):
    """-> float64 [num_points - 1, 3] = (dx, dy, pen_lift): offsets between consecutive points in file order (y flipped),
    the end-of-stroke flag shifted by one position so that it marks the offset that is NOT drawn, unit-std coordinates,
    then three rounds of `combine_strokes` on 20 % of the points."""
    stroke_set = ET.parse(xml_path).getroot().find("StrokeSet")
    if stroke_set is None:
        raise ValueError("No StrokeSet element found in XML file")
    pts, ends = [], []
    for stroke in stroke_set.findall("Stroke"):
        points = stroke.findall("Point")
        for k, p in enumerate(points):
            pts.append((int(p.attrib["x"]), -int(p.attrib["y"])))
            ends.append(1.0 if k == len(points) - 1 else 0.0)
    pts = np.asarray(pts, dtype=float).reshape(-1, 2)
    out = np.empty((max(len(pts) - 1, 0), 3), dtype=float)
    out[:, :2] = pts[1:] - pts[:-1]
    out[:, 2] = np.roll(np.asarray(ends[1:], dtype=float), 1)
    out[:, :2] /= np.std(out[:, :2])
    for _ in range(3):
        out = combine_strokes(out, int(len(out) * 0.2))
    return out


def parse_lines_txt(ascii_file):
    """{'<file stem>-<line number:02d>': text} for the lines after the 'CSR:' marker (and the blank line that follows it)."""
    ascii_file = Path(ascii_file)
    texts, seen_marker, index = {}, False, -1
    with ascii_file.open("r") as f:
        for line in f.readlines():
            seen_marker = seen_marker or "CSR" in line
            if not seen_marker:
                continue
            if index > 0 and line.strip():
                texts[f"{ascii_file.stem}-{index:02d}"] = line[:-1]
            index += 1
    return texts


def pad_stroke_seq(x, maxlength):
    """float32 [maxlength, 3]: `x` followed by (0, 0, 1) rows; None if the line is longer than `maxlength` or any value
    exceeds 15 in magnitude."""
    if len(x) > maxlength or np.amax(np.abs(x)) > 15:
        return None
    pad = np.zeros((maxlength - len(x), 3))
    pad[:, 2] = 1
    return np.concatenate((x, pad)).astype("float32")


def pad_img(img, width, height):
    """float32 [height, width]: `img` with white (255) columns appended on the right."""
    return np.concatenate((img, np.full((height, width - img.shape[1]), 255.0)), axis=1).astype("float32")
