"""Golden vectors for the IAM stroke pipeline, produced by the UNMODIFIED reference functions
(utils/io.py, utils/preprocessing.py) on a synthetic IAM-OnDB-style stroke file and transcription file.

    python tests/golden/make_golden_iam.py        # needs /root/reference (build container only)

Writes tests/golden/iam_line.xml, iam_ascii.txt (the inputs) and iam_pipeline.npz (the reference's outputs)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")


def synthetic_xml(rs, n_strokes=23):
    """IAM-OnDB layout: <WhiteboardCaptureSession><StrokeSet><Stroke><Point x= y= time=/>...  Smooth pen trajectories."""
    lines = ['<?xml version="1.0" encoding="ISO-8859-1"?>', "<WhiteboardCaptureSession>", "  <WhiteboardDescription/>", "  <StrokeSet>"]
    x, y, t = 900.0, 1400.0, 0.0
    for s in range(n_strokes):
        lines.append('    <Stroke colour="black" start_time="%.2f" end_time="%.2f">' % (t, t + 1))
        n = int(rs.randint(6, 40))
        vx, vy = rs.randn(2) * 12
        for _ in range(n):
            vx, vy = 0.8 * vx + rs.randn() * 6 + 3, 0.8 * vy + rs.randn() * 6
            x, y, t = x + vx, y + vy, t + 0.01
            lines.append('      <Point x="%d" y="%d" time="%.2f"/>' % (round(x), round(y), t))
        lines.append("    </Stroke>")
        x, y = x + rs.randint(20, 90), y + rs.randint(-40, 40)
    lines += ["  </StrokeSet>", "</WhiteboardCaptureSession>"]
    return "\n".join(lines) + "\n"


ASCII = """OCR:

A MOVE to stop Mr. Gaitskell from
nominating any more Labour life Peers

CSR:

A MOVE to stop Mr. Gaitskell from
nominating any more Labour life Peers
is to be made at a meeting of Labour

M Ps tomorrow. Mr. Michael Foot has
"""


def main():
    from diffusion_handwriting_generation.utils.io import combine_strokes, parse_lines_txt, parse_strokes_xml
    from diffusion_handwriting_generation.utils.preprocessing import pad_img, pad_stroke_seq
    from pathlib import Path

    rs = np.random.RandomState(7)
    xml = os.path.join(HERE, "iam_line.xml")
    with open(xml, "w") as f:
        f.write(synthetic_xml(rs))
    txt = os.path.join(HERE, "iam_ascii.txt")
    with open(txt, "w") as f:
        f.write(ASCII)
    strokes = parse_strokes_xml(xml)
    raw = rs.randn(101, 3)
    raw[:, 2] = rs.rand(101) < 0.1
    combined = combine_strokes(raw.copy(), 17)
    padded = pad_stroke_seq(strokes, 480)
    too_long = pad_stroke_seq(strokes, 10)
    wild = pad_stroke_seq(strokes * 100, 2000)
    img = rs.randint(0, 255, size=(96, 300)).astype(np.uint8)
    texts = parse_lines_txt(Path(txt))
    np.savez_compressed(
        os.path.join(HERE, "iam_pipeline.npz"), strokes=strokes, combine_in=raw, combine_n=17, combine_out=combined, padded=padded,
        too_long_is_none=too_long is None, wild_is_none=wild is None, img=img, img_padded=pad_img(img, 1400, 96),
        text_keys=np.array(list(texts.keys())), text_values=np.array(list(texts.values())),
    )
    print("strokes", strokes.shape, "padded", padded.shape, "texts", texts)


if __name__ == "__main__":
    main()
