"""the data formats either side of training -- IAM-OnDB stroke XML / transcriptions in, stroke arrays
out, and the checkpoint files train.py writes.  Bit-exact against the reference's own functions (fixtures from
tests/golden/make_golden_iam.py, which imports /root/reference).  CPU only."""
import os

import numpy as np
import pytest
import torch

from dhg_b200.checkpoint import save_checkpoint, save_model_final
from dhg_b200.iam import combine_strokes, pad_img, pad_stroke_seq, parse_lines_txt, parse_strokes_xml
from dhg_b200.writer import read_state_dict

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "iam_pipeline.npz"))


def test_parse_strokes_xml_is_bit_exact(g):
    strokes = parse_strokes_xml(os.path.join(GOLDEN, "iam_line.xml"))
    assert strokes.dtype == np.float64 and strokes.shape == g["strokes"].shape
    assert np.array_equal(strokes, g["strokes"])
    assert set(np.unique(strokes[:, 2])) <= {0.0, 1.0}


def test_combine_strokes_is_bit_exact(g):
    out = combine_strokes(g["combine_in"].copy(), int(g["combine_n"]))
    assert np.array_equal(out, g["combine_out"])
    assert len(out) == len(g["combine_in"]) - int(g["combine_n"])
    assert abs(np.std(out[:, :2]) - 1.0) < 1e-12


def test_pad_stroke_seq_and_pad_img(g):
    padded = pad_stroke_seq(g["strokes"], 480)
    assert padded.dtype == np.float32 and np.array_equal(padded, g["padded"])
    assert np.all(padded[len(g["strokes"]):] == np.array([0, 0, 1], np.float32))
    assert pad_stroke_seq(g["strokes"], 10) is None and bool(g["too_long_is_none"])       # longer than max_seq_len
    assert pad_stroke_seq(g["strokes"] * 100, 2000) is None and bool(g["wild_is_none"])   # |value| > 15
    img = pad_img(g["img"], 1400, 96)
    assert img.dtype == np.float32 and np.array_equal(img, g["img_padded"])


def test_parse_lines_txt(g):
    texts = parse_lines_txt(os.path.join(GOLDEN, "iam_ascii.txt"))
    assert list(texts.keys()) == [str(k) for k in g["text_keys"]]
    assert list(texts.values()) == [str(v) for v in g["text_values"]]
    assert "iam_ascii-04" not in texts          # the blank line keeps its number but is dropped


def test_xml_without_strokeset(tmp_path):
    p = tmp_path / "empty.xml"
    p.write_text("<WhiteboardCaptureSession></WhiteboardCaptureSession>")
    with pytest.raises(ValueError, match="StrokeSet"):
        parse_strokes_xml(str(p))


def test_checkpoint_files_round_trip(tmp_path, state_dict):
    """train.py:123-137 writes checkpoint_<n>.pth = {meta, state_dict[, optimizer]} and model_final.pth = raw state_dict;
    both must come back through the loader the sampling path uses (checkpoint.py:117-129 semantics)."""
    lin = torch.nn.Linear(3, 2)
    opt = torch.optim.Adam(lin.parameters(), lr=1e-3)
    lin(torch.ones(1, 3)).sum().backward()
    opt.step()
    save_checkpoint(state_dict, tmp_path / "checkpoint_100.pth", meta={"step": 100}, optimizer=opt)
    ck = torch.load(tmp_path / "checkpoint_100.pth", weights_only=False)
    assert set(ck) == {"meta", "state_dict", "optimizer"} and ck["meta"] == {"step": 100}
    assert ck["optimizer"]["state"][0]["step"] == 1
    back = read_state_dict(str(tmp_path / "checkpoint_100.pth"))
    assert list(back) == list(state_dict) and all(torch.equal(back[k], state_dict[k]) for k in back)

    class Wrapped(torch.nn.Module):   # DataParallel-style wrapper: saved without the `module.` prefix
        def __init__(self):
            super().__init__()
            self.module = lin

    save_checkpoint(Wrapped(), tmp_path / "w.pth")
    ck = torch.load(tmp_path / "w.pth", weights_only=False)
    assert set(ck) == {"meta", "state_dict"} and set(ck["state_dict"]) == {"weight", "bias"}
    save_model_final(state_dict, tmp_path / "model_final.pth")
    raw = torch.load(tmp_path / "model_final.pth", weights_only=False)
    assert "state_dict" not in raw and len(raw) == 323
    assert all(torch.equal(read_state_dict(str(tmp_path / "model_final.pth"))[k], state_dict[k]) for k in state_dict)
