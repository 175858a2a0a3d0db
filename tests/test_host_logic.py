"""Host-side mirror of the reference interface: tokenizer, config reader, checkpoint
discovery and formats, sharding arithmetic, stroke rasteriser.  CPU only."""
import os

import numpy as np
import pytest
import torch

from dhg_b200.config import DLConfig, parse_yaml
from dhg_b200.diffusion import get_alpha_bar, get_beta_set
from dhg_b200.inference import load_style, resolve_experiment
from dhg_b200.sharding import shard_bounds
from dhg_b200.tokenizer import Tokenizer, stroke_length
from dhg_b200.vis import save_strokes_png, strokes_to_polylines
from dhg_b200.writer import read_state_dict

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_tokenizer_matches_reference(golden):
    g = golden("schedule_tokenizer")
    tok = Tokenizer()
    assert tok.vocab_size == 73
    for i, prompt in enumerate(g["prompts"]):
        assert tok.encode(str(prompt)) == g[f"tok_{i}"].tolist()
    ids = tok.encode("Follow the White Rabbit")
    assert len(ids) == 24 and ids[-1] == 1 and stroke_length(len(ids)) == 392
    assert tok.decode(ids[:-1]) == "Follow the White Rabbit"
    assert tok.encode("") == [1]
    assert tok.encode("é~")[:2] == [2, 2]  # unknown -> '_' (id 2)


def test_stroke_length_rule():
    # inference.py:77-78: T = 16 n, then T - T % 8 + 8 (always adds 8 since 16 n % 8 == 0)
    for n in (1, 2, 24, 50, 81):
        assert stroke_length(n) == 16 * n + 8
        assert stroke_length(n) % 8 == 0


def test_schedule_matches_reference(golden):
    g = golden("schedule_tokenizer")
    assert np.array_equal(get_beta_set().numpy(), g["beta"])
    assert np.array_equal(get_alpha_bar().numpy(), g["alpha_bar"])


def test_config_reader():
    cfg = DLConfig.load(os.path.join(GOLDEN, "config.yml"))
    assert cfg.training_args.att_layers_num == 2
    assert cfg.training_args.channels == 128
    assert cfg.training_args.dropout == 0.0
    assert cfg.training_args.max_files is None          # empty value -> null
    assert cfg.training_args.not_there is None          # CfgDict: missing -> None
    assert cfg.dataset_args.max_seq_len == 480          # trailing comment stripped
    assert cfg.optimizer.params.betas == [0.9, 0.98]
    assert cfg.optimizer.type == "torch.optim.Adam"
    assert cfg.experiment.deterministic is False
    cfg.update({"training_args.att_layers_num": 4})
    assert cfg.training_args.att_layers_num == 4
    with pytest.raises(ValueError):
        parse_yaml("just a line without a colon")


def test_checkpoint_discovery_order(tmp_path):
    (tmp_path / "config.yml").write_text("training_args:\n  channels: 128\n")
    with pytest.raises(ValueError):
        resolve_experiment(experiment_path=str(tmp_path))        # no checkpoint at all
    with pytest.raises(ValueError):
        resolve_experiment()                                     # nothing given
    for name in ("checkpoint_1000.pth", "checkpoint_12000.pth", "checkpoint_last.pth"):
        (tmp_path / name).write_bytes(b"x")
    cfg, ck = resolve_experiment(experiment_path=str(tmp_path))
    assert ck.endswith("checkpoint_12000.pth") and cfg.endswith("config.yml")
    (tmp_path / "model_last.pth").write_bytes(b"x")
    assert resolve_experiment(experiment_path=str(tmp_path))[1].endswith("model_last.pth")
    (tmp_path / "model_final.pth").write_bytes(b"x")
    assert resolve_experiment(experiment_path=str(tmp_path))[1].endswith("model_final.pth")
    assert resolve_experiment("a.yml", "b.pth", str(tmp_path)) == ("a.yml", "b.pth")


def test_checkpoint_formats(tmp_path):
    sd = {"input_dense.weight": torch.randn(128, 2), "input_dense.bias": torch.randn(128)}
    torch.save(sd, tmp_path / "raw.pth")
    torch.save({"meta": {}, "state_dict": {"module." + k: v for k, v in sd.items()}}, tmp_path / "wrapped.pth")
    torch.save([1, 2, 3], tmp_path / "bad.pth")
    for f in ("raw.pth", "wrapped.pth"):
        got = read_state_dict(str(tmp_path / f))
        assert set(got) == set(sd) and torch.equal(got["input_dense.bias"], sd["input_dense.bias"])
    with pytest.raises(RuntimeError):
        read_state_dict(str(tmp_path / "bad.pth"))


def test_shard_bounds_cover_batch():
    for total in (0, 1, 7, 64, 8192, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(8192, 3, 8) == (3072, 4096)
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_style_loading(tmp_path):
    s = torch.randn(14, 1280)
    torch.save(s, tmp_path / "style.pt")
    np.save(tmp_path / "style.npy", s.numpy())
    assert load_style(str(tmp_path / "style.pt")).shape == (1, 14, 1280)
    assert torch.equal(load_style(str(tmp_path / "style.npy"))[0], s)


def test_read_img_and_remove_whitespace(tmp_path):
    """utils/preprocessing.py:47-62 and utils/io.py:98-115: crop to the dark rows / columns (the last one is cut too, as
    in the reference's rows[0]:rows[-1]), then bicubic resize to the requested height keeping the aspect ratio."""
    import cv2

    from dhg_b200.style import read_img, remove_whitespace

    img = np.full((50, 80), 255, np.uint8)
    img[10:30, 20:60] = 0
    img[12, 5] = 200          # lighter than the threshold: still whitespace
    out = remove_whitespace(img, thresh=127)
    assert out.shape == (19, 39) and out.max() == 0
    holes = img.copy()
    holes[15:20, :] = 255
    assert remove_whitespace(holes, 127, remove_middle=True).shape == (15, 40)
    cv2.imwrite(str(tmp_path / "w.png"), img)
    r = read_img(tmp_path / "w.png", 96)
    assert r.shape == (96, 96 * 39 // 19) and r.dtype == np.uint8
    with pytest.raises(FileNotFoundError):
        read_img(tmp_path / "nope.png", 96)
    with pytest.raises(ValueError, match="MobileNetV2 weights"):
        os.environ.pop("DHG_MOBILENET_WEIGHTS", None)
        from dhg_b200.style import _load_weights
        _load_weights(None)


def test_polylines_and_png(tmp_path):
    # pen lift at index 3: the move into point 3 is a jump (utils/vis.py:24-32)
    strokes = np.array([[1, 0, 0], [1, 1, 0], [1, 0, 0], [5, 5, 1], [1, 0, 0], [1, 1, 0.4]], dtype=np.float32)
    lines = strokes_to_polylines(strokes)
    assert len(lines) == 1 and len(lines[0]) == 3       # trailing open segment is not drawn, like the reference
    strokes[5, 2] = 0.6
    lines = strokes_to_polylines(strokes)
    assert [len(l) for l in lines] == [3, 2]
    p = save_strokes_png(strokes, str(tmp_path / "out.png"))
    data = open(p, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and data[-8:-4] == b"IEND"


def test_chain_shape_contract():
    """Every pointer handed to dhg_sample / dhg_sample_host is sized from these checks (sample_host used to
    skip them and could read past the end of a host buffer)."""
    from dhg_b200.writer import DiffusionWriter

    chk = DiffusionWriter._check_chain_shapes
    B, T, L = 3, 16, 5
    text, style = torch.zeros(B, L, dtype=torch.int64), torch.zeros(B, 14, 1280)
    x0, noise = torch.zeros(B, T, 2), torch.zeros(60, B, T, 2)
    assert chk(text, style, x0, noise) == (B, T, L, 14)
    bad = [
        (text, style[:1], x0, noise),                  # one style for B > 1 prompts
        (text, style[0], x0, noise),                   # 2-D style
        (text, torch.zeros(B, 14, 1279), x0, noise),   # wrong style width
        (text, style, x0[:, :, :1], noise),            # not (dx, dy)
        (text, style, x0, noise[:59]),                 # 59 draws
        (text, style, x0, noise[:, :2]),               # noise of another batch
        (text, style, torch.zeros(B, 12, 2), torch.zeros(60, B, 12, 2)),   # T not a multiple of 8
        (text[0], style, x0, noise),                   # 1-D text
    ]
    for args in bad:
        with pytest.raises(ValueError):
            chk(*args)
    assert DiffusionWriter._check_mode("new") == 0 and DiffusionWriter._check_mode("standard") == 1
    with pytest.raises(ValueError):
        DiffusionWriter._check_mode("ddim")


def test_cli_mirrors_make_infer(monkeypatch, capsys):
    """`make infer` of the reference passes --prompt= --source= --experiment_path= --config_path="" --checkpoint_path=""
    --output= (Makefile:14-21) to fire.Fire(infer) (inference.py:101-102)."""
    from dhg_b200 import inference

    seen = {}

    def fake_infer(prompt, source, config_path=None, checkpoint_path=None, experiment_path=None, output="result",
                   diffusion_mode="new", **kw):
        seen.update(prompt=prompt, source=source, config_path=config_path, checkpoint_path=checkpoint_path,
                    experiment_path=experiment_path, output=output, diffusion_mode=diffusion_mode, **kw)
        return torch.zeros(24, 3)

    monkeypatch.setattr(inference, "infer", fake_infer)
    rc = inference.main(["--prompt=Hello World and goodbye", "--source=style.pt", "--experiment_path=data/best_exp",
                         "--config_path=", "--checkpoint_path=", "--output=prediction"])
    assert rc == 0 and "prediction.png" in capsys.readouterr().out
    assert seen["prompt"] == "Hello World and goodbye" and seen["source"] == "style.pt"
    assert seen["experiment_path"] == "data/best_exp" and seen["output"] == "prediction"
    assert not seen["config_path"] and not seen["checkpoint_path"] and seen["diffusion_mode"] == "new"
    # fire also accepts the positional order of infer()
    inference.main(["Follow the White Rabbit", "style.npy", "--experiment_path", "exp", "--diffusion_mode", "standard"])
    assert seen["prompt"] == "Follow the White Rabbit" and seen["source"] == "style.npy" and seen["diffusion_mode"] == "standard"
    with pytest.raises(SystemExit):
        inference.main(["--prompt=only a prompt"])
