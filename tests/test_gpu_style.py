"""StyleExtractor on the GPU against torchvision's MobileNetV2 feature stack run by PyTorch on the CPU
in fp32 -- the network the reference wraps (text_style.py:11-59).  The pretrained weights cannot be downloaded here, so
the model is seeded random-init with randomised BatchNorm statistics (which exercises the BN folding); the forward is the
reference's: x / 127.5 - 1, repeat to 3 channels, features, AvgPool2d(3, 3), AdaptiveAvgPool2d((1, 14)), permute."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mobilenet(seed=0):
    from torchvision import models

    torch.manual_seed(seed)
    m = models.mobilenet_v2(weights=None)
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():   # non-trivial BatchNorm: random affine and running statistics
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data = 0.5 + torch.rand(mod.weight.shape, generator=g)
            mod.bias.data = 0.2 * torch.randn(mod.bias.shape, generator=g)
            mod.running_mean = 0.2 * torch.randn(mod.running_mean.shape, generator=g)
            mod.running_var = 0.5 + torch.rand(mod.running_var.shape, generator=g)
    return m.eval()


def _reference_forward(m, img_batch):
    """text_style.py:49-59 around the given torchvision model."""
    with torch.no_grad():
        x = torch.tensor(img_batch, dtype=torch.float32)
        x = (x / 127.5) - 1
        x = x.repeat(1, 3, 1, 1)
        x = m.features(x)
        x = torch.nn.AvgPool2d(kernel_size=3, stride=3)(x)
        x = torch.nn.AdaptiveAvgPool2d((1, 14))(x)
        return x.squeeze(2).permute(0, 2, 1)


@pytest.mark.parametrize("B,H,W", [(1, 96, 1400), (2, 96, 1000), (1, 96, 1353), (1, 96, 300), (1, 128, 777)])
def test_style_extractor_matches_torchvision(B, H, W):
    from dhg_b200 import StyleExtractor

    m = _mobilenet(3)
    img = np.random.RandomState(B + W).randint(0, 256, size=(B, 1, H, W)).astype(np.float32)
    ref = _reference_forward(m, img)
    ex = StyleExtractor(m.state_dict())
    got = ex(img).cpu()
    ex.close()
    assert got.shape == (B, 14, 1280) and torch.isfinite(got).all()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 1e-4, rel
    assert (got - ref).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())


def test_style_extractor_errors():
    from dhg_b200 import StyleExtractor
    from dhg_b200._abi import DhgError

    m = _mobilenet(1)
    sd = m.state_dict()
    bad = {k: v for k, v in sd.items() if k != "features.7.conv.1.1.running_var"}
    with pytest.raises(DhgError, match="missing key"):
        StyleExtractor(bad)
    ex = StyleExtractor(sd)
    with pytest.raises(DhgError, match="at least"):
        ex(np.zeros((1, 1, 32, 500), np.float32))
    with pytest.raises(ValueError):
        ex(np.zeros((1, 3, 96, 500), np.float32))
    ex.close()


def test_infer_from_a_writer_image(tmp_path, state_dict):
    """infer(prompt, source=<image>) end to end: read_img -> StyleExtractor -> 60-step chain -> PNG (inference.py:60-98)."""
    import os
    import shutil

    import cv2

    from dhg_b200 import infer

    exp = tmp_path / "exp"
    exp.mkdir()
    shutil.copy(os.path.join(os.path.dirname(__file__), "golden", "config.yml"), exp / "config.yml")
    torch.save(state_dict, exp / "model_final.pth")
    torch.save(_mobilenet(2).state_dict(), tmp_path / "mobilenet.pth")
    rs = np.random.RandomState(0)
    page = np.full((200, 1800), 255, np.uint8)
    for _ in range(60):   # dark scribbles on a white page with white margins
        x, y = rs.randint(100, 1700), rs.randint(40, 160)
        cv2.line(page, (x, y), (x + rs.randint(-40, 40), y + rs.randint(-30, 30)), 0, 2)
    cv2.imwrite(str(tmp_path / "writer.png"), page)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        strokes = infer("Hello", str(tmp_path / "writer.png"), experiment_path=str(exp), output="pred", seed=1, dtype="bf16",
                        style_weights=str(tmp_path / "mobilenet.pth"))
    finally:
        os.chdir(cwd)
    assert strokes.shape == (16 * 6 + 8, 3) and torch.isfinite(strokes).all()
    assert (tmp_path / "pred.png").exists()
