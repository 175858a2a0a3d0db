"""Behaviour of the C ABI entry points that the parity tests do not pin down: stream ordering between entry points that
share the plan's buffers, device-side input errors on the stream-ordered path, argument validation of the
host-buffer entry point, results independent of how a global batch is sharded, and the CLI.  Needs a B200."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(B, T, L, seed=11):
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(2, 73, (B, L), generator=g)
    text[:, -1] = 1
    return text, torch.randn(B, 14, 1280, generator=g), torch.randn(B, T, 2, generator=g), torch.randn(60, B, T, 2, generator=g)


@pytest.fixture(scope="module")
def writer(state_dict):
    from dhg_b200 import DiffusionWriter

    w = DiffusionWriter(state_dict=state_dict, num_layers=2, channels=128, dtype="bf16")
    yield w
    w.close()


def test_entry_points_on_different_streams_are_ordered(writer):
    """sample() on a side stream, sample_host() (the plan's private stream) right behind it, sample() on a third stream:
    all three share the plan's buffers and graph and must behave as if serialised, with no synchronisation in between."""
    dev = writer.device
    text, style, x0, noise = _inputs(16, 64, 10)
    text2, style2, x02, noise2 = _inputs(16, 64, 10, seed=12)
    ref1 = writer.sample(text, style, x0=x0, noise=noise).cpu()
    ref2 = writer.sample(text2, style2, x0=x02, noise=noise2).cpu()
    d1 = [t.to(dev) for t in (text, style, x0, noise)]
    d2 = [t.to(dev) for t in (text2, style2, x02, noise2)]
    torch.cuda.synchronize(dev)
    s1, s3 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for _ in range(3):
        with torch.cuda.stream(s1):
            a = writer.sample(d1[0], d1[1], x0=d1[2], noise=d1[3])
        b = writer.sample_host(text2, style2, x02, noise2)      # no sync since the call above
        with torch.cuda.stream(s3):
            c = writer.sample(d1[0], d1[1], x0=d1[2], noise=d1[3])
        with torch.cuda.stream(s1):
            d = writer.sample(d2[0], d2[1], x0=d2[2], noise=d2[3])
        torch.cuda.synchronize(dev)
        assert torch.equal(a.cpu(), ref1) and torch.equal(b, ref2) and torch.equal(c.cpu(), ref1) and torch.equal(d.cpu(), ref2)


def test_device_side_token_error_is_reported(writer):
    """A CUDA text tensor modified behind the Python guard's back (an alias write does not bump the version the guard
    remembers) reaches the kernel: the stream-ordered call cannot fail, dhg_check_errors does."""
    dev = writer.device
    text, style, x0, noise = _inputs(4, 64, 10)
    t_dev = text.to(dev)
    writer.sample(t_dev, style, x0=x0, noise=noise)          # passes the guard; remembered as checked
    t_dev.data[1, 2] = 99                                     # .data write: no version bump
    writer.sample(t_dev, style, x0=x0, noise=noise)
    with pytest.raises(IndexError, match="token id"):
        writer.check_errors()
    writer.check_errors()                                     # the flag was cleared
    # the host-buffer entry point validates before anything is copied
    bad = text.clone()
    bad[0, 0] = -1
    with pytest.raises(IndexError):
        writer.sample_host(bad, style, x0, noise)


def test_sample_host_validates_like_sample(writer):
    text, style, x0, noise = _inputs(4, 64, 10)
    for args in ((text, style[:1], x0, noise), (text, style[0], x0, noise), (text, style[..., :1279], x0, noise),
                 (text, style, x0, noise[:59]), (text, style, x0[:2], noise)):
        with pytest.raises(ValueError):
            writer.sample_host(*args)
    with pytest.raises(ValueError):
        writer.sample_host(text, style, x0, noise, diffusion_mode="ddim")
    from dhg_b200._abi import DhgError

    with pytest.raises((DhgError, ValueError)):
        writer.sample(text, style, x0=x0, noise=None, T=60)   # T not a multiple of 8


@pytest.mark.parametrize("dtype,G,T,L,worlds", [("bf16", 16, 64, 10, (1, 2, 4, 8, 3)), ("fp32", 16, 64, 10, (1, 2, 4, 8, 3)),
                                                 # BASELINE length: 196 keys at level 1, where two attention kernels with different
                                                 # summation orders apply (the plan-time timing must not choose between them)
                                                 ("bf16", 6, 392, 24, (1, 2, 3))])
def test_result_does_not_depend_on_the_sharding(state_dict, dtype, G, T, L, worlds):
    """'results independent of N'.  A G-prompt global batch through sharding.sample_sharded as 1, 2, 4 and 8
    'ranks' (each rank's slice is its own plan of its own batch size, exactly what a rank of an N-GPU job runs): the
    concatenated result has the same bits every time."""
    from dhg_b200 import DiffusionWriter
    from dhg_b200.sharding import sample_sharded, shard_bounds

    text, style, x0, noise = _inputs(G, T, L, seed=21)
    results = []
    for world in worlds:
        parts = []
        for rank in range(world):
            lo, hi = shard_bounds(G, rank, world)
            w = DiffusionWriter(state_dict=state_dict, num_layers=2, channels=128, dtype=dtype, chunk=hi - lo)
            parts.append(sample_sharded(lambda t, s, x0, noise: w.sample(t, s, x0=x0, noise=noise), text, style, x0, noise,
                                        rank, world, gather=False).cpu())
            w.close()
        results.append(torch.cat(parts))
    for r in results[1:]:
        assert torch.equal(r, results[0])


def test_cli_end_to_end(tmp_path, state_dict, golden):
    """python -m dhg_b200.inference with the reference's `make infer` arguments (Makefile:14-21)."""
    import shutil
    import subprocess
    import sys

    exp = tmp_path / "exp"
    exp.mkdir()
    shutil.copy(os.path.join(os.path.dirname(__file__), "golden", "config.yml"), exp / "config.yml")
    torch.save(state_dict, exp / "model_final.pth")
    torch.save(torch.tensor(golden("chain_c1")["style"][0]), tmp_path / "style.pt")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "diffusion-handwriting-generation.pytorch_b200"))
    r = subprocess.run([sys.executable, "-m", "dhg_b200.inference", "--prompt=Hello World and goodbye", f"--source={tmp_path / 'style.pt'}",
                        f"--experiment_path={exp}", "--config_path=", "--checkpoint_path=", "--output=prediction", "--seed=3"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (tmp_path / "prediction.png").exists() and "stroke points" in r.stdout


@pytest.mark.parametrize("chunk", [5, 2])
def test_host_entry_point_direct_and_staged_paths_give_the_device_bits(state_dict, chunk):
    """dhg_sample_host: a batch that fills the plan goes straight into the plan's buffers with the noise of the later
    steps copied while the first steps run (two graphs); any other batch is staged and chunked.  Both must return the
    bits of the device-resident call, twice in a row (the second call re-uses graphs and buffers)."""
    from dhg_b200 import DiffusionWriter

    text, style, x0, noise = _inputs(5, 64, 10, seed=33)
    w = DiffusionWriter(state_dict=state_dict, num_layers=2, channels=128, dtype="bf16", chunk=chunk)
    ref = w.sample(text.cuda(), style.cuda(), x0=x0.cuda(), noise=noise.cuda()).cpu()
    for _ in range(2):
        got = w.sample_host(text, style, x0.pin_memory(), noise.pin_memory())
        assert torch.equal(got, ref)
    w.close()
