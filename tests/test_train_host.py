"""Host side of the update half of the training step (dhg_b200/train.py): the learning-rate schedule of
scheduler.py:22-35, the flat parameter layout, and the one collective of a data-parallel step over gloo (world size 2)."""
import os
import socket
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion-handwriting-generation.pytorch_b200"))


def test_inv_sqrt_schedule_matches_scheduler_py():
    from dhg_b200.train import InvSqrtSchedule

    s = InvSqrtSchedule(1.0, 256, 10000)   # train.py:150-155 with config.yml: channels 128 -> d_model 256, warmup 10000
    for n in (1, 2, 9999, 10000, 10001, 60000):
        want = 1.0 * (256 ** -0.5) * min(n ** (-0.5), n * 10000 ** (-1.5))
        assert s.lr(n) == want
    assert s.lr(10000) == max(s.lr(n) for n in (1, 5000, 10000, 20000))   # the peak is at the end of the warm-up
    assert abs(s.lr(10000) - 6.25e-4) < 1e-12
    with pytest.raises(ValueError):
        s.lr(0)


def test_flatten_params_round_trip():
    from dhg_b200.train import flatten_params

    g = torch.Generator().manual_seed(0)
    ts = [torch.randn(3, 4, 5, generator=g), torch.randn(7, generator=g), torch.randn(1, 1, generator=g)]
    flat, table = flatten_params(ts, "cpu")
    assert flat.numel() == 60 + 7 + 1 and [o for o, _ in table] == [0, 60, 67]
    for (o, shp), t in zip(table, ts):
        assert torch.equal(flat[o:o + t.numel()].view(shp), t)


def _rank_main(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dhg_b200.train import exchange_gradients

    g = torch.full((1000,), float(rank + 1))
    n = exchange_gradients(g)
    out.put((rank, n, g[0].item(), g[-1].item()))
    dist.destroy_process_group()


def test_exchange_gradients_sums_over_two_ranks():
    """one all-reduce (sum) per step; the division by N is left to the optimiser kernel (world_size argument)."""
    import torch.multiprocessing as mp

    from dhg_b200.train import exchange_gradients

    assert exchange_gradients(torch.ones(4)) == 1   # no process group: nothing to exchange
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, 2, 3.0, 3.0), (1, 2, 3.0, 3.0)]


def test_param_layout_is_the_checkpoint_and_needs_no_gpu():
    """dhg_trainer_param_info through the product library (host-only call): 323 keys in model_final.pth order, back to back."""
    from dhg_b200.train import param_layout
    from oracle.dhg_oracle import state_dict_spec

    lay = param_layout(2, 128)
    spec = state_dict_spec(2, 128)
    assert list(lay) == [k for k, _ in spec] and len(lay) == 323
    off = 0
    for k, shape in spec:
        n = int(torch.Size(shape).numel())
        assert lay[k] == (off, n)
        off += n
    assert off == 10_028_451


def test_trainer_has_no_cpu_path():
    from dhg_b200.train import DenoiserTrainer, DhgTrainError
    from oracle.dhg_oracle import init_state_dict

    with pytest.raises(DhgTrainError, match="no CPU path"):
        DenoiserTrainer(init_state_dict(0), 2, 16, 5, device="cpu")


def test_get_alphas_follows_the_reference_draws():
    """utils/nn.py:42-61: same two draws from the same generator state give the same alphas (the reference uses the global one)."""
    from dhg_b200.diffusion import get_alpha_bar
    from dhg_b200.train import get_alphas

    alpha_set = torch.as_tensor(get_alpha_bar())
    g = torch.Generator().manual_seed(3)
    a = get_alphas(16, alpha_set, generator=g)
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(low=0, high=len(alpha_set) - 1, size=(16, 1), dtype=torch.int64, generator=g)
    want = torch.rand(16, 1, generator=g) * (alpha_set[idx + 1] - alpha_set[idx]) + alpha_set[idx]
    assert a.shape == (16, 1) and torch.equal(a, want)
    assert (a <= alpha_set[0]).all() and (a >= alpha_set[-1]).all()


class _StubTrainer:
    """Stands in for DenoiserTrainer in the loop test (the real one needs a GPU): records what fit() feeds it."""

    def __init__(self, interrupt_at=None):
        self.calls, self.interrupt_at = [], interrupt_at
        self.w = torch.zeros(3)

    def train_step(self, x, pen_lifts, text, style, alphas, eps, style_keep=None):
        if self.interrupt_at is not None and len(self.calls) + 1 == self.interrupt_at:
            raise KeyboardInterrupt
        self.calls.append((x.shape, pen_lifts.shape, text.shape, style.shape, alphas.shape, eps.shape, style_keep.shape))
        assert set(torch.unique(style_keep).tolist()) <= {0.0, float(torch.tensor(1.0) / 0.7)}
        self.w += 1
        n = float(len(self.calls))
        return torch.tensor(n), torch.tensor(n / 2), torch.tensor(n / 4)

    def state_dict(self):
        return {"w": self.w}


def test_fit_follows_the_reference_loop_cadence(tmp_path):
    """train.py:95-137: log when (count + 1) % log_freq == 0, checkpoint_<count + 1>.pth when (count + 1) % save_freq == 0,
    model_final.pth after `steps` steps; the batch iterable is restarted when it runs out."""
    from dhg_b200.train import fit

    batches = [{"strokes": torch.randn(2, 16, 3), "text": torch.ones(2, 5, dtype=torch.int64), "style": torch.randn(2, 14, 1280)}] * 2
    tr = _StubTrainer()
    lines = []
    log = type("L", (), {"info": staticmethod(lines.append)})
    hist = fit(tr, batches, steps=7, exp_dir=str(tmp_path), log_freq=3, save_freq=4, logger=log, generator=torch.Generator().manual_seed(0))
    assert len(tr.calls) == 7 and tr.calls[0] == ((2, 16, 2), (2, 16), (2, 5), (2, 14, 1280), (2, 1), (2, 16, 2), (2, 14, 1280))
    assert [h[0] for h in hist] == [3, 6] and hist[0][1] == pytest.approx(1.5) and hist[1][1] == pytest.approx(4.0)   # means of steps 1-2, 3-5
    assert sorted(p.name for p in tmp_path.iterdir()) == ["checkpoint_4.pth", "checkpoint_8.pth", "model_final.pth"]
    ck = torch.load(tmp_path / "checkpoint_4.pth")
    assert set(ck) == {"meta", "state_dict"} and ck["state_dict"]["w"].tolist() == [3.0, 3.0, 3.0]    # saved after 3 steps
    assert torch.load(tmp_path / "model_final.pth")["w"].tolist() == [7.0, 7.0, 7.0]
    assert any(l.startswith("Step 3 | Loss: 1.500 | Score: 0.750 | Pen: 0.375") for l in lines)


def test_fit_saves_the_last_state_on_interrupt(tmp_path):
    from dhg_b200.train import fit

    batches = [{"strokes": torch.randn(2, 16, 3), "text": torch.ones(2, 5, dtype=torch.int64), "style": torch.randn(2, 14, 1280)}]
    fit(_StubTrainer(interrupt_at=3), batches, steps=10, exp_dir=str(tmp_path), log_freq=100, save_freq=100)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["checkpoint_last.pth", "model_last.pth"]
    assert torch.load(tmp_path / "model_last.pth")["w"].tolist() == [2.0, 2.0, 2.0]
