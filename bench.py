"""Benchmark of the reverse-diffusion sampling hot path (BASELINE.json metric:
sampled handwriting lines/s for the full 60-step chain; us per denoiser step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--global-batch G]

One "step" = one pass of the hot path over one batch: the full 60-step reverse
chain (inference.py:81-96) for the per-GPU batch of synthetic prompts.
Workload at N=1: 1024 prompts per GPU, T=392 stroke points, L=24 tokens (the
shape of BASELINE configs[0]/[1], 'Follow the White Rabbit'; 1024 = the per-GPU
shard of configs[2]'s 8192 at 8 GPUs).  Default: WEAK scaling (per-GPU batch
fixed at 1024, global batch 1024*N).  `--global-batch 8192` is configs[2] as
written: a fixed global batch sharded over the ranks (STRONG scaling).  Either
way rank r owns the contiguous slice `sharding.shard_bounds(G, r, N)` of
GLOBALLY indexed prompts (sample g's text, style, x0 and noise depend on g
only), and there is no collective in the loop.

Under torchrun (N>1) every rank drives its own GPU; the timed region is
bracketed by barrier + synchronize, timed with CUDA events, max over ranks.

Beside the headline the line carries (rank 0, N=1 unless noted):
  parity                 4 rows of the timed B=1024 result against the CPU oracle
  equivalence            sha1 of a 64-prompt global batch sampled through sharding.sample_sharded
                         over the N ranks (every N): identical for every N
  fp32                   the same chain in the fp32-contract mode (a like-precision figure)
  gpu_eager_baseline     the UNMODIFIED reference (baseline/_ref) in PyTorch eager on the same B200
  cpu_baseline           the unmodified reference on the host cores (bounded sample)
  other_configs          BASELINE configs[1] (B=64), B=1 latency, configs[4] shard (T=1200, L=81)
  roofline*              chain (tensor), every GEMM family of a step, the posterior-update kernel
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

T_STROKES, L_TEXT, S_STYLE = 392, 24, 14
NUM_STEPS = 60
SEED = 1234


def flops_alg_per_sample_step(T, L, S5=70):
    """F_alg of SURVEY.md 8d: work that depends on x_t or the step (hoisted invariants excluded)."""
    gemm = 3_446_016 * T + 4_292_608 * L + 589_824 * S5
    attn = 768 * (T // 2) * (L + T // 2) + 1024 * (T // 4) * (L + T // 4) + 3072 * (T // 8) * (L + T // 8) + 1536 * L * S5
    return gemm + attn


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def train_update_leg(dev, peaks):
    """SURVEY 8e / 8f-3, the part that exists: one gradient-norm + clipped Adam step over the reference's
    10,028,451 parameters (dhg_b200.train.FlatAdam -> dhg_train_sqnorm + dhg_train_adam_step), device-timed."""
    try:
        from dhg_b200.train import FlatAdam

        n = 10_028_451
        opt = FlatAdam([torch.zeros(n, device=dev)], clip_grad=100.0, device=dev)
        grad = torch.randn(n, device=dev)
        for _ in range(3):
            opt.step_and_update_lr(grad)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            opt.step_and_update_lr(grad)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        gbs = n * 32 / us / 1e3
        return {"parameters": n, "us_per_step": us, "bytes_per_step": n * 32, "achieved_gbs": gbs, "peak_gbs": peaks["hbm_gbs"],
                "frac": gbs / peaks["hbm_gbs"], "what": "gradient 2-norm + clip + Adam (inv-sqrt schedule) on the flat fp32 buffer; "
                "28 B (Adam) + 4 B (norm) per parameter; the gradient comes from dhg_trainer_backward (train_step leg)"}
    except Exception as e:   # noqa: BLE001
        return {"failed": str(e)[:200]}


def train_step_leg(dev, B=96, T=480, L=50, steps=5):
    """BASELINE configs[3] / SURVEY 8f-3: TrainingLoop.train_step (train.py:26-67) at batch 96, seq_len 480, text 50 on ONE GPU:
    dhg_b200.train.DenoiserTrainer (perturb, forward with kept activations, loss, backward, clip + Adam; csrc/train_step.cu +
    train_update.cu, fp32 on the CUDA cores) and, beside it, the UNMODIFIED reference's step in PyTorch eager fp32 on the same
    GPU (model.train(), loss_fn, backward, clip_grad_norm_, Adam in InvSqrtScheduledOptim).  Device-timed, synthetic batch."""
    out = {"config": {"B": B, "T": T, "L": L, "workload": "configs[3]: train.py denoiser step, seq_len 480, batch 96 (per GPU)"}}
    from oracle.dhg_oracle import alpha_bar, beta_schedule, init_state_dict

    g = torch.Generator().manual_seed(SEED)
    strokes = torch.randn(B, T, 2, generator=g).to(dev)
    pen = (torch.rand(B, T, generator=g) < 0.05).float().to(dev)
    text = torch.randint(2, 73, (B, L), generator=g)
    text[:, -1] = 1
    text[: B // 2, L // 2:] = 0     # half of the prompts padded (config.yml: max_text_len 50)
    text[: B // 2, L // 2 - 1] = 1
    text = text.to(dev)
    style = torch.randn(B, S_STYLE, 1280, generator=g).to(dev)
    keep = ((torch.rand(B, S_STYLE, 1280, generator=g) >= 0.3).float() / 0.7).to(dev)
    eps = torch.randn(B, T, 2, generator=g).to(dev)
    abar = alpha_bar(beta_schedule())
    idx = torch.randint(0, len(abar) - 1, (B, 1), generator=g)
    alphas = (torch.rand(B, 1, generator=g) * (abar[idx + 1] - abar[idx]) + abar[idx]).to(dev)
    sd = init_state_dict(0)

    def timed(fn):
        for _ in range(2):
            r = fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps, r

    try:
        from dhg_b200.train import DenoiserTrainer

        tr = DenoiserTrainer(sd, B, T, L, device=dev)
        ms, losses = timed(lambda: tr.train_step(strokes, pen, text, style, alphas, eps, style_keep=keep))
        flop = 3.0 * B * (flops_alg_per_sample_step(T, L) + 983_040 * 70 + 1_323_008)   # forward F_exec x 3 for fwd + bwd
        out["b200"] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "loss": float(losses[0].item()), "finite": bool(torch.isfinite(losses[0]).item()),
                       "launches_fwd_bwd": int(tr.last_launch_count), "workspace_GB": tr.workspace_bytes / 1e9,
                       "tflops_fp32_3x_forward": flop / (ms * 1e-3) / 1e12, "dtype": "fp32 (CUDA cores)",
                       "frac_of_fp32_fma_peak": flop / (ms * 1e-3) / 1e12 / 72.7, "fp32_fma_peak_tflops": 72.7}
        # the same step with plain TF32 products on the tensor cores (torch's allow_tf32 precision class, outside the fp32 contract)
        from dhg_b200 import _abi
        _abi.lib().dhg_trainer_set_option(b"tiled_gemm", 4)
        try:
            ms4, l4 = timed(lambda: tr.train_step(strokes, pen, text, style, alphas, eps, style_keep=keep))
            out["b200_tf32"] = {"ms_per_step": ms4, "samples_per_s": B / (ms4 * 1e-3), "loss": float(l4[0].item()),
                                "dtype": "TF32 products (mma.sync), fp32 accumulate and storage"}
        finally:
            _abi.lib().dhg_trainer_set_option(b"tiled_gemm", 1)
        tr.close()
        del tr
        torch.cuda.empty_cache()
    except Exception as e:   # noqa: BLE001
        out["b200"] = {"failed": f"{type(e).__name__}: {str(e)[:300]}"}
    try:
        mods = reference_modules()
        if mods is None:
            raise RuntimeError("baseline/_ref is missing")
        from diffusion_handwriting_generation.loss import loss_fn
        from diffusion_handwriting_generation.scheduler import InvSqrtScheduledOptim
        from diffusion_handwriting_generation.utils.clip_grad import dispatch_clip_grad

        model = mods[0](num_layers=2, c1=128, c2=192, c3=256, drop_rate=0.0)
        model.load_state_dict(sd, strict=True)
        model.to(dev).train()
        opt = InvSqrtScheduledOptim(torch.optim.Adam(model.parameters(), lr=3e-4, weight_decay=1e-5, betas=(0.9, 0.98)), 1.0, 256, 10000)

        def ref_step():   # train.py:38-63
            x_p = torch.sqrt(alphas).unsqueeze(-1) * strokes + torch.sqrt(1 - alphas).unsqueeze(-1) * eps
            opt.zero_grad()
            score_pred, pen_pred, _ = model(x_p, text, torch.sqrt(alphas), style)
            loss, _, _ = loss_fn(eps, score_pred, pen, pen_pred, alphas)
            loss.backward()
            dispatch_clip_grad(model.parameters(), value=100.0)
            opt.step_and_update_lr()
            return loss

        ms, loss = timed(ref_step)
        out["reference_eager_fp32"] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "loss": float(loss.item()),
                                       "flags": {"matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32),
                                                 "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32)}}
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            ms, loss = timed(ref_step)
            out["reference_eager_tf32"] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "loss": float(loss.item()),
                                           "flags": {"matmul_allow_tf32": True, "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32)}}
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        del model, opt
        torch.cuda.empty_cache()
    except Exception as e:   # noqa: BLE001
        out["reference_eager_fp32"] = {"failed": f"{type(e).__name__}: {str(e)[:300]}"}
    return out


def synthetic_inputs(lo, hi, T, L, pin=False, seed=SEED):
    """Synthetic prompts [lo, hi) of a GLOBAL batch: random ids in [2,72] with end token 1,
    random style vectors, injected x0 and per-step noise.  Sample g is drawn from its own generator seeded
    with (seed, g), so its inputs -- and therefore its result -- do not depend on how the batch is sharded."""
    n = hi - lo
    text = torch.empty(n, L, dtype=torch.int64)
    style = torch.empty(n, S_STYLE, 1280)
    x0 = torch.empty(n, T, 2)
    noise = torch.empty(NUM_STEPS, n, T, 2)
    g = torch.Generator()
    for j in range(n):
        g.manual_seed(seed * 1_000_003 + lo + j)
        text[j] = torch.randint(2, 73, (L,), generator=g)
        style[j] = torch.randn(S_STYLE, 1280, generator=g)
        x0[j] = torch.randn(T, 2, generator=g)
        noise[:, j] = torch.randn(NUM_STEPS, T, 2, generator=g)
    text[:, -1] = 1
    if pin:
        text, style, x0, noise = (t.pin_memory() for t in (text, style, x0, noise))
    return text, style, x0, noise


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)   # the poller must be gone before the host-timed e2e leg starts
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_traffic():
    """DRAM bytes per denoiser step from the committed ncu capture (profiles/): dram__bytes_read + dram__bytes_write
    summed over the launches of one step.  None when the summary is missing."""
    for name in ("r2f_step_dram_traffic.json", "r2_step_dram_traffic.json", "r1_step_dram_traffic_v12.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            d["file"] = "profiles/" + name
            return d
        except (OSError, ValueError):
            continue
    return None


# ---------------------------------------------------------------------------
# The reference itself (baseline/_ref: the unmodified package, see baseline/install_ref.py)
# ---------------------------------------------------------------------------
def reference_modules():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import install_ref
        return install_ref.import_reference()
    except Exception:   # noqa: BLE001 -- absent or not importable: the oracle port stands in
        return None


def reference_model(mods, device):
    """The reference's DiffusionModel with the bench's seeded weights, eval mode, on `device`."""
    from oracle.dhg_oracle import init_state_dict

    DiffusionModel = mods[0]
    model = DiffusionModel(num_layers=2, c1=128, c2=192, c3=256, drop_rate=0.0)
    model.load_state_dict(init_state_dict(0), strict=True)
    return model.to(device).eval()


@torch.no_grad()
def reference_chain(mods, model, text, style, x0, device, steps=None):
    """The loop of infer(), inference.py:81-96, re-stated around the reference's own DiffusionModel.forward and
    new_diffusion_step (that module itself imports `fire`, which is not installed).  The step draws its own noise."""
    _, new_diffusion_step, get_beta_set = mods
    beta_set = get_beta_set().to(device)
    alpha_set = torch.cumprod(1 - beta_set, dim=0)
    bs = text.shape[0]
    x = x0.to(device)
    pen_lifts = None
    order = list(range(len(beta_set) - 1, -1, -1))
    for i in order[:steps] if steps else order:
        alpha = alpha_set[i] * torch.ones((bs, 1, 1), device=device)
        beta = beta_set[i] * torch.ones((bs, 1, 1), device=device)
        a_next = alpha_set[i - 1] if i > 1 else torch.tensor(1.0, device=device)
        model_out, pen_lifts, _ = model(x, text, torch.sqrt(alpha), style)
        x = new_diffusion_step(x, model_out, beta, alpha, a_next)
    return torch.cat((x, pen_lifts.unsqueeze(2)), dim=2)


def cpu_reference_leg(batch, chains, warmup):
    """The reference's CPU path on all host cores, on a bounded sample of the same workload: `batch` prompts of
    T=392 / L=24 through full 60-step chains.  kind "reference" = the unmodified reference package (baseline/_ref),
    "port" = the oracle restatement when that directory is missing."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    text, style, x0, noise = synthetic_inputs(0, batch, T_STROKES, L_TEXT)
    mods = reference_modules()
    if mods is not None:
        model = reference_model(mods, "cpu")
        run = lambda steps=None: reference_chain(mods, model, text, style, x0, "cpu", steps)   # noqa: E731
        kind = "reference"
    else:
        from oracle import dhg_oracle as O

        sd = O.init_state_dict(0)
        run = lambda steps=None: O.reverse_chain(sd, text, style, x0, noise, steps=steps)   # noqa: E731
        kind = "port"
    for _ in range(warmup):
        run(2)
    t0 = time.perf_counter()
    for _ in range(chains):
        run()
    dt = time.perf_counter() - t0
    return batch * chains / dt, dt / chains, cores, kind


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    batch = args.ref_batch
    lines_s, s_per_chain, cores, kind = cpu_reference_leg(batch, args.steps, min(args.warmup, 1))
    what = "unmodified reference package (baseline/_ref)" if kind == "reference" else "oracle port"
    sample = f"{batch} prompts x full 60-step chain per step, T={T_STROKES} L={L_TEXT}, {what}, torch CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": "sampled handwriting lines/s (full 60-step reverse chain)", "value": lines_s,
        "unit": "lines/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": s_per_chain * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"best_exp sampling, T={T_STROKES}, L={L_TEXT}, 60 steps; CPU sample of {batch} prompts per step"},
        "cpu_baseline": {"value": lines_s, "unit": "lines/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": lines_s, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "us_per_denoiser_step": s_per_chain / NUM_STEPS * 1e6,
    }))


def gpu_eager_leg(dev, batch):
    """'the existing Blackwell kernel the build must beat is the reference run in PyTorch eager on
    the same B200'.  The unmodified reference model on `dev`, the loop of inference.py:84-94, fp32 (torch defaults:
    cuBLAS fp32 matmul, cuDNN convolutions with TF32 allowed) and under bf16 autocast; one warm-up of 3 steps, then
    one full 60-step chain each, CUDA-event timed."""
    mods = reference_modules()
    if mods is None:
        return {"unavailable": "baseline/_ref is missing (run baseline/install_ref.py in the build container)"}
    out = {"batch": batch, "T": T_STROKES, "L": L_TEXT, "api": "reference DiffusionModel.forward + new_diffusion_step, "
           "loop of inference.py:84-94, torch %s eager" % torch.__version__,
           "flags": {"matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32), "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32)}}
    text, style, x0, _ = synthetic_inputs(0, batch, T_STROKES, L_TEXT)
    text, style, x0 = text.to(dev), style.to(dev), x0.to(dev)
    model = reference_model(mods, dev)
    for name, ctx in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        try:
            def chain(steps=None):
                if ctx is None:
                    return reference_chain(mods, model, text, style, x0, dev, steps)
                with torch.autocast("cuda", dtype=ctx):
                    return reference_chain(mods, model, text, style, x0, dev, steps)
            chain(3)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = chain()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            out[name] = {"value": batch / (ms * 1e-3), "unit": "lines/s", "ms_per_chain": ms,
                         "us_per_denoiser_step": ms / NUM_STEPS * 1e3, "finite": bool(torch.isfinite(res).all().item())}
        except Exception as e:   # noqa: BLE001 -- e.g. SDPA rejecting the reference's fp32 mask next to bf16 q/k/v
            out[name] = {"failed": f"{type(e).__name__}: {str(e)[:200]}"}
    del model
    torch.cuda.empty_cache()
    return out


def parity_rows(out_rows, rows, T, L):
    """The CPU oracle on the given GLOBAL sample ids (fp32), against the rows the GPU produced for them."""
    from oracle import dhg_oracle as O

    sd = O.init_state_dict(0)
    ins = [synthetic_inputs(g, g + 1, T, L) for g in rows]
    text, style, x0, noise = (torch.cat([i[k] for i in ins], dim=1 if k == 3 else 0) for k in range(4))
    ref = O.reverse_chain(sd, text, style, x0, noise)
    rel = ((out_rows[..., :2] - ref[..., :2]).norm() / ref[..., :2].norm()).item()
    sure = (ref[..., 2] - 0.5).abs() > 1e-2
    agree = ((out_rows[..., 2] > 0.5) == (ref[..., 2] > 0.5))[sure].float().mean().item()
    return {"rows": list(rows), "strokes_rel_l2": rel, "pen_agreement": agree, "pen_margin": 0.01,
            "pen_positions_counted": int(sure.sum().item()), "pen_abs_diff_mean": (out_rows[..., 2] - ref[..., 2]).abs().mean().item(),
            "oracle": "oracle/dhg_oracle.py fp32 CPU, same weights and injected noise"}


def time_chain(w, text, style, x0, noise, steps, warmup, dev):
    for _ in range(warmup):
        out = w.sample(text, style, x0=x0, noise=noise)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = w.sample(text, style, x0=x0, noise=noise)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps, out


def small_config(sd, dev, dtype, B, T, L, steps=2):
    """One of the other BASELINE configurations on this GPU: device-resident chain time."""
    from dhg_b200 import DiffusionWriter

    w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype=dtype, device=dev, chunk=B)
    text, style, x0, noise = (t.to(dev) for t in synthetic_inputs(0, B, T, L))
    ms, out = time_chain(w, text, style, x0, noise, steps, 2, dev)
    res = {"B": B, "T": T, "L": L, "dtype": dtype, "ms_per_chain": ms, "us_per_denoiser_step": ms / NUM_STEPS * 1e3,
           "lines_per_s": B / (ms * 1e-3), "finite": bool(torch.isfinite(out).all().item()),
           "tflops_alg": B * flops_alg_per_sample_step(T, L) * NUM_STEPS / (ms * 1e-3) / 1e12}
    w.close()
    del w, text, style, x0, noise, out
    torch.cuda.empty_cache()
    return res


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist

    from dhg_b200 import DiffusionWriter
    from dhg_b200.sharding import sample_sharded, shard_bounds
    from oracle.dhg_oracle import init_state_dict  # seeded random-init weights (no checkpoint offline)

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    T, L = T_STROKES, L_TEXT
    strong = args.global_batch > 0
    G = args.global_batch if strong else args.batch * world
    lo, hi = shard_bounds(G, rank, world)
    B = hi - lo
    sd = init_state_dict(0)
    w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype=args.dtype, device=dev, chunk=args.chunk)
    # this rank's contiguous slice of the GLOBAL batch (sharding.shard_bounds); sample g's inputs depend on g only
    text_h, style_h, x0_h, noise_h = synthetic_inputs(lo, hi, T, L, pin=True)
    text, style, x0, noise = (t.to(dev) for t in (text_h, style_h, x0_h, noise_h))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def maxreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        out = w.sample(text, style, x0=x0, noise=noise)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = w.sample(text, style, x0=x0, noise=noise)
    e1.record()
    barrier()
    ms_total = maxreduce(e0.elapsed_time(e1))
    clk = clocks.stop()
    launches = w.last_launch_count * args.steps
    finite = bool(torch.isfinite(out).all().item())
    ms_step = ms_total / args.steps
    value = G * args.steps / (ms_total * 1e-3)

    # ---- parity at the benchmarked configuration: rows of the TIMED result against the CPU oracle ----
    parity = None
    if rank == 0 and args.parity_rows > 0:
        n = min(args.parity_rows, B)
        local = sorted({int(round(i * (B - 1) / max(n - 1, 1))) for i in range(n)})
        parity = parity_rows(out[local].float().cpu(), [lo + r for r in local], T, L)
        parity["dtype"] = args.dtype
        parity["batch"] = B

    # ---- end to end through the host-buffer C-ABI call ("e2e") ----
    w.sample_host(text_h, style_h, x0_h, noise_h)
    barrier()
    t0 = time.perf_counter()
    e2e_calls = []
    for _ in range(args.e2e_steps):
        tc = time.perf_counter()
        out_h = w.sample_host(text_h, style_h, x0_h, noise_h)   # returns after the device->host copy has completed
        e2e_calls.append(round(1e3 * (time.perf_counter() - tc), 2))
    torch.cuda.synchronize(dev)
    e2e_s = maxreduce(time.perf_counter() - t0)
    barrier()
    e2e_value = G * args.e2e_steps / e2e_s
    h2d = sum(t.numel() * t.element_size() for t in (text_h, style_h, x0_h, noise_h))
    d2h = out_h.numel() * out_h.element_size()
    e2e_same = bool(torch.equal(out_h, out.cpu()))   # host-buffer call and device call: same bits

    # ---- HBM roofline of the fused posterior-update kernel: ONE launch over 64 chunks' worth of points ----
    n_big = 64 * min(B, 1024) * T * 2          # floats per array; 4 arrays -> >= 0.8 GB of traffic per launch at B >= 1024
    xb = torch.randn(n_big, device=dev)
    eb = torch.randn(n_big, device=dev)
    zb = torch.randn(n_big, device=dev)
    ob = torch.empty(n_big, device=dev)
    for _ in range(2):
        w.posterior_step(30, xb, eb, zb, out=ob)
    torch.cuda.synchronize(dev)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    p0.record()
    for _ in range(reps):
        w.posterior_step(30, xb, eb, zb, out=ob)
    p1.record()
    torch.cuda.synchronize(dev)
    post_us = p0.elapsed_time(p1) / reps * 1e3
    post_bytes = 16.0 * n_big                  # read x, eps, z + write x (fp32) = 32*T bytes per sample per step
    post_gbs = post_bytes / (post_us * 1e-6) / 1e9
    del xb, eb, zb, ob

    # ---- N-independence: a 64-prompt GLOBAL batch through sharding.sample_sharded over the ranks (NCCL gather) ----
    equivalence = None
    if args.equivalence:
        GE = 64
        ins = synthetic_inputs(0, GE, T, L, seed=SEED + 1)
        we = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype=args.dtype, device=dev, chunk=GE)
        fn = lambda t, s, x0, noise: we.sample(t.to(dev), s.to(dev), x0=x0.to(dev), noise=noise.to(dev))   # noqa: E731
        full = sample_sharded(fn, *ins, rank, world, gather=True)
        torch.cuda.synchronize(dev)
        if rank == 0:
            equivalence = {"global_batch": GE, "ranks": world, "per_rank": [shard_bounds(GE, r, world)[1] - shard_bounds(GE, r, world)[0] for r in range(world)],
                           "sha1": hashlib.sha1(full.float().cpu().numpy().tobytes()).hexdigest(),
                           "what": "sha1 of the gathered [64,T,3] fp32 result of a fixed 64-prompt global batch sampled through "
                                   "dhg_b200.sharding.sample_sharded on N ranks; the same string for every N"}
        we.close()
        del we, full
        torch.cuda.empty_cache()
    barrier()

    # ---- the dominant kernel on its own: every tcgen05 GEMM launch of one denoiser step, timed per family ----
    w_bytes = w.plan_bytes
    w.close()          # free the plan before the single-GPU extra legs
    del text, style, x0, noise, out
    torch.cuda.empty_cache()
    if rank != 0:
        return
    gemm = None
    if args.gemm_roofline and args.dtype == "bf16":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from gemm_cases import time_step_gemms

        gemm = time_step_gemms(1024, repeats=5)
        torch.cuda.empty_cache()

    peaks = measured_peaks()
    f_alg = flops_alg_per_sample_step(T, L)
    tflops = G / world * f_alg * NUM_STEPS / (ms_step * 1e-3) / 1e12     # per GPU
    peak_tf = peaks["bf16_tflops_sustained"]
    traffic = measured_traffic()
    line = {
        "metric": "sampled handwriting lines/s (full 60-step reverse chain)", "value": value, "unit": "lines/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {
            "workload": (f"best_exp sampling, BASELINE configs[2] as written: fixed global batch {G} sharded over {world} GPU(s)" if strong else
                         f"best_exp sampling, {args.batch} prompts per GPU (the per-GPU shard of BASELINE configs[2]'s 8192 at 8 GPUs), weak scaling") +
                        f", T={T}, L={L}, 60 steps, seeded random-init weights, injected noise",
            "per_gpu_batch": B, "global_batch": G, "T": T, "L": L, "chunk": min(B, args.chunk),
            "parallelism": f"batch-sharded x{world} (sharding.shard_bounds, globally indexed prompts), no collective in the loop",
            "l2": f"working set {w_bytes / 1e9:.1f} GB per chain >> 126 MB L2 (inputs larger than L2)",
        },
        "us_per_denoiser_step": ms_step / NUM_STEPS * 1e3 / max(1, -(-B // args.chunk)),
        "finite": finite,
        "gpu_launches": int(launches),
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "lines/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": args.e2e_steps, "ms_per_call": e2e_calls, "same_bits_as_device_call": e2e_same,
                "api": "DiffusionWriter.sample_host -> dhg_sample_host (pinned host buffers)"},
        "roofline": {
            "bound": "tensor", "achieved": tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": tflops / peak_tf,
            "traffic": (traffic or {}).get("step_dram_bytes") if B == 1024 else None,
            "traffic_unit": "DRAM bytes per denoiser step, all launches (ncu, %s)" % (traffic or {}).get("file"),
            "peak_source": peaks["source"] + " (bf16_tflops_sustained: chain timed inside a long step)",
            "frac_of_burst_peak": tflops / peaks["bf16_tflops"],
            "definition": "B * F_alg(T,L) * 60 / t_chain per GPU (SURVEY.md 8d); F_alg = %.1f MFLOP/sample/step" % (f_alg / 1e6),
        },
        "roofline_posterior_update": {
            "kernel": "posterior_kernel (dhg_posterior_step), one launch over %d floats" % n_big,
            "bound": "hbm", "achieved": post_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": post_gbs / peaks["hbm_gbs"],
            "us_per_launch": post_us, "bytes_per_launch": post_bytes, "peak_source": peaks["source"],
        },
    }
    if parity:
        line["parity"] = parity
    if equivalence:
        line["equivalence"] = equivalence
    if gemm:
        us = gemm["us"]
        line["roofline_gemm"] = {
            "kernel": "tc_gemm_kernel (all %d launches of one denoiser step, each family timed alone, 5 launches per family)" % gemm["launches"],
            "bound": "hbm", "achieved": gemm["bytes"] / us / 1e3, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": gemm["bytes"] / us / 1e3 / peaks["hbm_gbs"],
            "traffic": (traffic or {}).get("gemm_dram_bytes_per_step"),
            "traffic_unit": "DRAM bytes of the same launches (ncu, cold L2 per kernel; writes still in L2 at kernel end not counted)",
            "tflops": gemm["flop"] / us / 1e6, "tensor_frac": gemm["flop"] / us / 1e6 / peaks["bf16_tflops"],
            "us_per_step": us, "share_of_step": us / (ms_step / NUM_STEPS * 1e3),
            "algorithmic_bytes_per_step": gemm["bytes"], "flop_per_step": gemm["flop"],
            "peak_source": peaks["source"] + " (burst figures: families timed in isolation)",
        }
    if world == 1 and args.extras:
        # a like-precision figure: the same workload in the fp32-contract mode
        if args.dtype == "bf16":
            try:
                r = small_config(sd, dev, "fp32", args.fp32_batch, T, L, steps=1)
                r["path"] = "fp32 contract (parity 1e-3): see DESIGN.md for the kernels this mode runs"
                r["frac_of_sustained_bf16_peak"] = r["tflops_alg"] / peak_tf
                line["fp32"] = r
            except Exception as e:   # noqa: BLE001
                line["fp32"] = {"failed": str(e)[:200]}
        other = {}
        for key, (dt, b, t, l) in {"configs[1] B=64 bf16": ("bf16", 64, T, L), "configs[1] B=64 fp32": ("fp32", 64, T, L),
                                   "B=1 bf16 (latency)": ("bf16", 1, T, L), "B=1 fp32 (latency)": ("fp32", 1, T, L),
                                   "configs[4] shard: B=128 T=1200 L=81 bf16": ("bf16", 128, 1200, 81)}.items():
            try:
                other[key] = small_config(sd, dev, dt, b, t, l)
            except Exception as e:   # noqa: BLE001
                other[key] = {"failed": str(e)[:200]}
        line["other_configs"] = other
        line["gpu_eager_baseline"] = gpu_eager_leg(dev, args.eager_batch)
        # BASELINE configs[1] / configs[0] shapes: the reference's PyTorch path on this GPU beside `other_configs`
        line["gpu_eager_baseline_small"] = {f"B={b}": {k: v for k, v in gpu_eager_leg(dev, b).items() if k in ("fp32", "bf16_autocast")}
                                            for b in (64, 1)}
        line["train_update"] = train_update_leg(dev, peaks)
        line["train_step"] = train_step_leg(dev)
    if args.cpu_baseline and world >= 1:
        lines_s, s_chain, cores, kind = cpu_reference_leg(args.cpu_batch, 1, 1)
        line["cpu_baseline"] = {
            "value": lines_s, "unit": "lines/s", "cores": cores, "kind": kind,
            "sample": f"{args.cpu_batch} prompts x one full 60-step chain, T={T} L={L}, "
                      + ("unmodified reference package (baseline/_ref)" if kind == "reference" else "oracle port")
                      + f", torch CPU fp32, {s_chain:.1f} s",
        }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="prompts per GPU (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0, help="fixed global batch sharded over the ranks (strong scaling; BASELINE configs[2]: 8192)")
    ap.add_argument("--chunk", type=int, default=1024, help="prompts per captured chain")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-batch", type=int, default=32, help="--impl reference: prompts per step (one full chain each)")
    ap.add_argument("--cpu-batch", type=int, default=128, help="prompts of the cpu_baseline chain (about 10 s on 16 cores)")
    ap.add_argument("--parity-rows", type=int, default=4, help="rows of the timed result checked against the CPU oracle")
    ap.add_argument("--fp32-batch", type=int, default=1024, help="batch of the fp32-contract leg")
    ap.add_argument("--eager-batch", type=int, default=1024, help="batch of the GPU-eager reference baseline")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-gemm-roofline", dest="gemm_roofline", action="store_false")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip fp32 / other configs / GPU-eager baseline legs")
    ap.add_argument("--no-equivalence", dest="equivalence", action="store_false")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: >= 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
