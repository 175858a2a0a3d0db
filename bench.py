"""Benchmark of the reverse-diffusion sampling hot path (BASELINE.json metric:
sampled handwriting lines/s for the full 60-step chain; us per denoiser step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch: the full 60-step reverse
chain (inference.py:81-96) for the per-GPU batch of synthetic prompts.
Workload at N=1: the per-GPU shard of BASELINE configs[2] ("best_exp config
sampling, batch 8192 synthetic prompts, batch-sharded at 1/2/4/8 B200"):
8192 / 8 = 1024 prompts per GPU, T=392 stroke points, L=24 tokens -- the shape
of configs[0]/[1] ('Follow the White Rabbit').  Weak scaling: per-GPU batch is
fixed, no collective in the loop.

Under torchrun (N>1) every rank drives its own GPU; the timed region is
bracketed by barrier + synchronize, timed with CUDA events, max over ranks.
"""
import argparse
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


def synthetic_inputs(B, T, L, seed, pin=False):
    """Synthetic prompts of the BASELINE shape: random ids in [2,72] with end token 1, random style
    vectors, injected x0 and per-step noise."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(2, 73, (B, L), generator=g)
    text[:, -1] = 1
    style = torch.randn(B, S_STYLE, 1280, generator=g)
    x0 = torch.randn(B, T, 2, generator=g)
    noise = torch.randn(NUM_STEPS, B, T, 2, generator=g)
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
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_step_dram_traffic_v12.json")
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def cpu_oracle_leg(batch, chains, warmup):
    """The reference's CPU path (oracle port, torch CPU fp32, all host cores) on a bounded sample of
    the same workload: `batch` prompts of T=392/L=24 through full 60-step chains."""
    from oracle import dhg_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.init_state_dict(0)
    text, style, x0, noise = synthetic_inputs(batch, T_STROKES, L_TEXT, 1234)
    for _ in range(warmup):
        O.reverse_chain(sd, text, style, x0, noise, steps=2)
    t0 = time.perf_counter()
    for _ in range(chains):
        O.reverse_chain(sd, text, style, x0, noise)
    dt = time.perf_counter() - t0
    return batch * chains / dt, dt / chains, cores


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    batch = args.ref_batch
    lines_s, s_per_chain, cores = cpu_oracle_leg(batch, args.steps, min(args.warmup, 1))
    sample = f"{batch} prompts x full 60-step chain per step, T={T_STROKES} L={L_TEXT}, torch CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": "sampled handwriting lines/s (full 60-step reverse chain)", "value": lines_s,
        "unit": "lines/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": s_per_chain * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"best_exp sampling, T={T_STROKES}, L={L_TEXT}, 60 steps; CPU sample of {batch} prompts per step"},
        "cpu_baseline": {"value": lines_s, "unit": "lines/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": lines_s, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "us_per_denoiser_step": s_per_chain / NUM_STEPS * 1e6,
    }))


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist

    from dhg_b200 import DiffusionWriter
    from oracle.dhg_oracle import init_state_dict  # seeded random-init weights (no checkpoint offline)

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, T, L = args.batch, T_STROKES, L_TEXT
    sd = init_state_dict(0)
    w = DiffusionWriter(state_dict=sd, num_layers=2, channels=128, dtype=args.dtype, device=dev, chunk=args.chunk)
    # distinct prompts per rank: global sample index = rank * B + b
    text_h, style_h, x0_h, noise_h = synthetic_inputs(B, T, L, 1234 + rank, pin=True)
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
    value = world * B * args.steps / (ms_total * 1e-3)

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
    e2e_value = world * B * args.e2e_steps / e2e_s
    h2d = sum(t.numel() * t.element_size() for t in (text_h, style_h, x0_h, noise_h))
    d2h = out_h.numel() * out_h.element_size()

    # ---- HBM roofline of the standalone fused posterior-update kernel ----
    n = x0.numel()
    eps = torch.randn_like(x0)
    big = torch.randn(64, *x0.shape, device=dev)  # rotate through 64 x-buffers (> L2 together with eps/z/out)
    outb = torch.empty_like(big)
    for i in range(8):
        w.posterior_step(30, big[i], eps, noise[i % NUM_STEPS], out=outb[i])
    torch.cuda.synchronize(dev)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(64):
        w.posterior_step(30, big[i], eps, noise[i % NUM_STEPS], out=outb[i])
    p1.record()
    torch.cuda.synchronize(dev)
    post_us = p0.elapsed_time(p1) / 64 * 1e3
    post_gbs = 16.0 * n / (post_us * 1e-6) / 1e9   # 32*T bytes per sample: read x, eps, z + write x (fp32)

    # ---- the dominant kernel on its own: every tcgen05 GEMM launch of one denoiser step, timed per family ----
    gemm = None
    if rank == 0 and args.gemm_roofline and B == 1024:
        w_bytes = w.plan_bytes
        w.close()          # free the 6 GB plan before allocating the GEMM operands
        del text, style, x0, noise, big, outb, eps, out
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from gemm_cases import time_step_gemms

        gemm = time_step_gemms(B, repeats=5)
    else:
        w_bytes = w.plan_bytes

    if rank != 0:
        return
    peaks = measured_peaks()
    f_alg = flops_alg_per_sample_step(T, L)
    tflops = B * f_alg * NUM_STEPS / (ms_step * 1e-3) / 1e12
    peak_tf = peaks["bf16_tflops_sustained"]
    traffic = measured_traffic()
    line = {
        "metric": "sampled handwriting lines/s (full 60-step reverse chain)", "value": value, "unit": "lines/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {
            "workload": f"best_exp sampling (BASELINE configs[2] shard): {B} prompts/GPU, T={T}, L={L}, 60 steps, "
                        "seeded random-init weights, injected noise",
            "per_gpu_batch": B, "global_batch": B * world, "T": T, "L": L, "chunk": min(B, args.chunk),
            "parallelism": f"batch-sharded x{world}, no collective in the loop",
            "l2": f"working set {w_bytes / 1e9:.1f} GB per chain >> 126 MB L2 (inputs larger than L2)",
        },
        "us_per_denoiser_step": ms_step / NUM_STEPS * 1e3,
        "finite": finite,
        "gpu_launches": int(launches),
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "lines/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": args.e2e_steps, "ms_per_call": e2e_calls, "api": "DiffusionWriter.sample_host -> dhg_sample_host (pinned host buffers)"},
        "roofline": {
            "bound": "tensor", "achieved": tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": tflops / peak_tf,
            "traffic": (traffic or {}).get("step_dram_bytes") if B == 1024 else None,
            "traffic_unit": "DRAM bytes per denoiser step, all launches (ncu, profiles/r1_step_dram_traffic_v12.json)",
            "peak_source": peaks["source"] + " (bf16_tflops_sustained: chain timed inside a long step)",
            "definition": "B * F_alg(T,L) * 60 / t_chain (SURVEY.md 8d); F_alg = %.1f MFLOP/sample/step" % (f_alg / 1e6),
        },
        "roofline_posterior_update": {
            "bound": "hbm", "achieved": post_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": post_gbs / peaks["hbm_gbs"],
            "us_per_launch": post_us, "bytes_per_launch": 16 * n, "peak_source": peaks["source"],
        },
    }
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
    if args.cpu_baseline and world >= 1:
        lines_s, s_chain, cores = cpu_oracle_leg(args.cpu_batch, 1, 1)
        line["cpu_baseline"] = {
            "value": lines_s, "unit": "lines/s", "cores": cores, "kind": "port",
            "sample": f"{args.cpu_batch} prompts x one full 60-step chain, T={T} L={L}, oracle port (torch CPU fp32), {s_chain:.1f} s",
        }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="prompts per GPU")
    ap.add_argument("--chunk", type=int, default=1024, help="prompts per captured chain")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-batch", type=int, default=32, help="--impl reference: prompts per step (one full chain each)")
    ap.add_argument("--cpu-batch", type=int, default=128, help="prompts of the cpu_baseline chain (about 10 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-gemm-roofline", dest="gemm_roofline", action="store_false")
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
