"""Time every GEMM family of one denoiser step at batch B through the C-ABI test hook.
Usage: python tools/gemm_bench.py [B]   (prints us per launch, TFLOP/s, algorithmic GB/s)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_cases import time_step_gemms  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
print(f"B={B}  {'case':38s} {'us':>8s} {'TF/s':>7s} {'GB/s':>7s}  x")
r = time_step_gemms(B, verbose=True)
print(f"total per step: {r['us']:.0f} us, {r['flop']/r['us']/1e6:.1f} TFLOP/s, {r['bytes']/r['us']/1e3:.0f} GB/s algorithmic "
      f"({r['flop']/1e9:.1f} GFLOP, {r['bytes']/1e9:.2f} GB)")
