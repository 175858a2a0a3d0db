#!/bin/bash
# One GPU round: GEMM unit tests + GEMM family timings, full parity tests, bench, launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > gpurun_out/t_gemm.log 2>&1; echo "gemm tests rc=$?"
tail -8 gpurun_out/t_gemm.log
timeout 300 python tools/gemm_bench.py 1024 > gpurun_out/gemm_bench.log 2>&1; echo "gemm bench rc=$?"
tail -60 gpurun_out/gemm_bench.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_gemm.py > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -8 gpurun_out/t_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench.log
