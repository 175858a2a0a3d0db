#!/bin/bash
# One GPU round: unit tests + family timings, full parity tests, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -x -q -m gpu > gpurun_out/t_unit.log 2>&1; echo "unit tests rc=$?"
tail -8 gpurun_out/t_unit.log
timeout 300 python tools/attn_bench.py 1024 > gpurun_out/attn_bench.log 2>&1; echo "attn bench rc=$?"
tail -10 gpurun_out/attn_bench.log
if [ -z "$SKIP_GEMM_BENCH" ]; then
timeout 300 python tools/gemm_bench.py 1024 > gpurun_out/gemm_bench.log 2>&1; echo "gemm bench rc=$?"
tail -60 gpurun_out/gemm_bench.log
fi
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_gemm.py --deselect tests/test_gpu_attention.py > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -8 gpurun_out/t_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench.log
