#!/bin/bash
# One GPU visit: GEMM unit test, bring-up check, full gpu test suite.  Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"
tail -5 gpurun_out/t_gemm.log
timeout 300 python tools/gpu_check.py bf16 1 > gpurun_out/check_bf16_tc.log 2>&1; echo "check rc=$?"
tail -25 gpurun_out/check_bf16_tc.log
