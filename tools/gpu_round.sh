#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -15 gpurun_out/t_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench.log
timeout 120 python tools/profile_step.py 1024 bf16 3 > gpurun_out/step_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 90 --csv --log-file gpurun_out/step_launches.csv python tools/profile_step.py 1024 bf16 3 > gpurun_out/step_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/step_plain.log
