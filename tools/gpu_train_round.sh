#!/bin/bash
# One GPU round for the training step: its tests, BASELINE configs[3] timing beside the reference's eager step, launch list of one step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py -x -q -m gpu -s > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?"
tail -12 gpurun_out/t_train.log
timeout 300 python tools/train_step_bench.py > gpurun_out/train_bench.log 2>&1; echo "train bench rc=$?"
tail -3 gpurun_out/train_bench.log
if [ -z "$SKIP_NCU" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_step_bench.py --one > gpurun_out/train_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/train_ncu.log
fi
