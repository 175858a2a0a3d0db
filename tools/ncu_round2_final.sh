#!/bin/bash
# Profiling pass of the final round-2 code (run under gpurun): launch list of the bf16 chain and `ncu --set full` of the
# level-1 self-attention and of the q|k|v projection.  Same recipe as tools/ncu_round2.sh (plain run first, reports
# reduced to text on the box).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
R=201729
prof() {   # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/r2_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  local rc=$?
  if [ -f /tmp/r2_$name.ncu-rep ]; then
    python tools/ncu_top.py /tmp/r2_$name.ncu-rep 40 > gpurun_out/r2f_ncu_full_$name.txt 2>&1
    rm -f /tmp/r2_$name.ncu-rep
  fi
  echo "$name rc=$rc"
}
DHG_OPTS=autotune=0 python tools/profile_chain.py 1024 bf16 > gpurun_out/plain_chain.log 2>&1 && \
DHG_OPTS=autotune=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 330 --csv \
    --log-file gpurun_out/r2f_step_launches.csv python tools/profile_chain.py 1024 bf16 > gpurun_out/ncu_chain.log 2>&1
echo "launch list rc=$?"
DHG_OPTS=autotune=0 prof attn_self_l1 attn_tc_kernel 3 python tools/profile_chain.py 1024 bf16
prof gemm_qkv tc_gemm 2 python tools/gemm_one.py $R 192 576 1 period=197 rowbias
