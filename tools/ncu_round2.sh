#!/bin/bash
# Round-2 profiling pass (run under gpurun): plain run first, then ncu on the same command (B200_PROFILING.md).
# The .ncu-rep files are ~20 MB each and gpurun_out/ is capped at 64 MiB, so each capture is reduced to text on the box
# (headline metrics + the most-stalled source lines, tools/ncu_top.py; the raw metric page as csv) and the report deleted.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
R=201729   # level-1 rows at B = 1024 (1024 * 197 + 1)
prof() {   # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/r2_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  local rc=$?
  if [ -f /tmp/r2_$name.ncu-rep ]; then
    python tools/ncu_top.py /tmp/r2_$name.ncu-rep 40 > gpurun_out/r2_ncu_full_$name.txt 2>&1
    ncu -i /tmp/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_${name}_raw.csv 2>/dev/null
    rm -f /tmp/r2_$name.ncu-rep
  fi
  echo "$name rc=$rc"
}
# launch list + DRAM bytes of the chain (no graph, built-in tile rule so that plan time stays short under ncu)
DHG_OPTS=autotune=0 python tools/profile_chain.py 1024 bf16 > gpurun_out/plain_chain.log 2>&1 && \
DHG_OPTS=autotune=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 330 --csv \
    --log-file gpurun_out/r2_step_launches.csv python tools/profile_chain.py 1024 bf16 > gpurun_out/ncu_chain.log 2>&1
echo "launch list rc=$?"
DHG_OPTS=autotune=0 python tools/profile_chain.py 1024 fp32 > gpurun_out/plain_chain_fp32.log 2>&1 && \
DHG_OPTS=autotune=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv \
    --log-file gpurun_out/r2_step_launches_fp32.csv python tools/profile_chain.py 1024 fp32 > gpurun_out/ncu_chain_fp32.log 2>&1
echo "fp32 launch list rc=$?"
prof gemm_qkv tc_gemm 2 python tools/gemm_one.py $R 192 576 1 period=197 rowbias
prof gemm_dense2 tc_gemm 2 python tools/gemm_one.py $R 192 192 1 period=197 ln film res_pre both
prof gemm_ffn1 tc_gemm 2 python tools/gemm_one.py $R 192 384 1 period=197 act
prof gemm_dual_dec2 tc_gemm 2 python tools/gemm_one.py $R 192 192 1 period=197 dual=256
prof gemm_conv2_l0 tc_gemm 2 python tools/gemm_one.py 402433 64 128 3 period=393 film act
DHG_OPTS=autotune=0 prof heads heads_from_dots 1 python tools/profile_chain.py 1024 bf16
DHG_OPTS=autotune=0 prof attn_self_l1 attn_tc_kernel 3 python tools/profile_chain.py 1024 bf16
