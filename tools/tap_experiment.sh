mkdir -p gpurun_out
for o in "tap_base_offset=0" "tap_base_offset=1" "tap_shift=0"; do
  DHG_OPTS=$o timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu 2>&1 | tail -4 > gpurun_out/t_gemm_$o.log; echo "== $o"; tail -3 gpurun_out/t_gemm_$o.log
done
