for N in 64 128 256; do
for o in "interleave=0,mma_repeat=1" "interleave=0,mma_repeat=8" "interleave=1,mma_repeat=8"; do
  echo "== N=$N $o"
  DHG_OPTS=$o DHG_DESCRIBE=1 DHG_TRACE=1 python tools/gemm_one.py 402433 128 $N 3 2> gpurun_out/tr.txt | tail -1
  python tools/trace_summary.py gpurun_out/tr.txt 6 6 | grep -E "plan|M_a_ok|M_commit|tile period|MMA warp"
done; done
