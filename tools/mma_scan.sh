for N in 64 128 256; do for K in 64 128 256 512; do python tools/gemm_one.py 402433 $K $N 1; done; done
for K in 64 128; do python tools/gemm_one.py 402433 $K 128 3; done
