set -x
mkdir -p gpurun_out
SH="qkv:2304:768:0,o:768:768:1,wi:3072:768:2,wo:768:3072:1"
RB200_BN=128 timeout 200 python tools/gemm_bench.py --precision fp16x3 --M 71680 --iters 30 --rotate-mb 3000 --shapes $SH 2>&1 | tee gpurun_out/gemm_tail_bn128.txt
RB200_BN=256 timeout 200 python tools/gemm_bench.py --precision fp16x3 --M 71680 --iters 30 --rotate-mb 3000 --shapes $SH 2>&1 | tee gpurun_out/gemm_tail_bn256.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench21.json | cut -c1-250
