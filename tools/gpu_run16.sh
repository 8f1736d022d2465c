set -x
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --parity-queries 0"
RB200_XATTN_B=10 timeout 300 $B 2>&1 | tail -1 | tee gpurun_out/bench16_xb10.json | cut -c1-200
timeout 400 $B --steps 2 --batch 128 --beams 100 2>&1 | tail -1 | tee gpurun_out/bench16_c3.json | cut -c1-200
timeout 400 $B --steps 2 --model t5-large --batch 512 2>&1 | tail -1 | tee gpurun_out/bench16_c4.json | cut -c1-200
timeout 400 $B --steps 2 --docid-len 16 --codebook 1024 2>&1 | tail -1 | tee gpurun_out/bench16_c5.json | cut -c1-200
timeout 300 python bench.py --steps 3 --warmup 3 --precision tf32x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench16_tf32x3.json | cut -c1-200
