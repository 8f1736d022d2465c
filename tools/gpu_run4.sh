set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "gemm_modes or golden or t5base_search or long_docid or overflow" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu4.log
timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench4_fp16x3.json | cut -c1-300
RB200_PDL=1 timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench4_fp16x3_pdl.json | cut -c1-300
RB200_GEMM_TRACE=1 timeout 200 python tools/gemm_bench.py --precision fp16x3 --shapes "o64:768:64:1,o:768:768:1,wi:3072:768:2" 2>&1 | tee gpurun_out/gemm_trace4.txt
RB200_PDL=1 timeout 200 python tools/gemm_bench.py --precision fp16x3 2>&1 | tee gpurun_out/gemm_bench4_pdl.txt
RB200_PDL=1 timeout 200 python tools/gemm_bench.py --precision tf32x3 2>&1 | tee gpurun_out/gemm_bench4_tf32x3_pdl.txt
