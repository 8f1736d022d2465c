set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "golden or t5base_search or long_docid or overflow" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu2.log
timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench2_fp16x3.json | cut -c1-400
RB200_PDL=1 timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline --parity-queries 0 2>&1 | tail -1 | tee gpurun_out/bench2_fp16x3_pdl.json | cut -c1-400
timeout 200 python tools/gemm_bench.py --precision fp16x3 2>&1 | tee gpurun_out/gemm_bench_fp16x3.txt
RB200_PDL=1 timeout 200 python tools/gemm_bench.py --precision fp16x3 2>&1 | tee gpurun_out/gemm_bench_fp16x3_pdl.txt
timeout 200 python tools/gemm_bench.py --precision fp16x3 --shapes "o64:768:64:1,o128:768:128:1,o256:768:256:1,o1536:768:1536:1" 2>&1 | tee gpurun_out/gemm_bench_k.txt
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches2_fp16x3.csv python tools/profile_step.py --precision fp16x3 > gpurun_out/prof2_fp16x3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches2_fp16x3.csv | tee gpurun_out/launch_summary2_fp16x3.txt | head -12
