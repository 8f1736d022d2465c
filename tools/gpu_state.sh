set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 --precision tf32x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tf32x3.json | cut -c1-2500
timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp16x3.json | cut -c1-2500
RB200_GEMM=1cta timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline --parity-queries 0 2>&1 | tail -1 | tee gpurun_out/bench_fp16x3_1cta.json | cut -c1-600
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp16x3.csv python tools/profile_step.py --precision fp16x3 > gpurun_out/prof_fp16x3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_fp16x3.csv | tee gpurun_out/launch_summary_fp16x3.txt | head -14
