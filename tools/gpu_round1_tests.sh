set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -x -k "gemm_modes" 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -k "golden or t5base_search or overflow" 2>&1 | tail -8
RB200_GEMM=1cta timeout 600 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline --parity-queries 0 2>&1 | tail -1 | cut -c1-1800
timeout 600 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2200
timeout 600 python bench.py --steps 3 --warmup 3 --precision tf32x3 --no-cpu-baseline --parity-queries 0 2>&1 | tail -1 | cut -c1-1800
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp16x3.csv python tools/profile_step.py --precision fp16x3 --steps-only 8 > gpurun_out/prof4.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_fp16x3.csv
