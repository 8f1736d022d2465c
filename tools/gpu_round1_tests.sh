set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -q --timeout 200 -k "golden or t5base_search or cli" 2>&1 | tail -4
timeout 200 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2200
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp16x3.csv python tools/profile_step.py --precision fp16x3 --steps-only 8 > gpurun_out/prof4.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_fp16x3.csv > gpurun_out/launch_summary.txt; head -9 gpurun_out/launch_summary.txt
timeout 200 python bench.py --steps 3 --warmup 2 --precision fp16x3 --no-cpu-baseline --parity-queries 0 --batch 512 2>&1 | tail -1 | cut -c1-300
