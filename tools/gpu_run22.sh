set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 400 -k "forced_tail or t5base_search" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu22.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench22.json | cut -c1-250
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches22.csv python tools/profile_step.py --precision fp16x3 > gpurun_out/prof22.log 2>&1
python tools/summarize_launches.py gpurun_out/launches22.csv | tee gpurun_out/launch_summary22.txt | head -7
