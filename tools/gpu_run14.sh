set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -8 | tee gpurun_out/pytest_gpu14.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $B 2>&1 | tail -1 | tee gpurun_out/bench14.json | cut -c1-250
RB200_TAIL=0 timeout 300 $B --parity-queries 0 2>&1 | tail -1 | tee gpurun_out/bench14_notail.json | cut -c1-250
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches14.csv python tools/profile_step.py --precision fp16x3 > gpurun_out/prof14.log 2>&1
python tools/summarize_launches.py gpurun_out/launches14.csv | tee gpurun_out/launch_summary14.txt | head -16
