set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -8 | tee gpurun_out/pytest_gpu10.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $B 2>&1 | tail -1 | tee gpurun_out/bench10.json | cut -c1-200
RB200_FOLD=0 timeout 300 $B --parity-queries 0 2>&1 | tail -1 | tee gpurun_out/bench10_nofold.json | cut -c1-200
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches10.csv python tools/profile_step.py --precision fp16x3 > gpurun_out/prof10.log 2>&1
python tools/summarize_launches.py gpurun_out/launches10.csv | tee gpurun_out/launch_summary10.txt | head -10
