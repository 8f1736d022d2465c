# Round-end validation on one B200: GPU parity suite, smoke, the default bench line, the launch list of one search.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 | tee gpurun_out/smoke_final.log
timeout 400 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_final.json | cut -c1-300
P="python tools/profile_step.py --precision fp16x3"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv $P > gpurun_out/prof_final.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_final.csv | tee gpurun_out/launch_summary_final.txt | head -12
