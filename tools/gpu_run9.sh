set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 400 -k "golden or t5base_search or long_docid or wide_beam or t5large" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu9.log
B="python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline"
timeout 300 $B 2>&1 | tail -1 | tee gpurun_out/bench9.json | cut -c1-200
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches9_fp16x3.csv python tools/profile_step.py --precision fp16x3 > gpurun_out/prof9_fp16x3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches9_fp16x3.csv | tee gpurun_out/launch_summary9_fp16x3.txt | head -10
timeout 300 python bench.py --steps 2 --warmup 3 --precision fp16x3 --no-cpu-baseline --parity-queries 0 --batch 128 --beams 100 2>&1 | tail -1 | tee gpurun_out/bench9_c3.json | cut -c1-200
timeout 300 python bench.py --steps 2 --warmup 3 --precision fp16x3 --no-cpu-baseline --parity-queries 0 --model t5-large --batch 512 2>&1 | tail -1 | tee gpurun_out/bench9_c4.json | cut -c1-200
