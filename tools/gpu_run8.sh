set -x
mkdir -p gpurun_out
RB200_SELF=smem timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "golden or t5base_search or long_docid" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu8.log
B="python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline"
RB200_SELF=smem timeout 300 $B 2>&1 | tail -1 | tee gpurun_out/bench8_smem.json | cut -c1-200
timeout 300 $B --parity-queries 0 2>&1 | tail -1 | tee gpurun_out/bench8_reg.json | cut -c1-200
RB200_SELF=smem timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches8_fp16x3.csv python tools/profile_step.py --precision fp16x3 > gpurun_out/prof8_fp16x3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches8_fp16x3.csv | tee gpurun_out/launch_summary8_fp16x3.txt | head -9
