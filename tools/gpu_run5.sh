set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "golden or t5base_search or long_docid" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu5.log
timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench5_fp16x3.json | cut -c1-300
RB200_XATTN_Q=smem timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline --parity-queries 0 2>&1 | tail -1 | tee gpurun_out/bench5_fp16x3_qsmem.json | cut -c1-300
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches5_fp16x3.csv python tools/profile_step.py --precision fp16x3 > gpurun_out/prof5_fp16x3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches5_fp16x3.csv | tee gpurun_out/launch_summary5_fp16x3.txt | head -9
RB200_XATTN_Q=smem timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches5q_fp16x3.csv python tools/profile_step.py --precision fp16x3 --steps-only 12 > gpurun_out/prof5q_fp16x3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches5q_fp16x3.csv | grep cross
timeout 100 python tools/gemm_bench.py --precision fp16x3 2>&1 | tee gpurun_out/gemm_bench5.txt
