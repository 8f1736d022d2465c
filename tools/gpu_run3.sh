set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "golden or t5base_search or long_docid" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu3.log
timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench3_fp16x3.json | cut -c1-300
RB200_PDL=1 timeout 300 python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench3_fp16x3_pdl.json | cut -c1-300
RB200_GEMM_TRACE=1 timeout 200 python tools/gemm_bench.py --precision fp16x3 --shapes "o64:768:64:1,o:768:768:1,cq:768:768:0,wi:3072:768:2" 2>&1 | tee gpurun_out/gemm_trace.txt
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches3_fp16x3.csv python tools/profile_step.py --precision fp16x3 > gpurun_out/prof3_fp16x3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches3_fp16x3.csv | tee gpurun_out/launch_summary3_fp16x3.txt | head -9
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"attn_warp" -s 190 -c 4 -o gpurun_out/attn3 python tools/profile_step.py --precision fp16x3 --steps-only 20 > gpurun_out/prof3b.log 2>&1
ls -la gpurun_out/*.ncu-rep
