set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "golden or t5base_search or long_docid" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu7.log
B="python bench.py --steps 3 --warmup 3 --precision fp16x3 --no-cpu-baseline"
RB200_LANES=1 RB200_SELF=v3 timeout 300 $B --parity-queries 0 2>&1 | tail -1 | tee gpurun_out/bench7_l1_v3.json | cut -c1-200
RB200_LANES=1 timeout 300 $B --parity-queries 0 2>&1 | tail -1 | tee gpurun_out/bench7_l1.json | cut -c1-200
RB200_SELF=v3 timeout 300 $B 2>&1 | tail -1 | tee gpurun_out/bench7_l2_v3.json | cut -c1-200
timeout 300 $B 2>&1 | tail -1 | tee gpurun_out/bench7_l2.json | cut -c1-200
RB200_LANES=1 RB200_SELF=v3 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"self_attn" -s 340 -c 2 -o gpurun_out/self7_v3 python tools/profile_step.py --precision fp16x3 > gpurun_out/prof7a.log 2>&1
RB200_LANES=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"self_attn" -s 340 -c 2 -o gpurun_out/self7_diet python tools/profile_step.py --precision fp16x3 > gpurun_out/prof7b.log 2>&1
ls -la gpurun_out/*.ncu-rep
