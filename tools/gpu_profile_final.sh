set -x
mkdir -p gpurun_out
P="python tools/profile_step.py --precision fp16x3"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_sm100_2cta -s 356 -c 6 -o gpurun_out/r01_gemm_tail $P > gpurun_out/proff1.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_sm100_2cta -s 140 -c 6 -o gpurun_out/r01_gemm_step $P > gpurun_out/proff2.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:self_attn_tail -c 1 -o gpurun_out/r01_self_tail $P > gpurun_out/proff3.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:cross_attn_warp -s 60 -c 1 -o gpurun_out/r01_cross_tail $P > gpurun_out/proff4.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_final.json | cut -c1-300
