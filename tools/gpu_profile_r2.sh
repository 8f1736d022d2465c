# Round-2 profiling pass (run under gpurun, one GPU): launch lists of one search and ncu --set full captures of the
# dominant kernels. Reports land in gpurun_out/; tools/ncu_extract.py summarises them into profiles/.
set -x
mkdir -p gpurun_out
P="python tools/profile_step.py --precision fp16x3"
NCU="ncu --profile-from-start off --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_fp16x3_c2_raw.csv $P > gpurun_out/prof0.log 2>&1
RB200_SELF_TAIL=v1 timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_fp16x3_c2_selfv1_raw.csv $P > gpurun_out/prof0b.log 2>&1
# GEMM launch order in one search: encoder 60, steps 0..3 4 x 73, then the forced tail (72 big GEMMs + LM heads)
timeout 600 $NCU --set full --import-source on -k regex:gemm_sm100_2cta -s 352 -c 6 -o gpurun_out/r02_gemm_tail $P > gpurun_out/prof1.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:gemm_sm100_2cta -s 206 -c 6 -o gpurun_out/r02_gemm_step $P > gpurun_out/prof2.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:self_attn_tail -c 1 -o gpurun_out/r02_self_tail $P > gpurun_out/prof3.log 2>&1
RB200_SELF_TAIL=v1 timeout 600 $NCU --set full --import-source on -k regex:self_attn_tail -c 1 -o gpurun_out/r02_self_tail_v1 $P > gpurun_out/prof3b.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:cross_attn_mma16 -c 1 -o gpurun_out/r02_cross_tail $P > gpurun_out/prof4.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:rmsnorm -s 150 -c 2 -o gpurun_out/r02_rmsnorm_tail $P > gpurun_out/prof5.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/prof*.log
