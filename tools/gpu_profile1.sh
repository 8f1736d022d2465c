set -x
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tf32x3.csv python tools/profile_step.py --precision tf32x3 > gpurun_out/prof1.log 2>&1
tail -3 gpurun_out/prof1.log
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_sm100 -s 200 -c 7 -o gpurun_out/gemm_tf32x3 python tools/profile_step.py --precision tf32x3 --steps-only 4 > gpurun_out/prof2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"self_attn|cross_attn|beam_step|rmsnorm" -s 300 -c 8 -o gpurun_out/misc_tf32x3 python tools/profile_step.py --precision tf32x3 --steps-only 6 > gpurun_out/prof3.log 2>&1
ls -la gpurun_out
