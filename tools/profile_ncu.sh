#!/bin/bash
# ncu evidence for the round: launch list of ONE bench step (cudaProfilerStart/Stop brackets it,
# see bench.py --ncu-step) + a full capture of the top kernel families of that same step.
mkdir -p gpurun_out
B="python bench.py --ncu-step --warmup 3"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv $B > gpurun_out/launch_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_mlp_fused -s 8 -c 2 \
    -o gpurun_out/prof_mlp -f $B > gpurun_out/ncu_mlp.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_gather_gemm -s 40 -c 8 \
    -o gpurun_out/prof_gemm -f $B > gpurun_out/ncu_gemm.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_window_attn -s 10 -c 2 \
    -o gpurun_out/prof_attn -f $B > gpurun_out/ncu_attn.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_cpe_ln -s 10 -c 2 \
    -o gpurun_out/prof_cpe -f $B > gpurun_out/ncu_cpe.log 2>&1
ls -la gpurun_out
