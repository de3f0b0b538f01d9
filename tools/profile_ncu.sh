#!/bin/bash
# ncu evidence for the round: launch list of one bench step + full capture of the top kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1110 -c 372 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gather_gemm -s 210 -c 8 \
    -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_window_attn -s 40 -c 2 \
    -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cpe_ln -s 40 -c 2 \
    -o gpurun_out/prof_cpe -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_cpe.log 2>&1
ls -la gpurun_out
