mkdir -p gpurun_out
B="python bench.py --ncu-step --warmup 3"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launch_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_mlp_fused -s 8 -c 2 -o gpurun_out/r02_mlp -f $B > gpurun_out/r02_ncu_mlp.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_qkv_attn -s 10 -c 2 -o gpurun_out/r02_qkv_attn -f $B > gpurun_out/r02_ncu_qkv_attn.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_pool_(mma|stats)' -s 0 -c 2 -o gpurun_out/r02_pool -f $B > gpurun_out/r02_ncu_pool.log 2>&1
