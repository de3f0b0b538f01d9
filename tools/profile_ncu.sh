#!/bin/bash
# ncu evidence for the round: launch list of ONE bench step (cudaProfilerStart/Stop brackets it,
# see bench.py --ncu-step) + full captures of the top kernel families of that same step and of the
# small kernels the roofline table names (octree neighbours, pooling, exact top-k).
mkdir -p gpurun_out
R=${1:-r02}
B="python bench.py --ncu-step --warmup 3"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${R}_launches.csv $B > gpurun_out/${R}_launch_bench.log 2>&1
cap() {  # name regex skip count
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 \
      -o gpurun_out/${R}_$1 -f $B > gpurun_out/${R}_ncu_$1.log 2>&1
}
cap mlp k_mlp_fused 8 2
cap qkv_attn k_qkv_attn 10 2
cap cpe k_cpe_ln 10 2
cap gemm k_gather_gemm 4 4
cap neigh k_neigh_child 0 2
cap pool 'k_pool_(mma|stats)' 0 2
ncu --set full --clock-control none --import-source on -k regex:k_knn -s 3 -c 1 -o gpurun_out/${R}_knn -f \
    python tools/knn_bench.py > gpurun_out/${R}_ncu_knn.log 2>&1
ls -la gpurun_out | grep ${R}_
