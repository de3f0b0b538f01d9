#!/bin/bash
# A/B harness for one GPU box:  tools/ab_bench.sh VAR "v1 v2 ..." [bench args]
# runs bench.py --no-cpu once per value of the environment toggle VAR and prints one summary line each
# (submaps/s resident, end to end, ms/step, ms per kernel family).  Toggles of the product path:
#   HFL_ATTN_V=1|2          window attention: one head per warp | two heads per warp (default: v3, 8 heads per pass)
#   HFL_ATTN_WPH=1|2        v3: warps per head (default: 2 only when the tables allow one CTA per SM)
#   HFL_ATTN_CTAS=3         EXPERIMENTAL (kernel tests pass, speed unmeasured): three windows per SM, compact tables, 80 registers
#   HFL_GEMM_DENSE_TMA=0    dense GEMMs: A by cp.async producers instead of TMA
#   HFL_GEMM_PLAIN2=0       +bias / bf16-store epilogue without the tensor-memory load prefetch
#   HFL_GEMM_DENSE320=1     EXPERIMENTAL (kernel tests pass, speed unmeasured): 320-thread dense GEMM instantiation, 168 registers
#   HFL_FUSED_MLP=""|128|256|128,256   channel widths that use the fused MLP kernel
#   HFL_LEVEL_STREAMS=1     pyramid levels of an H-OSA block on three streams
#   HFL_LOADER_THREADS=N    eval file loader threads (1 = serial)
VAR=$1; shift
VALS=$1; shift
mkdir -p gpurun_out
for v in $VALS; do
  env "$VAR=$v" python bench.py --no-cpu --steps 10 --warmup 3 "$@" > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.err
  python - "$VAR" "$v" <<'P'
import json, sys
var, v = sys.argv[1:3]
try:
    d = json.load(open(f'gpurun_out/ab_{var}_{v}.json'))
    print(f'{var}={v}', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 2),
          {k: round(x['ms_per_step'], 2) for k, x in d['roofline_by_kernel'].items()})
except Exception as e:
    print(f'{var}={v}', 'FAILED', e)
P
done
