python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -1
python bench.py --no-cpu --steps 10 --warmup 3 > gpurun_out/ab_3.json 2>gpurun_out/ab_3.err
python -c "
import json; d=json.load(open('gpurun_out/ab_3.json')); print(round(d['value']), round(d['e2e']['value']), d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline_by_kernel'].items()})"
