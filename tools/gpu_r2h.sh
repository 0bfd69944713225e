#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2h}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/red_bench tools/red_scatter_bench.cu && timeout 120 /tmp/red_bench > gpurun_out/${TAG}_red_scatter.txt 2>&1; cat gpurun_out/${TAG}_red_scatter.txt
timeout 900 python -m pytest tests/test_gpu_naming.py tests/test_gpu_scale.py tests/test_gpu_multirank.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_diag.py --run small_kernels > gpurun_out/${TAG}_small_kernels.txt 2>&1; cat gpurun_out/${TAG}_small_kernels.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-e2e > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "n1 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_n1.json'))
print('value', d['value'], 'split', d['split'])
for k in ('c5','c4_rows'):
    if k in d: print('  ', k, {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac','phases_us_rank0')})
PY
