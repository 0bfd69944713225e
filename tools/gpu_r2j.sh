#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2j}
timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_diag.py --run small_kernels > gpurun_out/${TAG}_small_kernels.txt 2>&1; tail -6 gpurun_out/${TAG}_small_kernels.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "n1 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_n1.json'))
print('value', d['value'], 'split', d['split'])
print('e2e', d['e2e'])
for k in ('c5','c4_rows'):
    if k in d: print('  ', k, {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac','phases_us_rank0')})
PY
tail -5 gpurun_out/${TAG}_n1.err
