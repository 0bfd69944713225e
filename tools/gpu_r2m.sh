#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2m}
timeout 300 python -m pytest tests/test_gpu_kmeans.py -q -x -k "fused or host_features or full_size" > gpurun_out/${TAG}_pytest_fused.log 2>&1; echo "pytest fused rc=$?"; tail -25 gpurun_out/${TAG}_pytest_fused.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-extra > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "n1 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_n1.json'))
print('value', d['value'], 'split', d['split'], d.get('kmeans_pass'))
print('e2e', d['e2e'])
print('parity', d['parity'])
PY
tail -5 gpurun_out/${TAG}_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-extra --no-e2e --no-fused-em > gpurun_out/${TAG}_n1_unfused.json 2> gpurun_out/${TAG}_n1_unfused.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_n1_unfused.json'))
print('unfused value', d['value'], 'split', d['split'])
print('parity', d['parity'])
PY
