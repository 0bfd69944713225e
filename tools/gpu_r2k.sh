#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2k}
timeout 600 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_naming.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-extra > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "n1 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_n1.json'))
print('value', d['value'], 'split', d['split'])
print('e2e', d['e2e'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vote_kernel" -s 2 -c 1 -o gpurun_out/${TAG}_vote -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-clocks --no-extra --no-graph > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
