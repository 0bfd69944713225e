#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2y}
timeout 600 python -m pytest tests/test_gpu_naming.py tests/test_gpu_scale.py tests/test_gpu_multirank.py -q -x -k "vote or round or rank or c1" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_diag.py --run small_kernels > gpurun_out/${TAG}_small_kernels.txt 2>&1; tail -6 gpurun_out/${TAG}_small_kernels.txt
timeout 600 python bench.py --no-cpu-baseline --no-torch-baseline --no-extra --no-e2e > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_n1.json'))
print('value', d['value'], 'split', d['split'], d['parity']['voted_sha'], d['parity']['vote_counts_sha'])
PY
