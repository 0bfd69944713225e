#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --no-torch-baseline --no-extra > gpurun_out/r2x_n1.json 2> gpurun_out/r2x_n1.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2x_n1.json'))
print('value', d['value'], 'split', d['split'])
print('e2e', d['e2e'])
PY
tail -3 gpurun_out/r2x_n1.err
