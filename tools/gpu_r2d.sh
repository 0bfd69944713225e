#!/bin/bash
# Round 2: naming kernel with the linear work partition + one-round-trip exact scan: parity suite, shape sweep, item timeline, bench
mkdir -p gpurun_out
TAG=${1:-r2d}
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_diag.py --run naming_scale > gpurun_out/${TAG}_naming_scale.txt 2>&1; head -16 gpurun_out/${TAG}_naming_scale.txt
timeout 300 python tools/gpu_diag.py --run name_items > gpurun_out/${TAG}_name_items.txt 2>&1; head -75 gpurun_out/${TAG}_name_items.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
for k in ('value','clocks','e2e','roofline','split','sustained'): print(k, d.get(k))
for k in ('c5','c4_vocab_shard','c4_rows'): print(k, {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac')} if k in d else None)
PY
tail -5 gpurun_out/${TAG}_bench.err
