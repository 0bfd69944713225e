#!/bin/bash
# Round 2, second visit: the whole parity suite (no -x), per-item timeline of the naming kernel, bench
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 1800 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_diag.py --run name_items > gpurun_out/${TAG}_name_items.txt 2>&1; cat gpurun_out/${TAG}_name_items.txt | head -90
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/'+'r2b'+'_bench.json'))
for k in ('value','clocks','e2e','roofline','split','sustained'): print(k, d.get(k))
for k in ('c5','c4_vocab_shard','c4_rows'): print(k, {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac')} if k in d else None)
PY
tail -5 gpurun_out/${TAG}_bench.err
