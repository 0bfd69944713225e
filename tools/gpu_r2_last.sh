#!/bin/bash
# last visit of the round: the driver's checks on the final build (parity suite, smoke, the bare bench line)
mkdir -p gpurun_out
TAG=${1:-r2z}
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -10 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
for k in ('value','clocks','e2e','roofline','split','gpu_launches'): print(k, d.get(k))
PY
