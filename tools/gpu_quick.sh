#!/bin/bash
# Quick GPU visit: bring-up diagnostics, parity tests of the two contraction kernels, kernel timings (+ optional bench line)
mkdir -p gpurun_out
TAG=${1:-q}
timeout 300 python tools/gpu_diag.py naming_tiny naming_shapes > gpurun_out/${TAG}_shapes.log 2>&1
grep -E "naming n=|bad rows|rc=|Error|error|trap|timed out" gpurun_out/${TAG}_shapes.log | head -40
timeout 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_naming.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_diag.py naming_time name_prof > gpurun_out/${TAG}_diag.log 2>&1
grep -E "name_topk C2|estep C2|mstep C2|sample idx|issuer|epi |tiles" gpurun_out/${TAG}_diag.log
if [ "$2" == "bench" ]; then
  timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
fi
