#!/bin/bash
# GPU visit: E-step bring-up + per-role cycle counters and the k-means parity tests
mkdir -p gpurun_out
TAG=${1:-v3}
timeout 300 python tools/gpu_diag.py kmeans estep_prof > gpurun_out/${TAG}_diag.log 2>&1
grep -E "pd n=|estep n=|issuer|converter|producer|epilogue|rc=|rror|trap|timed out" gpurun_out/${TAG}_diag.log | head -70
timeout 600 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_constrained.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
