#!/bin/bash
# GPU visit: E-step (TMEM-A variant) bring-up + counters, k-means / eval / naming parity tests, bench line
mkdir -p gpurun_out
TAG=${1:-v2}
timeout 300 python tools/gpu_diag.py kmeans estep_prof > gpurun_out/${TAG}_diag.log 2>&1
grep -E "pd n=|estep n=|issuer|converter|producer|epilogue|rc=|rror|trap|timed out" gpurun_out/${TAG}_diag.log | head -60
timeout 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_eval.py tests/test_gpu_constrained.py tests/test_gpu_naming.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
