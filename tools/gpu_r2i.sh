#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2i}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fused_proxy tools/fused_em_proxy.cu && timeout 120 /tmp/fused_proxy > gpurun_out/${TAG}_fused_proxy.txt 2>&1; cat gpurun_out/${TAG}_fused_proxy.txt
timeout 900 python -m pytest tests/test_gpu_naming.py tests/test_gpu_scale.py tests/test_gpu_multirank.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_diag.py --run small_kernels > gpurun_out/${TAG}_small_kernels.txt 2>&1; tail -6 gpurun_out/${TAG}_small_kernels.txt
