#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kmeans.py -q -x -k "kpp" --durations=5 > gpurun_out/r2w_pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2w_pytest.log
