#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2t}
timeout 300 python -m pytest tests/test_gpu_naming.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for seed in 1 2 3; do
  timeout 200 python tools/naming_stress.py $seed 90 > gpurun_out/${TAG}_naming_stress_$seed.txt 2>&1; echo "naming stress seed $seed rc=$?"; tail -2 gpurun_out/${TAG}_naming_stress_$seed.txt
done
