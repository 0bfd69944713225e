#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2s}
timeout 200 python tools/naming_stress.py 1 75 > gpurun_out/${TAG}_naming_stress.txt 2>&1; echo "naming stress rc=$?"; tail -3 gpurun_out/${TAG}_naming_stress.txt
timeout 400 python tools/naming_shapes_big.py > gpurun_out/${TAG}_naming_shapes_big.txt 2>&1; echo "shapes rc=$?"; cat gpurun_out/${TAG}_naming_shapes_big.txt
timeout 200 python tools/estep_stress2.py 1 60 > gpurun_out/${TAG}_estep_stress.txt 2>&1; echo "estep stress rc=$?"; tail -3 gpurun_out/${TAG}_estep_stress.txt
timeout 600 python -m pytest tests/test_gpu_naming.py tests/test_gpu_scale.py -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-clocks --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1
grep -E "topk_merge|vote_kernel" gpurun_out/${TAG}_launches.csv | tail -4 | cut -c1-60,200-400
