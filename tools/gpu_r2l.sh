#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2l}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vote_kernel" -s 2 -c 1 -o gpurun_out/${TAG}_vote -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-clocks --no-extra --no-graph > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log; ls -la gpurun_out/${TAG}_vote*
