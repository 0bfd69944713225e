#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-e1}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"estep_tc_kernel" -s 2 -c 1 -o gpurun_out/${TAG}_estep -f python tools/estep_repro.py 127000 768 100 4 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log; ls -la gpurun_out/${TAG}_estep.ncu-rep
