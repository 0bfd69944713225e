#!/bin/bash
# One GPU-box visit: diagnostics, parity tests, bench, ncu launch list and full captures -> gpurun_out/
mkdir -p gpurun_out
TAG=${1:-r}
timeout 300 python tools/gpu_diag.py kmeans naming_time > gpurun_out/${TAG}_diag.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-clocks > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"name_topk_kernel|estep_tc_kernel|segment_sum_kernel|vote_kernel" -s 8 -c 4 -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-clocks > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
head -40 gpurun_out/${TAG}_diag.log
