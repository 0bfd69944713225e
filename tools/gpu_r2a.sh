#!/bin/bash
# Round 2, first GPU visit: parity suite (incl. the new scale / multi-rank tests), bench with the extra blocks, single-GPU
# proxies of the per-rank shapes with and without the suspend-time hint, ncu launch list + full capture.
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -8 > gpurun_out/${TAG}_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
for H in 0 2000 50000; do
  SCD_NAME_WAIT_HINT_NS=$H timeout 300 python tools/gpu_diag.py --run naming_scale > gpurun_out/${TAG}_naming_scale_h$H.txt 2>&1
  head -8 gpurun_out/${TAG}_naming_scale_h$H.txt
done
timeout 300 python tools/gpu_diag.py --run small_kernels > gpurun_out/${TAG}_small_kernels.txt 2>&1; cat gpurun_out/${TAG}_small_kernels.txt
timeout 300 python tools/gpu_diag.py --run name_prof > gpurun_out/${TAG}_name_prof.txt 2>&1; head -24 gpurun_out/${TAG}_name_prof.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-clocks --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"name_topk_kernel|estep_tc_kernel|segment_sum_kernel|vote_kernel|label_scatter|label_hist" -s 12 -c 7 -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-clocks --no-extra --no-graph > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
