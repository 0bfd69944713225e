#!/bin/bash
# Round 2 evidence, one GPU: parity suite, the driver's bench line (all blocks), reference arm, ncu launch list + full capture
mkdir -p gpurun_out
TAG=${1:-r2p}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/${TAG}_smi.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -14 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
for k in ('value','clocks','e2e','roofline','split','sustained','cpu_baseline','torch_cuda_baseline','gpu_launches'): print(k, d.get(k))
for k in ('c5','c4_vocab_shard','c4_rows'): print(k, {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac')} if k in d else None)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"; cat gpurun_out/${TAG}_bench_reference.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-clocks --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"name_topk_kernel|estep_tc_kernel|segment_sum_kernel|vote_kernel|topk_merge|label_scatter|label_hist|finalize" -s 18 -c 9 -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-clocks --no-extra --no-graph > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
