#!/bin/bash
# 8 GPUs: peer test, then the bench at N = 8 and 4 after the parallel peer loads / register-resident merge
mkdir -p gpurun_out
TAG=${1:-r2r}
timeout 300 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/${TAG}_pytest_peer.log 2>&1; echo "peer test rc=$?"; tail -15 gpurun_out/${TAG}_pytest_peer.log
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${TAG}_n${n}.json 2> gpurun_out/${TAG}_n${n}.err
  echo "N=$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_n${n}.json'))
    print('  value', d['value'], 'split', {k:v for k,v in d['split'].items()}, 'frac', d['roofline']['frac'], 'parity', d['parity'].get('equals_n1'))
    for k in ('c5','c4_vocab_shard','c4_rows','c4_grid_2d'):
        if k in d: print('  ', k, d[k]['sharding'], {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac')}, d[k]['parity'].get('equals_n1'), d[k].get('phases_us_rank0'))
except Exception as e:
    print('  no line:', e)
PY
  tail -2 gpurun_out/${TAG}_n${n}.err
done
