#!/bin/bash
# Round 2, 2-GPU visit: peer-memory exchange test, then the bench at N=2 with both exchanges
mkdir -p gpurun_out
TAG=${1:-r2f}
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/${TAG}_pytest.log
for ex in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex > gpurun_out/${TAG}_n2_$ex.json 2> gpurun_out/${TAG}_n2_$ex.err
  echo "exchange=$ex rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_n2_$ex.json'))
    print('value', d['value'], 'split', d['split'], 'frac', d['roofline']['frac'], 'parity', {k:v for k,v in d['parity'].items() if 'mismatch' in k or k=='equals_n1'})
    for k in ('c5','c4_vocab_shard','c4_rows','c4_grid_2d'):
        if k in d: print(k, d[k]['sharding'], {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac')}, d[k]['parity'].get('equals_n1'))
except Exception as e:
    print('no line:', e)
PY
  tail -5 gpurun_out/${TAG}_n2_$ex.err
done
