#!/bin/bash
# Round 2, multi-GPU visit: the driver's bench line at N ranks (C2 headline + c5 / c4 blocks + parity vs a 1-rank recomputation)
mkdir -p gpurun_out
TAG=${1:-r2c}; shift
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in "$@"; do
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${TAG}_n$n.json 2> gpurun_out/${TAG}_n$n.err
  fi
  echo "N=$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_n$n.json'))
    print('value', d['value'], 'split', d['split'], 'frac', d['roofline']['frac'], 'parity', {k:v for k,v in d['parity'].items() if 'mismatch' in k or k=='equals_n1'})
    for k in ('c5','c4_vocab_shard','c4_rows','c4_grid_2d'):
        if k in d: print(k, d[k]['sharding'], {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac')}, d[k]['parity'].get('equals_n1'))
except Exception as e:
    print('no line:', e)
PY
  tail -4 gpurun_out/${TAG}_n$n.err
done
