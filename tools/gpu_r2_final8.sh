#!/bin/bash
# Round 2 evidence, 8 GPUs: the driver's bench line at N = 1, 2, 4, 8 on ONE box (C2 headline + c5 + c4 blocks + parity vs a
# 1-rank recomputation), the NCCL-exchange variant at N = 8 beside it, and the two-GPU peer-exchange test
mkdir -p gpurun_out
TAG=${1:-r2q}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/${TAG}_pytest_peer.log 2>&1; echo "peer test rc=$?"; tail -3 gpurun_out/${TAG}_pytest_peer.log
run() {  # n exchange suffix
  local n=$1 ex=$2 sfx=$3
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 20 --warmup 5 --exchange $ex > gpurun_out/${TAG}_n${n}${sfx}.json 2> gpurun_out/${TAG}_n${n}${sfx}.err
  fi
  echo "N=$n exchange=$ex rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_n${n}${sfx}.json'))
    print('  value', d['value'], 'split', {k:v for k,v in d['split'].items()}, 'frac', d['roofline']['frac'], 'parity', d['parity'].get('equals_n1'))
    for k in ('c5','c4_vocab_shard','c4_rows','c4_grid_2d'):
        if k in d: print('  ', k, d[k]['sharding'], {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac')}, d[k]['parity'].get('equals_n1'), d[k].get('phases_us_rank0'))
except Exception as e:
    print('  no line:', e)
PY
  tail -2 gpurun_out/${TAG}_n${n}${sfx}.err
}
run 8 peer ""
run 4 peer ""
run 2 peer ""
run 1 peer ""
run 8 nccl "_nccl"
