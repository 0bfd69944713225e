#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r3a}
timeout 300 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/${TAG}_pytest_peer.log 2>&1; echo "peer test rc=$?"; tail -12 gpurun_out/${TAG}_pytest_peer.log
timeout 300 python -m pytest tests/test_gpu_naming.py tests/test_gpu_scale.py tests/test_gpu_multirank.py -q -x -k "vote or round or rank or c1" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 > gpurun_out/${TAG}_n2.json 2> gpurun_out/${TAG}_n2.err; echo "N=2 rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_n2.json'))
    print('  value', d['value'], 'split', d['split'], 'parity', {k:v for k,v in d['parity'].items() if 'mismatch' in k or k=='equals_n1'})
    for k in ('c5','c4_vocab_shard','c4_rows'):
        if k in d: print('  ', k, {x:d[k][x] for x in ('ms_per_step','rest_ms')}, d[k]['parity'].get('equals_n1'), d[k].get('phases_us_rank0'))
except Exception as e:
    print('  no line:', e)
PY
tail -3 gpurun_out/${TAG}_n2.err
