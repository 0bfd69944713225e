#!/bin/bash
# per-phase split of the round at N=1 and N=2 (event nodes inside the graphs)
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-e2e > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_n2.json 2> gpurun_out/${TAG}_n2.err; echo "n2 rc=$?"
python - <<PY
import json
for n in (1, 2):
    try:
        d=json.load(open('gpurun_out/${TAG}_n%d.json' % n))
        print(n, 'value', d['value'], 'split', d['split'])
        for k in ('c5','c4_vocab_shard','c4_rows'):
            if k in d: print('  ', k, d[k]['sharding'], {x:d[k][x] for x in ('ms_per_step','naming_ms','rest_ms','kernel_frac','phases_us_rank0')})
    except Exception as e:
        print('no line:', e)
PY
tail -3 gpurun_out/${TAG}_n1.err gpurun_out/${TAG}_n2.err
