#!/bin/bash
# Multi-GPU visit: N-rank bench lines (strong scaling of C2), optional C5, reference arm
mkdir -p gpurun_out
TAG=${1:-m}; N=${2:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in $(seq 1 $N); do
  if [ $n -eq 1 ] || [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 30 --warmup 3 > gpurun_out/${TAG}_n$n.json 2> gpurun_out/${TAG}_n$n.err
    fi
    echo "N=$n rc=$?"; tail -c 1500 gpurun_out/${TAG}_n$n.json; tail -3 gpurun_out/${TAG}_n$n.err
  fi
done
if [ "$3" == "c5" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 5 --warmup 3 --config C5 > gpurun_out/${TAG}_c5_n$N.json 2> gpurun_out/${TAG}_c5_n$N.err
  echo "C5 N=$N rc=$?"; tail -c 1500 gpurun_out/${TAG}_c5_n$N.json; tail -3 gpurun_out/${TAG}_c5_n$N.err
fi
