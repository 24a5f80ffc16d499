#!/bin/bash
# Multi-GPU session (runs under `gpurun --gpus N`): the NCCL exchange test at world size 2, then bench.py at 1..N ranks.
# usage: profiles/tools/gpu_session_multi.sh <tag> <N>
TAG=${1:-r2}; N=${2:-2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_multigpu_nccl.py -q -m gpu -rs > $O/${TAG}_nccl_test.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_nccl_test.log
timeout 300 python bench.py --gpus 1 --quick > $O/${TAG}_scale_n1.json 2> $O/${TAG}_scale_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
        bench.py --gpus $n --steps 20 --warmup 5 > $O/${TAG}_scale_n${n}.json 2> $O/${TAG}_scale_n${n}.err
  fi
done
tail -3 $O/${TAG}_nccl_test.log; cat $O/${TAG}_scale_n*.json | cut -c1-400
