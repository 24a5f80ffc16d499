#!/bin/bash
O=gpurun_out; mkdir -p $O
for MODE in overlap nooverlap allreduce; do
  EX="--exchange compact"; [ $MODE = nooverlap ] && EX="--exchange compact --no-overlap"; [ $MODE = allreduce ] && EX="--exchange allreduce"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 20 --warmup 5 $EX > $O/r2n_n2_${MODE}.json 2> $O/r2n_n2_${MODE}.err
  python - <<PY
import json
l=json.loads([x for x in open("$O/r2n_n2_${MODE}.json").read().splitlines() if x.startswith("{")][-1])
print("$MODE", l["ms_per_step"], l["e2e_resident"]["ms_per_step"], l["exchange"])
PY
done
