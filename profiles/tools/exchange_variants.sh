#!/bin/bash
O=gpurun_out; mkdir -p $O
for MODE in overlap nooverlap peer; do
  EX="--exchange compact"; [ $MODE = nooverlap ] && EX="--exchange compact --no-overlap"; [ $MODE = peer ] && EX="--exchange peer"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 4 --steps 20 --warmup 5 $EX > $O/r2s_n4_${MODE}.json 2> $O/r2s_n4_${MODE}.err
  python - <<PY
import json
try:
    l=json.loads([x for x in open("$O/r2s_n4_${MODE}.json").read().splitlines() if x.startswith("{")][-1])
    e=l["exchange"]
    print("$MODE", round(l["ms_per_step"],3), "mode", e["mode"], e.get("peer_fallback_reason"), "no-coll", e.get("step_without_collectives_ms"), "phases", e.get("phases_ms"))
except Exception as ex:
    print("$MODE failed", ex); print(open("$O/r2s_n4_${MODE}.err").read()[-600:])
PY
done
