#!/bin/bash
# usage: tools/sweep_env_2gpu.sh "VAR=val ..." ...  -> ms/step of the 2-GPU bench per environment setting
for cfg in "$@"; do
  ms=$(env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 5 --profile 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'])")
  echo "$cfg -> $ms ms/step"
done
