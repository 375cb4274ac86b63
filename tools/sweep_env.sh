#!/bin/bash
# usage: tools/sweep_env.sh "VAR=val VAR2=val" "VAR=val" ...   -> one bench line (ms/step) per environment setting
for cfg in "$@"; do
  ms=$(env $cfg python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'])")
  echo "$cfg -> $ms ms/step"
done
