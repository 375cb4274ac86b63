#!/usr/bin/env bash
# the data-parallel step under every gradient-exchange mode at N GPUs:  tools/bench_exchange_modes.sh N [modes...]
N=${1:-2}; shift
MODES=${@:-dense factored inside}
port=29600
for m in $MODES; do
  port=$((port+1))
  MHE_BENCH_EXCHANGE=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 30 --warmup 5 --no-configs --profile > gpurun_out/exch_${N}gpu_$m.json 2> gpurun_out/exch_${N}gpu_$m.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/exch_${N}gpu_$m.json'))
    print('N=$N mode=$m ms/step %.4f' % d['ms_per_step'])
except Exception as e:
    print('N=$N mode=$m FAILED', e)
PY
done
