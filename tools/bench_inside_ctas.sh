#!/usr/bin/env bash
# in-step bucketed exchange with NCCL limited to few CTAs (so its kernels fit beside the cluster kernels):
#   tools/bench_inside_ctas.sh N "ctas..." [chunks]
N=${1:-2}; CT=${2:-"2 4 8"}; CH=${3:-2}
port=29700
run() {  # label, env...
  port=$((port+1)); label=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 30 --warmup 5 --no-configs --profile > gpurun_out/inside_${N}gpu_$label.json 2> gpurun_out/inside_${N}gpu_$label.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/inside_${N}gpu_$label.json'))
    print('N=$N $label ms/step %.4f' % d['ms_per_step'])
except Exception as e:
    print('N=$N $label FAILED', e)
PY
}
run dense MHE_BENCH_EXCHANGE=dense
for c in $CT; do
  run inside_ctas$c MHE_BENCH_EXCHANGE=inside NCCL_MAX_CTAS=$c MHE_FUSED_BWD_CHUNKS=$CH
done
