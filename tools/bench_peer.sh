#!/usr/bin/env bash
# peer-memory exchange at N GPUs: equivalence check, then the step under dense NCCL / peer (in-step) / peer after the step
N=${1:-2}; shift
MODES=${@:-"dense peer peer_after inside"}
[ -n "$SKIP_CHECK" ] || timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29901 tools/check_engine_dp.py 2>&1 | grep -v "^W\|^\*\*\|UserWarning\|return func\|OMP_NUM" | tail -14
port=29910
for m in $MODES; do
  port=$((port+1))
  ex=$m; extra=
  case $m in inside) ex=inside;; peer_after) ex=peer; extra="MHE_ENGINE_PEER_AFTER=1";; peer_c3) ex=peer; extra="MHE_FUSED_BWD_CHUNKS=3";; peer_c4) ex=peer; extra="MHE_FUSED_BWD_CHUNKS=4";; peer_k1) ex=peer; extra="MHE_PEER_COPIES=1";; peer_k2) ex=peer; extra="MHE_PEER_COPIES=2";; peer_k8) ex=peer; extra="MHE_PEER_COPIES=8";; esac
  env MHE_BENCH_EXCHANGE=$ex $extra timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 30 --warmup 5 --no-configs --profile > gpurun_out/peer_${N}gpu_$m.json 2> gpurun_out/peer_${N}gpu_$m.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/peer_${N}gpu_$m.json'))
    print("N=$N mode=$m ms/step %.4f" % d["ms_per_step"])
except Exception as e:
    print('N=$N mode=$m FAILED', e)
    import subprocess; print(subprocess.run('tail -5 gpurun_out/peer_${N}gpu_$m.err', shell=True, capture_output=True, text=True).stdout)
PY
done
