#!/usr/bin/env bash
# config-2 step (value leg only) under a few schedule switches
run() { echo -n "$* : "; env "$@" python bench.py --steps 30 --warmup 5 --profile 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.read())['ms_per_step'])"; }
run A=0
run MHE_COND_DIRECT=0
run MHE_ENGINE_PIPELINED_COND_BWD=1
run MHE_COND_DIRECT=0 MHE_ENGINE_PIPELINED_COND_BWD=1
run MHE_FUSED_BWD_CHUNKS=3
run MHE_FUSED_BWD_CHUNKS=1
run MHE_FUSED_HEAD_START_NS=0
