#!/usr/bin/env bash
# One compute-sanitizer pass over the smoke step: memcheck | racecheck | synccheck | initcheck.
# The cluster-fused flow kernels synchronise through hand-written mbarrier protocols (remote release-arrivals instead of a separate
# fence.acq_rel.cluster, csrc/flow_fused.cu:70-80), so racecheck / synccheck are the tools that have to bless them.
#   usage (GPU box, ONE tool per gpurun call - the pool's rule, B200_PROFILING.md):  tools/sanitize.sh <tool> [outdir]
#   log: <outdir>/sanitize_<tool>.log, one summary line on stdout; exit status = the tool's (3 = errors reported)
set -u
cd "$(dirname "$0")/.."
tool=${1:?usage: tools/sanitize.sh memcheck|racecheck|synccheck|initcheck [outdir]}
OUT=${2:-gpurun_out}
mkdir -p "$OUT"
SAN=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
log="$OUT/sanitize_${tool}.log"
extra=""
[ "$tool" = memcheck ] && extra="--leak-check no"
[ "$tool" = racecheck ] && extra="--racecheck-report all"
# bounded: a hang under instrumentation must not take the box down
timeout "${SANITIZE_TIMEOUT:-900}" "$SAN" --tool "$tool" $extra --error-exitcode 3 --print-limit 30 \
    python -c 'import __graft_entry__ as g; g.smoke()' > "$log" 2>&1
rc=$?
summary=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
echo "sanitize $tool: exit $rc, ${summary:-no summary line} ($(grep -c 'smoke:' "$log") smoke line(s))"
exit $rc
