#!/bin/bash
# Profiling pass (run under gpurun on one B200): launch list + one full capture of the dominant kernel.
# Usage: scripts/gpu_profile.sh <tag> [workload]
set -u
TAG=${1:-r1}
WL=${2:-configs}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --workload $WL"
KREGEX=k_states_tile
[ "$WL" = edges ] && KREGEX=k_states_tile
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_${WL}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_${WL}_launches.csv $CMD > gpurun_out/${TAG}_${WL}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 2 \
    -f -o gpurun_out/${TAG}_${WL}_prof $CMD > gpurun_out/${TAG}_${WL}_ncu_full.log 2>&1
tail -n 3 gpurun_out/${TAG}_${WL}_ncu_launches.log; tail -n 3 gpurun_out/${TAG}_${WL}_ncu_full.log
