#!/bin/bash
# Round-2 evidence pass (run under gpurun on one B200): GPU tests, the default bench line, the CPU arm,
# per-workload launch lists and one full ncu capture of the dominant kernels of each workload.
# Usage: scripts/r2_pass.sh <tag> [stages]   stages: any of t(ests) b(ench) r(eference) l(aunch lists) f(ull captures)
set -u
T=${1:-r2a}
ST=${2:-tbrlf}
O=gpurun_out
mkdir -p $O
if [[ $ST == *t* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc $?"; tail -n 3 $O/${T}_pytest.log
fi
if [[ $ST == *b* ]]; then
  timeout 600 python bench.py > $O/${T}_bench_all.json 2> $O/${T}_bench_all.err; echo "bench rc $?"; cut -c1-600 $O/${T}_bench_all.json
fi
if [[ $ST == *r* ]]; then
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "ref rc $?"
fi
if [[ $ST == *l* ]]; then
  for WL in configs edges blocks clutter; do
    CMD="python bench.py --steps 2 --warmup 3 --no-cpu --workload $WL"
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $O/${T}_${WL}_launches.csv $CMD > $O/${T}_${WL}_ncu_launches.log 2>&1; echo "launches $WL rc $?"
  done
fi
if [[ $ST == *f* ]]; then
  for WL in configs clutter; do
    CMD="python bench.py --steps 2 --warmup 3 --no-cpu --workload $WL"
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_wave' -s 15 -c 5 \
      -f -o $O/${T}_${WL}_prof $CMD > $O/${T}_${WL}_ncu_full.log 2>&1; echo "full $WL rc $?"
  done
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --workload edges"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_states_tile -s 6 -c 1 \
    -f -o $O/${T}_edges_prof $CMD > $O/${T}_edges_ncu_full.log 2>&1; echo "full edges rc $?"
fi
for R in $O/${T}_*_prof.ncu-rep; do
  ncu -i $R --page raw --csv > ${R%.ncu-rep}_raw.csv 2>/dev/null
done
du -sm $O; ls -la $O | tail -n 30
