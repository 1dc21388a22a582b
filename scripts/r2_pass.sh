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
  for WL in configs edges clutter; do
    CMD="python bench.py --steps 2 --warmup 3 --no-cpu --workload $WL"
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $O/${T}_${WL}_launches.csv $CMD > $O/${T}_${WL}_ncu_launches.log 2>&1; echo "launches $WL rc $?"
  done
fi
if [[ $ST == *f* ]]; then
  # full captures: prof_step.py runs three ungraphed steps of one workload (the kernels are the bench's)
  cap() { # <name> <workload> <kernel regex> <skip> <count>
    TMK_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 \
      -f -o $O/${T}_$1 python scripts/prof_step.py $2 3 > $O/${T}_$1_ncu.log 2>&1; echo "full $1 rc $?"
    : > $O/${T}_$1_ncu_summary.txt
    for ((i = 0; i < $5; i++)); do python scripts/ncu_summary.py $O/${T}_$1.ncu-rep $i 0 2>&1 | grep -v "^instruction share\|^ *[0-9.]*%" >> $O/${T}_$1_ncu_summary.txt; echo >> $O/${T}_$1_ncu_summary.txt; done
    ncu -i $O/${T}_$1.ncu-rep --page raw --csv > $O/${T}_$1_raw.csv 2>/dev/null
  }
  cap configs_main configs 'k_wave_fk|k_wave_export|k_wave_bits' 3 3
  cap configs_broad configs 'k_wave_broad|k_wave_quick' 4 4
  cap configs_tail configs 'k_conv_items|k_mesh_tasks|k_tri_items|k_recheck_items' 4 4
  cap clutter clutter 'k_wave_fk|k_wave_broad|k_wave_quick' 0 3
  cap edges edges 'k_states_tile' 14 1
  cap edges_heavy edges 'k_heavy_mixed' 4 1
  for K in k_wave_broad k_wave_quick; do python scripts/ncu_lines.py $O/${T}_configs_broad.ncu-rep $K 14 > $O/${T}_lines_$K.txt 2>&1; done
  for K in k_wave_fk; do python scripts/ncu_lines.py $O/${T}_configs_main.ncu-rep $K 14 > $O/${T}_lines_$K.txt 2>&1; done
  for K in k_mesh_tasks k_tri_items; do python scripts/ncu_lines.py $O/${T}_configs_tail.ncu-rep $K 14 > $O/${T}_lines_$K.txt 2>&1; done
  python scripts/ncu_lines.py $O/${T}_edges.ncu-rep k_states_tile 14 > $O/${T}_lines_k_states_tile.txt 2>&1
  rm -f $O/${T}_*.ncu-rep   # the summaries travel back, the reports (60+ MB) do not
fi
du -sm $O; ls -la $O | tail -n 40
