#!/bin/bash
# A/B probe (run under gpurun): parity tests that cover the changed kernels, then short bench runs per env setting.
# Usage: scripts/r2_probe.sh <tag> "<pytest -k expr>" "<workloads>" "ENV1=.. ENV2=..|ENV1=..|..."
set -u
T=$1; K=$2; WLS=$3; SETS=$4; O=gpurun_out; mkdir -p $O
if [ -n "$K" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > $O/${T}_pytest.log 2>&1; echo "pytest rc $?"; tail -n 4 $O/${T}_pytest.log; fi
IFS='|' read -ra ARR <<< "$SETS"
i=0
for SET in "${ARR[@]}"; do
  for WL in $WLS; do
    env $SET timeout 300 python bench.py --no-cpu --workload $WL --steps 40 --warmup 3 > $O/${T}_${WL}_$i.json 2> $O/${T}_${WL}_$i.err
    python - "$O/${T}_${WL}_$i.json" "$SET" "$WL" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d['roofline']
    print(f"{sys.argv[3]:8s} [{sys.argv[2]}] ms_per_step {d['ms_per_step']:.4f} value {d['value']:.4g} e2e {d['e2e']['value']:.4g} f32 {d['e2e'].get('value_f32_rows')} kernel_ms {r.get('kernel_ms'):.4f} valid {d['counters']['valid_frac']}")
except Exception as e:
    print('FAILED',sys.argv[1:],e)
PY
  done
  i=$((i+1))
done
