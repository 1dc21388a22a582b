#!/bin/bash
# per-kernel times of one configuration batch at several batch sizes (launch list under ncu)
# Usage: scripts/r2_sizes.sh <tag> "<log2 sizes>" "ENVSET|ENVSET|..." [pytest -k expr]
set -u
T=$1; SIZES=${2:-"16 18 20 22"}; SETS=${3:-"TMK_PARTS=1"}; K=${4:-}; O=gpurun_out; mkdir -p $O
if [ -n "$K" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > $O/${T}_pytest.log 2>&1; echo "pytest rc $?"; tail -n 4 $O/${T}_pytest.log; fi
IFS='|' read -ra ARR <<< "$SETS"
i=0
for SET in "${ARR[@]}"; do
for L in $SIZES; do
  env $SET TMK_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
     --log-file $O/${T}_sizes_${i}_$L.csv python scripts/prof_step.py ${WL:-configs} 3 $L > $O/${T}_sizes_${i}_$L.log 2>&1 || echo "size $L rc $?"
  python - $O/${T}_sizes_${i}_$L.csv "$SET" $L <<'PY'
import csv,re,sys
f,SET,L=sys.argv[1:4]
lines=[l for l in open(f) if not l.startswith('==')]
r=list(csv.reader(lines)); h=r[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
rows=[(re.sub(r'<.*|\(.*','',x[ki]).replace('void ','').replace('tmk::',''),float(x[vi].replace(',',''))/1000) for x in r[1:]]
idx=[i for i,(n,_) in enumerate(rows) if n=='k_wave_fk']
last=rows[idx[-1]:]
print(f'[{SET}] 2^{L}:',' '.join(f'{n[2:]}={v:.1f}' for n,v in last), 'sum=%.1f'%sum(v for _,v in last))
PY
done
i=$((i+1))
done
