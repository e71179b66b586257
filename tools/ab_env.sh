#!/bin/bash
# A/B of an environment switch of the library on ONE box:  VAR=SNN_GAUSS_HD bash tools/ab_env.sh
# runs the default bench alternately with VAR unset and VAR=1 (REPS rounds).
VAR=${VAR:?name of the switch}
REPS=${REPS:-2}
mkdir -p gpurun_out
for rep in $(seq $REPS); do
  for v in 0 1; do
    if [ $v = 1 ]; then export $VAR=1; else unset $VAR; fi
    python bench.py --no-cpu-baseline --no-score-mode --steps 4 --warmup 3 2>/dev/null > gpurun_out/abenv_${VAR}_$v$rep.json
    python -c "
import json; d=json.load(open('gpurun_out/abenv_${VAR}_$v$rep.json'))
print('bench $VAR=$v', round(d['ms_per_step'],2), 'ms/step', {k: round(x['ms_per_step'],2) for k, x in d['roofline']['stages'].items() if k in ('gray','gaussian','sobel_nms','search')})"
  done
done
unset $VAR
