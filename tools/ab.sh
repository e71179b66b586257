#!/bin/bash
# A/B of several builds of the library on ONE box (box-to-box variance is ~5-8 %).
#   LIBS="BASE=stripenn_b200/lib/libsnn_base.so NEW=stripenn_b200/lib/libstripenn_b200.so" bash tools/ab.sh
# runs the default bench alternately (REPS rounds) and prints ms/step and the image-chain stage
# times side by side; SCORE=1 also runs the score micro-benchmark.
LIBS=${LIBS:-"BASE=stripenn_b200/lib/libsnn_base.so NEW=stripenn_b200/lib/libstripenn_b200.so"}
REPS=${REPS:-2}
mkdir -p gpurun_out
if [ -n "$SCORE" ]; then
  for rep in $(seq $REPS); do
    for kv in $LIBS; do
      v=${kv%%=*}; lib=${kv#*=}
      STRIPENN_B200_LIB=$PWD/$lib python tools/score_bench.py --config cfg2 --stripes 100000 --reps 5 --tag $v 2>/dev/null \
        | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('score', d['tag'], round(d['kernel_ms'],3), 'ms')"
    done
  done
fi
for rep in $(seq $REPS); do
  for kv in $LIBS; do
    v=${kv%%=*}; lib=${kv#*=}
    STRIPENN_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --no-score-mode --steps 4 --warmup 3 2>/dev/null > gpurun_out/ab_$v$rep.json
    python -c "
import json; d=json.load(open('gpurun_out/ab_$v$rep.json'))
print('bench $v', round(d['ms_per_step'],2), 'ms/step', {k: round(v['ms_per_step'],2) for k, v in d['roofline']['stages'].items() if k in ('gray','gaussian','sobel_nms','search','score')})"
  done
done
