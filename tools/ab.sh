#!/bin/bash
# A/B of two builds of the library on ONE box (box-to-box variance is ~5-8 %): BASE and NEW are
# paths of libstripenn_b200.so builds; runs the score micro-benchmark and the default bench
# alternately and prints the numbers side by side.
BASE=${BASE:-stripenn_b200/lib/libsnn_base.so}
NEW=${NEW:-stripenn_b200/lib/libstripenn_b200.so}
mkdir -p gpurun_out
for rep in ${SCORE_REPS:-1 2}; do
  for v in BASE NEW; do
    lib=${!v}
    STRIPENN_B200_LIB=$PWD/$lib python tools/score_bench.py --config cfg2 --stripes 100000 --reps 5 --tag $v 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('score', d['tag'], round(d['kernel_ms'],3), 'ms')"
  done
done
for rep in 1 2; do
  for v in BASE NEW; do
    lib=${!v}
    STRIPENN_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --no-score-mode --steps 4 --warmup 3 2>/dev/null > gpurun_out/ab_$v$rep.json
    python -c "
import json; d=json.load(open('gpurun_out/ab_$v$rep.json'))
print('bench $v', round(d['ms_per_step'],2), 'ms/step', {k: round(v['ms_per_step'],2) for k, v in d['roofline']['stages'].items() if k in ('gray','gaussian','sobel_nms','search','score')})"
  done
done
