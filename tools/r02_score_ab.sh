#!/bin/bash
# Run under gpurun: score-mode bench lines for the default library and every variant under
# stripenn_b200/lib/variants (same box, interleaved ROUNDS times); TESTS=1 runs the fast-scoring tests first.
TAG=${TAG:-r02w}
CFG=${CFG:-cfg3}
mkdir -p gpurun_out
if [ "${TESTS:-0}" = "1" ]; then
  python -m pytest tests/test_gpu_score_fast.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
fi
for round in $(seq 1 ${ROUNDS:-2}); do
  python tools/score_bench.py --config $CFG --stripes 100000 --reps 5 --tag ${TAG}_default_$round 2>>gpurun_out/${TAG}_score.err | tee -a gpurun_out/${TAG}_score_ab.jsonl | cut -c1-200
  for v in stripenn_b200/lib/variants/*.so; do
    [ -f "$v" ] || continue
    STRIPENN_B200_LIB=$v python tools/score_bench.py --config $CFG --stripes 100000 --reps 5 --tag ${TAG}_$(basename $v .so)_$round 2>>gpurun_out/${TAG}_score.err | tee -a gpurun_out/${TAG}_score_ab.jsonl | cut -c1-200
  done
done
