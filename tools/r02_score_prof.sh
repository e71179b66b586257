#!/bin/bash
# Run under gpurun: the whole GPU suite, a score-mode bench line, one --set full capture of the
# k_score_fast launch of the third call of tools/score_bench.py (whole genome, 100k stripes).
TAG=${TAG:-r02z}
CFG=${CFG:-cfg3}
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
  python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
fi
python tools/score_bench.py --config $CFG --stripes 100000 --reps 5 --tag ${TAG} > gpurun_out/${TAG}_score_${CFG}.json 2> gpurun_out/${TAG}_score.err
cat gpurun_out/${TAG}_score_${CFG}.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_score_fast -s 2 -c 1 -f \
    -o gpurun_out/prof_${TAG}_k_score_fast python tools/score_bench.py --config $CFG --stripes 100000 --reps 1 \
    > gpurun_out/prof_${TAG}_k_score_fast.log 2>&1
ls -la gpurun_out | grep ${TAG}
