#!/bin/bash
# Run under gpurun: closing pass of a session — whole GPU suite, score-mode lines (fast / warp kernels),
# one --set full capture of k_score_fast on the whole-genome table, the default bench line.
TAG=${TAG:-r02zb}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
python tools/score_bench.py --config cfg3 --stripes 100000 --reps 5 --tag ${TAG}_fast > gpurun_out/${TAG}_score_fast.json 2> gpurun_out/${TAG}_score.err
SNN_SCORE_FAST=0 python tools/score_bench.py --config cfg3 --stripes 100000 --reps 5 --tag ${TAG}_warp > gpurun_out/${TAG}_score_warp.json 2>> gpurun_out/${TAG}_score.err
python tools/score_bench.py --config cfg2 --stripes 100000 --reps 5 --tag ${TAG}_fast_cfg2 > gpurun_out/${TAG}_score_fast_cfg2.json 2>> gpurun_out/${TAG}_score.err
cut -c1-260 gpurun_out/${TAG}_score_fast.json gpurun_out/${TAG}_score_warp.json gpurun_out/${TAG}_score_fast_cfg2.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_score_fast -s 2 -c 1 -f \
    -o gpurun_out/prof_${TAG}_k_score_fast python tools/score_bench.py --config cfg3 --stripes 100000 --reps 1 \
    > gpurun_out/prof_${TAG}_k_score_fast.log 2>&1
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['score_mode']['kernel_ms'], d['score_mode']['roofline']['frac'], d['score_mode']['stripes_routed'])
"
ls gpurun_out | grep ${TAG}
