#!/bin/bash
# Run under gpurun: closing pass after the k_gh_nms ring — whole GPU suite, the default bench line,
# one --set full capture of k_gh_nms on a full 51-window launch (cfg1).
TAG=${TAG:-r02zd}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['kernel'], round(d['roofline']['frac'],4), {k: round(v['ms_per_step'],1) for k,v in d['roofline']['stages'].items() if v['ms_per_step']>1})
print(d['score_mode']['kernel_ms'], d['score_mode']['roofline']['frac'], d['clocks'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_gh_nms -s 2 -c 1 -f \
    -o gpurun_out/prof_${TAG}_k_gh_nms python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline \
    --no-score-mode --config cfg1 > gpurun_out/prof_${TAG}_k_gh_nms.log 2>&1
ls gpurun_out | grep ${TAG}
