#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02j_pytest.log
for m in 0 a 1; do
SNN_FUSED=$m timeout 400 python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-score-mode > gpurun_out/r02j_bench_cfg3_$m.json 2> gpurun_out/r02j_bench_cfg3_$m.err
echo "bench cfg3 mode $m rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/r02j_bench_cfg3_$m.json').read().strip().splitlines()[-1])
print('ms/step %.1f e2e %.1f'%(d['ms_per_step'], d['e2e']['ms_per_step']), {k:round(v['ms_per_step'],1) for k,v in d['roofline']['stages'].items()})
PY
done
