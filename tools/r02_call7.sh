#!/bin/bash
mkdir -p gpurun_out
SNN_FUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_chain or stage_images or narrow or raw_candidates" > gpurun_out/r02g_pytest_1.log 2>&1
echo "pytest mode 1 rc=$?"; tail -3 gpurun_out/r02g_pytest_1.log
for m in a 1; do
SNN_FUSED=$m timeout 300 python bench.py --config cfg2 --steps 4 --warmup 3 --no-cpu-baseline --no-score-mode > gpurun_out/r02g_bench_cfg2_$m.json 2> gpurun_out/r02g_bench_cfg2_$m.err
echo "bench mode $m rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/r02g_bench_cfg2_$m.json').read().strip().splitlines()[-1])
print('ms/step %.1f'%d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['roofline']['stages'].items()})
PY
done
export SNN_FUSED=1
TAG=r02g KERNELS="k_gh_nms" LAUNCHES=0 KSKIP=0 bash tools/profile.sh > /dev/null 2>&1
ls gpurun_out | grep r02g
