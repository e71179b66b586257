#!/bin/bash
mkdir -p gpurun_out
tools/bin/check_sqrt | tee gpurun_out/r02e_check_sqrt.json
SNN_FUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_chain or stage_images or narrow or raw_candidates" > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02e_pytest.log
SNN_FUSED=1 timeout 300 python bench.py --config cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-score-mode > gpurun_out/r02e_bench_cfg2.json 2> gpurun_out/r02e_bench_cfg2.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench_cfg2.json').read().strip().splitlines()[-1])
print('ms/step %.1f'%d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})
PY
export SNN_FUSED=1
TAG=r02e KERNELS="k_gray_gv k_gh_nms" LAUNCHES=0 KSKIP=0 bash tools/profile.sh > /dev/null 2>&1
ls gpurun_out | grep r02e
