#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
( time python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err ) 2>&1 | tail -3
echo "native rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench_n1.json').read().strip().splitlines()[-1])
print('ms/step %.1f value %.0f e2e %.0f'%(d['ms_per_step'],d['value'],d['e2e']['value']))
r=d['roofline']; print({k:r[k] for k in r if k not in ('stages','hbm','fp64_pipe_rates_per_clk_per_sm')}); print(r['hbm']); print(d['cpu_baseline']); print(d['score_mode']['kernel_ms'], d['score_mode']['roofline']['frac'])
PY
tail -3 gpurun_out/r02m_bench_n1.err
( time python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err ) 2>&1 | tail -3
echo "ref rc=$?"; cut -c1-1200 gpurun_out/r02m_bench_ref.json
