#!/bin/bash
# strong scaling at N=8 and N=4: one genome (cfg3) sharded, 2 host threads per GPU
mkdir -p gpurun_out
for n in 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 12 --warmup 3 --no-score-mode > gpurun_out/r02u_bench_n$n.json 2> gpurun_out/r02u_bench_n$n.err
echo "bench n$n rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/r02u_bench_n$n.json').read().strip().splitlines()[-1])
print('N=$n ms/step %.1f value %.0f e2e %.0f (%.1f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))
print(d['host_phases_ms_last_step'])
print(d['config']['parallelism'][:700])
PY
tail -3 gpurun_out/r02u_bench_n$n.err
done
