#!/bin/bash
# strong scaling at N=2: one genome (cfg3) sharded over two GPUs, 2 host threads each
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 --no-score-mode > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err
echo "bench n2 rc=$?"; tail -c 2500 gpurun_out/r02q_bench_n2.json; tail -5 gpurun_out/r02q_bench_n2.err
