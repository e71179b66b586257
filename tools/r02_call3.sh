#!/bin/bash
# round-2 call 3: first run of the fused image chain
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_chain or stage_images or narrow or raw_candidates" > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/r02c_pytest.log
timeout 300 python bench.py --config cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-score-mode > gpurun_out/r02c_bench_cfg2.json 2> gpurun_out/r02c_bench_cfg2.err
echo "bench rc=$?"; tail -c 1200 gpurun_out/r02c_bench_cfg2.json; tail -5 gpurun_out/r02c_bench_cfg2.err
