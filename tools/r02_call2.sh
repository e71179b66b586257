#!/bin/bash
# round-2 call 2: GPU tests after the host-side rewrite (typed exchanges, device-resident background)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_cfg3.json 2> gpurun_out/r02b_bench_cfg3.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02b_bench_cfg3.json; tail -5 gpurun_out/r02b_bench_cfg3.err
