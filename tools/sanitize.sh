#!/bin/bash
# compute-sanitizer pass over the small parity cases (run under gpurun): memcheck on the whole
# compute + score path, racecheck on the shared-memory kernels.  Logs land in gpurun_out/.
mkdir -p gpurun_out
K=${K:-"compute_matches_reference_tiny2 or score_matches_reference"}
timeout ${T:-900} compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_memcheck.log \
    python -m pytest tests/test_gpu_parity.py tests/test_next_rows.py -x -q -m gpu -k "$K or mask_matches or slow_quantiles" \
    > gpurun_out/sanitize_memcheck.out 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/sanitize_memcheck.log; tail -2 gpurun_out/sanitize_memcheck.out
timeout ${T:-900} compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitize_racecheck.log \
    python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "score_matches_reference or small_and_empty" \
    > gpurun_out/sanitize_racecheck.out 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/sanitize_racecheck.log; tail -2 gpurun_out/sanitize_racecheck.out
