#!/bin/bash
# compute-sanitizer over the small-size GPU parity tests (memcheck, then racecheck + synccheck on the kernels that use
# shared memory).  Slow (10-100x): only the tests whose sizes are small.   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
set -u
mkdir -p gpurun_out
SMALL='test_field_arithmetic or test_point_arithmetic or test_bucket_reduction_kernels or (test_ntt_matches_oracle and not 16) or test_multiexp_g1_matches_oracle or test_multiexp_g2_matches_oracle or test_affine_rounds or test_multiexp_error_semantics or test_prove_mimc322 or test_fr_dot'
for tool in memcheck racecheck synccheck; do
    timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SMALL" > gpurun_out/sanitizer_$tool.log 2>&1
    echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Error" gpurun_out/sanitizer_$tool.log | tail -5
done
