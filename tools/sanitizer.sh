#!/bin/bash
# compute-sanitizer memcheck over the kernel test suites (bounded; the tests run 10-50x slower under the tool)
OUT=${1:-gpurun_out/sanitizer}; mkdir -p $OUT
for t in tests/test_attn_gpu.py tests/test_gemm_gpu.py tests/test_lm_ops_gpu.py tests/test_generate_gpu.py; do
  n=$(basename $t .py)
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/$n.log python -m pytest $t -m gpu -q -x > $OUT/$n.pytest.log 2>&1
  echo "$n rc=$? $(grep -E 'ERROR SUMMARY' $OUT/$n.log | tail -1) | $(tail -1 $OUT/$n.pytest.log)"
done
