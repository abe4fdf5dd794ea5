#!/bin/bash
# compute-sanitizer memcheck over the suites of the kernels added late in round 2 (sampling, fuse_obj maps, per-layer call)
OUT=${1:-gpurun_out/sanitizer_new}; mkdir -p $OUT
run() { n=$1; shift
  timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/$n.log python -m pytest "$@" -m gpu -q -x > $OUT/$n.pytest.log 2>&1
  echo "$n rc=$? $(grep -E 'ERROR SUMMARY' $OUT/$n.log | tail -1) | $(tail -1 $OUT/$n.pytest.log)"; }
run test_sampling_gpu tests/test_sampling_gpu.py
run test_pano_fuse_obj tests/test_pano_gpu.py -k "fuse_obj"
run test_layer_call tests/test_llama_gpu.py -k "layer_call"
run test_generate_b20 tests/test_generate_gpu.py -k "batch_over_16"
