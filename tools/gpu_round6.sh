#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_properties_gpu.py tests/test_llama_gpu.py tests/test_prefix_reuse_gpu.py tests/test_generate_gpu.py -m gpu -q > gpurun_out/r2/pytest_attn6.log 2>&1; echo "attn+generate tests rc=$?"; tail -4 gpurun_out/r2/pytest_attn6.log
timeout 300 python tools/attn_bench.py > gpurun_out/r2/attn_bench_6.txt 2>&1; cat gpurun_out/r2/attn_bench_6.txt
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_c3_6.json 2> gpurun_out/r2/bench_c3_6.err; python -c "
import json;j=json.load(open('gpurun_out/r2/bench_c3_6.json'));print('c3',j['value'],j['config']['ms_per_token'],j['config']['prefill_ms'],j['e2e']['value'],j['roofline']['frac'])"
NV_NVCC_EXTRA=-DNV_ATTN_TRACE timeout 600 python -m navillm_b200.build --force > gpurun_out/r2/build_trace.log 2>&1; echo "trace build rc=$?"
timeout 120 python tools/attn_fwd_trace.py > gpurun_out/r2/attn_fwd_trace_ragged6.txt 2>&1; timeout 120 python tools/attn_fwd_trace.py dense > gpurun_out/r2/attn_fwd_trace_dense6.txt 2>&1; echo "trace rc=$?"
timeout 120 python tools/attn_trace.py 1 > gpurun_out/r2/attn_bwd_trace_dq6.txt 2>&1; timeout 120 python tools/attn_trace.py 0 > gpurun_out/r2/attn_bwd_trace_dkv6.txt 2>&1
