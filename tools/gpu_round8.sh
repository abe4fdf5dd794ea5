#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_properties_gpu.py tests/test_llama_gpu.py tests/test_fullwidth_parity_gpu.py -m gpu -q -k "not generate" > gpurun_out/r2/pytest_attn8.log 2>&1; echo "attn tests rc=$?"; tail -4 gpurun_out/r2/pytest_attn8.log
timeout 300 python tools/attn_bench.py > gpurun_out/r2/attn_bench_8.txt 2>&1; cat gpurun_out/r2/attn_bench_8.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_c2_8.json 2> gpurun_out/r2/bench_c2_8.err; python -c "
import json;j=json.load(open('gpurun_out/r2/bench_c2_8.json'));print('c2',j['value'],j['ms_per_step'],j['e2e']['value'],j['roofline']['frac'],j['roofline']['share_of_step'],j['clocks'])"
NV_NVCC_EXTRA=-DNV_ATTN_TRACE timeout 600 python -m navillm_b200.build --force > gpurun_out/r2/build_trace.log 2>&1; echo "trace build rc=$?"
timeout 120 python tools/attn_trace.py 1 > gpurun_out/r2/attn_bwd_trace_dq8.txt 2>&1; timeout 120 python tools/attn_trace.py 0 > gpurun_out/r2/attn_bwd_trace_dkv8.txt 2>&1
