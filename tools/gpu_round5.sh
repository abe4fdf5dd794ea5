#!/bin/bash
mkdir -p gpurun_out/r2
for pdl in 0 1; do NAVILLM_DECODE_PDL=$pdl timeout 600 python -m pytest tests/test_generate_gpu.py -m gpu -q -s -k "summarization or 3dqa_generation" > gpurun_out/r2/pytest_gen_pdl$pdl.log 2>&1; echo "pdl=$pdl rc=$?"; grep -E "passed|failed|AssertionError" gpurun_out/r2/pytest_gen_pdl$pdl.log | head -5; done
timeout 300 python tools/attn_bench.py > gpurun_out/r2/attn_bench_3.txt 2>&1; cat gpurun_out/r2/attn_bench_3.txt
NV_NVCC_EXTRA=-DNV_ATTN_TRACE timeout 600 python -m navillm_b200.build --force > gpurun_out/r2/build_trace.log 2>&1; echo "trace build rc=$?"
timeout 120 python tools/attn_fwd_trace.py > gpurun_out/r2/attn_fwd_trace_ragged.txt 2>&1; timeout 120 python tools/attn_fwd_trace.py dense > gpurun_out/r2/attn_fwd_trace_dense.txt 2>&1; echo "trace rc=$?"
