#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_properties_gpu.py tests/test_llama_gpu.py tests/test_prefix_reuse_gpu.py -m gpu -q -x > gpurun_out/r2/pytest_attn.log 2>&1; echo "attn tests rc=$?"; tail -5 gpurun_out/r2/pytest_attn.log
timeout 300 python tools/attn_bench.py > gpurun_out/r2/attn_bench_fwd2.txt 2>&1; cat gpurun_out/r2/attn_bench_fwd2.txt
timeout 600 python -m pytest tests/test_fullwidth_parity_gpu.py -m gpu -q -x -s -k "generate or stack" > gpurun_out/r2/pytest_fw.log 2>&1; echo "fullwidth rc=$?"; grep -E "c3 generate|passed|failed" gpurun_out/r2/pytest_fw.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2/launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --profile > gpurun_out/r2/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
