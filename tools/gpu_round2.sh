#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu2.log
tail -15 gpurun_out/r2/pytest_gpu2.log
timeout 120 tools/bin/mma_ts_check > gpurun_out/r2/mma_ts_check.txt 2>&1; echo "ts rc=$?"
timeout 300 python tools/c3_e2e_breakdown.py > gpurun_out/r2/c3_breakdown.txt 2>&1; echo "c3 rc=$?"; cat gpurun_out/r2/c3_breakdown.txt
