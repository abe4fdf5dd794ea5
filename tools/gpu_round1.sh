#!/bin/bash
# GPU round: tests + the three bench workloads (1 GPU)
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu.log
tail -5 gpurun_out/r2/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_c2.json 2> gpurun_out/r2/bench_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r2/bench_c5.json 2> gpurun_out/r2/bench_c5.err; echo "c5 rc=$?"
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/r2/bench_c3.json 2> gpurun_out/r2/bench_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_ref.json 2> gpurun_out/r2/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r2/bench_c2.json | head -c 1500
