#!/bin/bash
# What the driver runs at round end, on one box: GPU tests, smoke(), the default bench line, the reference arm.
O=${1:-gpurun_out/final_check}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python -c "
import json
r=json.load(open('$O/bench_reference.json')); j=json.load(open('$O/bench.json'))
print('reference', r['value'], r['cpu_baseline']['cores']); print('bench', round(j['value'],2), round(j['ms_per_step'],1), 'e2e', round(j['e2e']['value'],2), 'roofline', round(j['roofline']['frac'],3), j['clocks'], 'cpu', j['cpu_baseline']['value'], 'e2e ratio', j['e2e']['value']/r['value'])"
