#!/bin/bash
O=gpurun_out/r2/r18; mkdir -p $O
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_generate_gpu.py tests/test_fullwidth_parity_gpu.py -m gpu -q -k "decode or generate" > $O/pytest.log 2>&1; echo "tests rc=$?"; tail -3 $O/pytest.log
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/c3.json 2> $O/c3.err; python -c "
import json;j=json.load(open('$O/c3.json'));print('c3',round(j['value'],1),j['config']['ms_per_token'],j['roofline']['frac'],j['e2e']['value'])"
