#!/bin/bash
O=gpurun_out/r2/r17; mkdir -p $O
NAVILLM_WGRAD_STREAM=1 timeout 900 python -m pytest tests/test_fullwidth_parity_gpu.py tests/test_navmodel_gpu.py -m gpu -q -k "not generate" > $O/pytest_ws.log 2>&1; echo "tests(ws=1) rc=$?"; tail -2 $O/pytest_ws.log
for ws in 0 1 0 1; do
NAVILLM_WGRAD_STREAM=$ws timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c2_ws$ws.json 2> $O/c2_ws$ws.err
python -c "
import json;j=json.load(open('$O/c2_ws$ws.json'));print('ws=$ws',round(j['value'],2),round(j['ms_per_step'],1),round(j['e2e']['value'],2),j['clocks']['sm_mhz'])"
done
