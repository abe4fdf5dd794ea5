#!/bin/bash
mkdir -p gpurun_out/r2/n2
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s > gpurun_out/r2/n2/pytest_ddp_b.log 2>&1; echo "ddp test rc=$?"; grep -E "nvls|ddp overlap|passed|failed|Error" gpurun_out/r2/n2/pytest_ddp_b.log | head -20
for nvls in 1 0; do
NAVILLM_NVLS=$nvls timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2/n2/c2_nvls$nvls.json 2> gpurun_out/r2/n2/c2_nvls$nvls.err; echo "nvls=$nvls rc=$?"
python -c "
import json
j=json.loads(open('gpurun_out/r2/n2/c2_nvls$nvls.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['config']['grad_exchange'][-90:])"
grep -i "warn\|nvls" gpurun_out/r2/n2/c2_nvls$nvls.err | head -5
done
