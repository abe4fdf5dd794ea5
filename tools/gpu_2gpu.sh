#!/bin/bash
mkdir -p gpurun_out/r2/n2
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s > gpurun_out/r2/n2/pytest_ddp.log 2>&1; echo "ddp test rc=$?"; tail -5 gpurun_out/r2/n2/pytest_ddp.log
for mode in overlap none; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --grad-sync $mode > gpurun_out/r2/n2/c2_$mode.json 2> gpurun_out/r2/n2/c2_$mode.err; echo "$mode rc=$?"
python -c "
import json
j=json.loads(open('gpurun_out/r2/n2/c2_$mode.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['config']['grad_exchange'][-60:])"
done
