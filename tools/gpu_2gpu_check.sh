#!/bin/bash
# 2-GPU check of the data-parallel wrapper: NVLS unit test + DDP semantics over NCCL and NVLS, then a short C2 weak-scaling step
O=gpurun_out/r2/n2c; mkdir -p $O
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s > $O/pytest_ddp.log 2>&1; echo "ddp test rc=$?"; grep -E "nvls|ddp overlap|passed|failed|Error|warn" $O/pytest_ddp.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 > $O/c2.json 2> $O/c2.err; echo "c2 rc=$?"
python -c "
import json
j=json.loads(open('$O/c2.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['config'].get('grad_exchange'), j['clocks'])"
