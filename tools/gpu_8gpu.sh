#!/bin/bash
# 8-GPU evidence run (one box, N = 8): C2 weak scaling with the gradient-exchange variants, C5 (B = 32 = 8 x 4), NCCL algo log
mkdir -p gpurun_out/r2/n8
N=${1:-8}
run() { # name, extra env, args
  name=$1; shift; envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
    > gpurun_out/r2/n8/$name.json 2> gpurun_out/r2/n8/$name.err
  echo "$name rc=$? $(python -c "
import json,sys
try:
    j=json.loads(open('gpurun_out/r2/n8/$name.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'])
except Exception as e: print('no json', e)")"
}
run c2_overlap "NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL" --steps 10 --warmup 3
grep -E "NVLS|Channel|Algo|algo|Using network|nChannels|comm .* rank 0" gpurun_out/r2/n8/c2_overlap.err | head -30 > gpurun_out/r2/n8/nccl_info.txt
run c2_end "X=1" --steps 10 --warmup 3 --grad-sync end
run c2_none "X=1" --steps 10 --warmup 3 --grad-sync none
run c2_overlap_cta8 "NCCL_MAX_CTAS=8" --steps 10 --warmup 3
run c5_overlap "X=1" --workload c5 --steps 6 --warmup 3
nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit --format=csv > gpurun_out/r2/n8/smi_after.txt 2>&1
