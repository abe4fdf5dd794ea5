#!/bin/bash
# second 8-GPU sweep: overlapped-slice CTA count (exposed parts use 32), slice size, raw (non-overlapped) exchange time
N=${1:-8}
OUT=gpurun_out/r2/n${N}b
mkdir -p $OUT
run() { name=$1; shift; envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
    > $OUT/$name.json 2> $OUT/$name.err
  echo "$name rc=$? $(python -c "
import json,sys
try:
    j=json.loads(open('$OUT/$name.json').read().strip().splitlines()[-1]); print(round(j['value'],2), round(j['ms_per_step'],1), j['clocks']['sm_mhz'])
except Exception as e: print('no json', e)")"
}
run none "X=1" --steps 8 --warmup 3 --grad-sync none
run nvls8 "NAVILLM_NVLS_CTAS=8" --steps 8 --warmup 3
run nvls4 "NAVILLM_NVLS_CTAS=4" --steps 8 --warmup 3
run nvls2 "NAVILLM_NVLS_CTAS=2" --steps 8 --warmup 3
run nvls8_chunk4 "NAVILLM_NVLS_CTAS=8 NAVILLM_SYNC_CHUNK=4" --steps 8 --warmup 3
run nvls32_end "NAVILLM_NVLS_CTAS=32" --steps 8 --warmup 3 --grad-sync end
