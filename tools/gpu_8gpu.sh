#!/bin/bash
# N-GPU evidence run (one box): C2 weak scaling with the gradient-exchange variants, C5 (B = 4 per rank), NCCL algo log
N=${1:-8}
OUT=gpurun_out/r2/n$N
mkdir -p $OUT
run() { # name, "ENV=.. ENV=..", bench args
  name=$1; shift; envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
    > $OUT/$name.json 2> $OUT/$name.err
  echo "$name rc=$? $(python -c "
import json,sys
try:
    j=json.loads(open('$OUT/$name.json').read().strip().splitlines()[-1]); print(round(j['value'],2), round(j['ms_per_step'],1), round(j['e2e']['value'],2), j['clocks']['sm_mhz'], j['clocks']['reasons'], j['config'].get('grad_exchange','')[-70:])
except Exception as e: print('no json', e)")"
}
run c2_none "X=1" --steps 8 --warmup 3 --grad-sync none
run c2_nccl "NAVILLM_NVLS=0 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT NCCL_DEBUG_FILE=$OUT/nccl_debug_%p.log" --steps 8 --warmup 3
cat $OUT/nccl_debug_*.log 2>/dev/null | grep -E "NVLS|nvls|Algo|nChannels|Connected all|channels" | sed "s/^.*NCCL INFO //" | sort | uniq -c | sort -rn | head -20 > $OUT/nccl_info.txt; rm -f $OUT/nccl_debug_*.log
run c2_nvls32 "NAVILLM_NVLS_CTAS=32" --steps 8 --warmup 3
run c2_nvls8 "NAVILLM_NVLS_CTAS=8" --steps 8 --warmup 3
run c5_nvls "X=1" --workload c5 --steps 6 --warmup 3
nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit --format=csv > $OUT/smi_after.txt 2>&1
