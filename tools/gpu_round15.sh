#!/bin/bash
O=gpurun_out/r2/dec; mkdir -p $O
for cfg in "8 8" "4 8" "2 8" "8 16" "8 4"; do set -- $cfg
NV_SKINNY_MAX_SPLITS=$1 NV_SKINNY_MIN_KB=$2 timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/c3_s$1_k$2.json 2> $O/c3_s$1_k$2.err
python -c "
import json;j=json.load(open('$O/c3_s$1_k$2.json'));print('splits<=$1 min_kb=$2',round(j['value'],1),j['config']['ms_per_token'],j['roofline']['frac'])"
done
