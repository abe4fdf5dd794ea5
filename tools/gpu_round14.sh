#!/bin/bash
O=gpurun_out/r2/final; mkdir -p $O
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"; tail -3 $O/bench_c4.err
timeout 600 python bench.py --workload c1 --steps 10 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err; echo "c1 rc=$?"; tail -3 $O/bench_c1.err
for f in c4 c1; do python -c "
import json;j=json.load(open('$O/bench_$f.json'));print('$f',round(j['value'],2),round(j['ms_per_step'],1),round(j['e2e']['value'],2),j['roofline']['frac'],j['clocks']['sm_mhz'],j.get('cpu_baseline',{}).get('value'))"; done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"attn_bwd_dq_kernel|attn_bwd_dkv_kernel" -c 2 -o $O/prof_attn_bwd python tools/attn_bench.py > $O/prof_attn_bwd.log 2>&1; echo "ncu attn bwd rc=$?"
bash tools/sanitizer.sh gpurun_out/r2/sanitizer
