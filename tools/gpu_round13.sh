#!/bin/bash
O=gpurun_out/r2/final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 rc=$?"
# launch list of one C2 step and of the C3 decode (per-launch times: cold cache, serialised -> shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/launches_c2.csv python bench.py --steps 1 --warmup 1 --profile > $O/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --profile > $O/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
# full captures: attention kernels (c2_ragged shape of tools/attn_bench.py), then the decode kernels
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"attn_fwd_kernel|attn_bwd_dq_kernel|attn_bwd_dkv_kernel" -c 3 -o $O/prof_attn python tools/attn_bench.py > $O/prof_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"gemm_skinny_tcgen05|decode_attn_kernel" -s 1200 -c 6 -o $O/prof_decode python bench.py --workload c3 --steps 1 --warmup 1 --profile > $O/prof_decode.log 2>&1; echo "ncu decode rc=$?"
for f in c2 c5 c3; do python -c "
import json;j=json.load(open('$O/bench_$f.json'));print('$f',round(j['value'],2),round(j['ms_per_step'],1),round(j['e2e']['value'],2),j['roofline']['frac'],j['clocks']['sm_mhz'],j.get('cpu_baseline',{}).get('value'))"; done
