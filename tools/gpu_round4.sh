#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 ncu --set full --import-source on --clock-control none -k regex:attn_fwd_kernel -c 2 -o gpurun_out/r2/prof_attn_fwd_v2 python tools/attn_bench.py > gpurun_out/r2/prof_attn_fwd_v2.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests/test_generate_gpu.py tests/test_fullwidth_parity_gpu.py -m gpu -q -x -s -k "generate" > gpurun_out/r2/pytest_gen.log 2>&1; echo "generate tests rc=$?"; grep -E "generate,|c3 generate|passed|failed|Error" gpurun_out/r2/pytest_gen.log | head -20
for pdl in 1 0; do NAVILLM_DECODE_PDL=$pdl timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_c3_pdl$pdl.json 2> gpurun_out/r2/bench_c3_pdl$pdl.err; echo "c3 pdl=$pdl rc=$?"; python -c "
import json;j=json.load(open('gpurun_out/r2/bench_c3_pdl$pdl.json'));print(j['value'],j['config']['ms_per_token'],j['config']['prefill_ms'],j['e2e']['value'],j['roofline']['frac'])"; done
timeout 600 python tools/ref_eager_b200.py > gpurun_out/r2/ref_eager.json 2> gpurun_out/r2/ref_eager.err; echo "ref eager rc=$?"; cat gpurun_out/r2/ref_eager.json | head -40
