# Developer A/B: 2-GPU bench with the gradient all-reduce overlapped / at the end / disabled, static vs dynamic GEMM tiles.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { # name, env, mode
  env $2 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --grad-sync $3 2>&1 | grep '^{' > gpurun_out/sync_$1.json
  python -c "import json; d=json.load(open('gpurun_out/sync_$1.json')); print('$1', round(d['value'],2), round(d['ms_per_step'],1), round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"
}
run dyn_overlap NV_X=0 overlap
run static_overlap NV_GEMM_STATIC_SCHED=1 overlap
run dyn_none NV_X=0 none
run dyn_end NV_X=0 end
