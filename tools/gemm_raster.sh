# Developer experiment: DRAM traffic and time of the big-K GEMMs for different rasterisation group sizes.
for g in 2 4 8 16; do
  echo "== GROUP_M=$g"
  NV_GEMM_GROUP_M=$g python tools/gemm_bench.py --no-cublas --only-512 --tokens 10400 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['name'], round(d['nv_bn512_ms'],3), 'ms', round(d['nv_bn512_tflops']))
"
done
for g in 4 8 16; do
  NV_GEMM_GROUP_M=$g ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:2cta --csv --log-file gpurun_out/raster_g$g.csv python tools/gemm_bench.py --no-cublas --only-512 --iters 1 --warmup 1 --tokens 10400 > /dev/null 2>&1
done
