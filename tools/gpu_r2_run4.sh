#!/bin/bash
# round 2, run 4 (one B200): the sort-free pipeline (direct record lists per band group)
TAG=${1:-r2d}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
{
for E in "X=1" "CAMA_GEO_STAGE=1"; do
  env $E timeout 200 python tools/quick_bench.py --workload config2 --steps 40 --tag "config2 $E" 2>&1 | tail -1
  env $E timeout 200 python tools/quick_bench.py --workload config3 --steps 20 --tag "config3 $E" 2>&1 | tail -1
done
timeout 200 python tools/quick_bench.py --workload config3 --frames 0:40 --steps 40 --tag config3_block40 2>&1 | tail -1
timeout 200 python tools/quick_bench.py --workload config3 --frames 0:40 --sparse --steps 40 --tag config3_block40_sparse 2>&1 | tail -1
timeout 200 python tools/quick_bench.py --workload config2_cama --steps 20 --tag config2_cama 2>&1 | tail -1
} > gpurun_out/${TAG}_workloads.jsonl
cat gpurun_out/${TAG}_workloads.jsonl
for L in 1 2 3; do
  echo "== bench lanes $L"
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --lanes $L 2>gpurun_out/${TAG}_bench_l$L.err | tee gpurun_out/${TAG}_bench_l$L.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); ph=d['roofline']['phase_ms']
        print('value %.0f  ms/step %.4f single %.4f geometry %.1f us  lists %.1f us  raster %.1f us  frac %.3f whole %.3f e2e %.0f dropin %.0f launches %d' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], ph['geometry']*1e3, ph['lists']*1e3, ph['raster']*1e3, d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['e2e']['value'], d['dropin']['value'], d['gpu_launches']))
"; tail -2 gpurun_out/${TAG}_bench_l$L.err
done
B="python bench.py --steps 2 --warmup 3 --ramp-seconds 0 --no-cpu-baseline --lanes 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'clip_geometry|binned_raster' -s 6 -c 2 -f -o gpurun_out/${TAG}_prof $B > gpurun_out/${TAG}_prof.log 2>&1
