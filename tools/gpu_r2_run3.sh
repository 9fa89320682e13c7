#!/bin/bash
# round 2, run 3 (one B200): runtime knobs of the geometry kernel and the frame-group pipeline, with the 2-lane bench
TAG=${1:-r2c}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
{
for E in "X=1" "CAMA_GEO_STATIC_PCT=0" "CAMA_GEO_STATIC_PCT=50" "CAMA_GEO_STATIC_PCT=90" "CAMA_GEO_STATIC_PCT=100" "CAMA_GEO_UNIT_FRAMES=8" "CAMA_PIPE_FRAMES=24" "CAMA_PIPE_FRAMES=16" "CAMA_GEO_CTAS=3" "CAMA_GEO_CTAS=2"; do
  env $E timeout 200 python tools/quick_bench.py --workload config2 --steps 40 --tag "config2 $E" 2>&1 | tail -1
done
for E in "X=1" "CAMA_GEO_UNIT_FRAMES_SITE=4" "CAMA_PIPE_FRAMES=40" "CAMA_PIPE_FRAMES=160"; do
  env $E timeout 200 python tools/quick_bench.py --workload config3 --steps 20 --tag "config3 $E" 2>&1 | tail -1
done
} > gpurun_out/${TAG}_workloads.jsonl
cat gpurun_out/${TAG}_workloads.jsonl
for E in "X=1" "CAMA_GEO_CTAS=3" "CAMA_GEO_CTAS=2"; do
 for L in 2 3; do
  echo "== bench lanes $L $E"
  env $E timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --lanes $L 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); ph=d['roofline']['phase_ms']
        print('value %.0f  ms/step %.4f single %.4f geometry %.1f us  lists %.1f us  raster %.1f us  frac %.3f whole %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], ph['geometry']*1e3, ph['lists']*1e3, ph['raster']*1e3, d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['e2e']['value']))
"
 done
done
timeout 300 python tools/dropin_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_dropin.json
