#!/bin/bash
# round 2, run 5 (one B200): lanes x geometry occupancy x unit frames on the 2-lane bench
TAG=${1:-r2f}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for E in "X=1" "CAMA_GEO_STATIC_PCT=40" "CAMA_GEO_UNIT_FRAMES=2"; do
  env $E timeout 200 python tools/quick_bench.py --workload config2 --steps 40 --tag "config2 $E" 2>&1 | tail -1
done
for E in "X=1" "CAMA_GEO_CTAS=3" "CAMA_GEO_CTAS=2" "CAMA_RASTER_CTAS=3"; do
 for L in 2 3 4; do
  echo "== bench lanes $L $E"
  env $E timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --lanes $L 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); ph=d['roofline']['phase_ms']
        print('value %.0f  ms/step %.4f single %.4f geometry %.1f us  lists %.1f us  raster %.1f us  frac %.3f whole %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], ph['geometry']*1e3, ph['lists']*1e3, ph['raster']*1e3, d['roofline']['frac'], d['roofline']['whole_step']['frac']))
"
 done
done
