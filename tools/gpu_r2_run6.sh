#!/bin/bash
TAG=${1:-r2g}
for E in "CAMA_GEO_CTAS=4 CAMA_RASTER_CTAS=3" "CAMA_GEO_CTAS=3 CAMA_RASTER_CTAS=3" "CAMA_GEO_CTAS=2 CAMA_RASTER_CTAS=3" "CAMA_GEO_CTAS=3 CAMA_RASTER_CTAS=2" "CAMA_GEO_CTAS=2 CAMA_RASTER_CTAS=2"; do
 for L in 3 4; do
  echo "== bench lanes $L $E"
  env $E timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --lanes $L 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); ph=d['roofline']['phase_ms']
        print('value %.0f  ms/step %.4f single %.4f geometry %.1f us  raster %.1f us  frac %.3f whole %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], ph['geometry']*1e3, ph['raster']*1e3, d['roofline']['frac'], d['roofline']['whole_step']['frac']))
"
 done
done
