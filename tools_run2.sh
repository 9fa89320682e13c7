for D in 0 1; do
  echo "== CAMA_RASTER_DEBUG=$D"
  CAMA_BAND_ROWS=16 CAMA_RASTER_DEBUG=$D timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['roofline']['frac'])
    else: print(l.rstrip()[-300:])
"
done
