#!/bin/bash
# usage: tools/gpu_bench_matrix.sh [--tests] "ENV=.. ENV2=.." "ENV=.." ...   (each argument = one bench run with that environment)
if [ "$1" == "--tests" ]; then shift; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
for E in "$@"; do
  echo "== $E"
  env $E timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline ${BENCH_ARGS} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); ph=d['roofline']['phase_ms']
        print('value %.0f  ms/step %.4f  geometry %.1f us  lists %.1f us  raster %.1f us  frac %.3f' % (d['value'], d['ms_per_step'], ph['geometry']*1e3, ph['lists']*1e3, ph['raster']*1e3, d['roofline']['frac']))
    else: print(l.rstrip()[-300:])
"
done
