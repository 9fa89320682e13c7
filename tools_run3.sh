timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools_run2.sh
for R in 0 20 12; do echo "== R=$R"; CAMA_BAND_ROWS=$R timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['roofline']['frac'])
    else: print(l.rstrip()[-300:])
"; done
