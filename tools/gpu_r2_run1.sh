#!/bin/bash
# round 2, run 1 (one B200): GPU tests with the new geometry kernel, A/B of its parts, the site block a rank of 8 renders
TAG=${1:-r2a}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
{
for w in config2 config3; do
  timeout 200 python tools/quick_bench.py --workload $w --steps 40 --tag ${w}_new 2>&1 | tail -1
  CAMA_GEO_NO_CAMTABLE=1 timeout 200 python tools/quick_bench.py --workload $w --steps 40 --tag ${w}_no_camtable 2>&1 | tail -1
  CAMA_GEO_NO_WARPBOUNDS=1 timeout 200 python tools/quick_bench.py --workload $w --steps 40 --tag ${w}_no_warpbounds 2>&1 | tail -1
done
timeout 200 python tools/quick_bench.py --workload config3 --frames 0:40 --steps 40 --tag config3_block40 2>&1 | tail -1
timeout 200 python tools/quick_bench.py --workload config3 --frames 0:40 --sparse --steps 40 --tag config3_block40_sparse 2>&1 | tail -1
timeout 200 python tools/quick_bench.py --workload config2_cama --steps 20 --tag config2_cama 2>&1 | tail -1
} > gpurun_out/${TAG}_workloads.jsonl
cat gpurun_out/${TAG}_workloads.jsonl
timeout 500 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --steps 50 --warmup 5 --lanes 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_lanes3.json 2>/dev/null
B="python bench.py --steps 2 --warmup 3 --ramp-seconds 0 --no-cpu-baseline --lanes 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'clip_geometry' -s 3 -c 1 -f -o gpurun_out/${TAG}_geo $B > gpurun_out/${TAG}_geo.log 2>&1
python -c "
import json
for f in ['gpurun_out/${TAG}_bench.json','gpurun_out/${TAG}_bench_lanes3.json']:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, 'value %.0f ms %.4f single %.4f e2e %.0f frac %.3f whole %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['e2e']['value'], r['frac'], r['whole_step']['frac']), r['phase_ms'])
    except Exception as e: print(f, 'ERR', e)
"
