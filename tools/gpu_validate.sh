#!/bin/bash
# Round validation on one B200: GPU tests, smoke, both bench arms, ncu launch list, ncu --set full of the two main kernels.
TAG=${1:-r2}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"
timeout 500 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ['gpurun_out/${TAG}_bench.json','gpurun_out/${TAG}_bench_reference.json']:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        if 'roofline' in d:
            r=d['roofline']
            print('value %.0f ms %.4f single %.4f e2e %.0f (%.3f ms) dropin %.0f frac %.3f whole %.3f launches %d cpu %.0f with_images %.0f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['dropin']['value'], r['frac'], r['whole_step']['frac'], d['gpu_launches'], d['cpu_baseline']['value'], d['cpu_baseline']['with_images']['value']), r['phase_ms'])
            print(d['e2e']['host_memory'], d['dropin']['ms_per_frame'])
        else:
            print('reference arm value %.1f cores %s' % (d['value'], d['cpu_baseline']['cores']))
    except Exception as e: print(f, 'ERR', e)
PY
B="python bench.py --steps 2 --warmup 3 --ramp-seconds 0 --no-cpu-baseline --lanes 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 16 --csv --log-file gpurun_out/${TAG}_launches.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'binned_raster|clip_geometry' -s 6 -c 2 -f -o gpurun_out/${TAG}_prof_final $B > gpurun_out/${TAG}_prof_final.log 2>&1
for w in config2 config2_cama config3; do timeout 200 python tools/quick_bench.py --workload $w --steps 30 --tag $w 2>&1 | tail -1; done > gpurun_out/${TAG}_workloads.jsonl
cat gpurun_out/${TAG}_workloads.jsonl
timeout 300 python tools/dropin_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_dropin.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt; nproc >> gpurun_out/${TAG}_smi.txt
