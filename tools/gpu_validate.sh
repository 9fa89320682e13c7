#!/bin/bash
# Round validation on one B200: GPU tests, smoke, both bench arms, ncu launch list, ncu --set full of the two main kernels.
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"
timeout 500 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err
B="python bench.py --steps 2 --warmup 3 --ramp-seconds 0 --no-cpu-baseline --lanes 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 20 --csv --log-file gpurun_out/${TAG}_launches.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'binned_raster|clip_geometry' -s 6 -c 2 -f -o gpurun_out/${TAG}_prof_final $B > gpurun_out/${TAG}_prof_final.log 2>&1
timeout 100 python tools/raster_timeline.py > gpurun_out/${TAG}_raster_timeline.txt 2>&1
timeout 100 python tools/e2e_breakdown.py > gpurun_out/${TAG}_e2e_breakdown.txt 2>&1
for w in config2 config2_cama config3; do timeout 200 python tools/quick_bench.py --workload $w --steps 30 --tag $w 2>&1 | tail -1; done > gpurun_out/${TAG}_workloads.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt; nproc >> gpurun_out/${TAG}_smi.txt
ls -la gpurun_out | tail -14
