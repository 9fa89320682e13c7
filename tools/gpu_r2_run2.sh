#!/bin/bash
# round 2, run 2 (one B200): geometry kernel variants (compile-time), A/B on config 2 / config 3 / a 40-frame site block
TAG=${1:-r2b}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
{
for v in default uf8 uf2 minb5 minb3 flush64; do
  if [ $v == default ]; then L=""; else L="CAMA_B200_LIB=build_variants/lib_$v.so"; fi
  env $L timeout 200 python tools/quick_bench.py --workload config2 --steps 40 --tag config2_$v 2>&1 | tail -1
  env $L timeout 200 python tools/quick_bench.py --workload config3 --steps 20 --tag config3_$v 2>&1 | tail -1
done
timeout 200 python tools/quick_bench.py --workload config3 --frames 0:40 --steps 40 --tag config3_block40 2>&1 | tail -1
timeout 200 python tools/quick_bench.py --workload config2_cama --steps 20 --tag config2_cama 2>&1 | tail -1
} > gpurun_out/${TAG}_workloads.jsonl
cat gpurun_out/${TAG}_workloads.jsonl
B="python bench.py --steps 2 --warmup 3 --ramp-seconds 0 --no-cpu-baseline --lanes 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'clip_geometry' -s 3 -c 1 -f -o gpurun_out/${TAG}_geo $B > gpurun_out/${TAG}_geo.log 2>&1
