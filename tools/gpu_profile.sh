#!/bin/bash
# usage: tools/gpu_profile.sh NAME KERNEL_REGEX [ENV=...]   -> gpurun_out/NAME.ncu-rep (one launch, ncu --set full)
NAME=$1; shift; K=$1; shift
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -f -o gpurun_out/$NAME \
  python bench.py --steps 2 --warmup 3 --ramp-seconds 0 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/$NAME.log 2>&1
tail -2 gpurun_out/$NAME.log
