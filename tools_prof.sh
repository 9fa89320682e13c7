B="python bench.py --steps 2 --warmup 3 --ramp-seconds 0 --no-cpu-baseline"
CAMA_BAND_ROWS=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'binned_raster' -s 3 -c 1 -f -o gpurun_out/r1_prof4 $B > gpurun_out/r1_prof4.log 2>&1
tail -2 gpurun_out/r1_prof4.log
