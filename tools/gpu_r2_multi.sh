#!/bin/bash
# round 2, multi-GPU run: N-GPU parity of the three assemblies, then bench.py --gpus N (configs[3] sharded by frame)
N=${1:-2}; TAG=${2:-r2m$N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/multi_gpu_check.py > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?"; grep -E "^\{|multi_gpu_check|Error|error" gpurun_out/${TAG}_check.log | tail -5
timeout 300 $TR --master-port 29513 tools/peer_breakdown.py 2>/dev/null | grep "^{" | tee gpurun_out/${TAG}_breakdown.json
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-30} --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench.json') if l.startswith('{')][-1])
    print('value %.0f (%.3f ms) single_gpu %.0f (%.3f ms) compute_only %.3f ms dense_allgather %.3f ms e2e %.0f (%.3f ms)' % (d['value'], d['ms_per_step'], d['single_gpu_same_workload']['value'], d['single_gpu_same_workload']['ms_per_step'], d['compute_only']['ms_per_step'], d['dense_allgather']['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
    print(d['verified'], d['config']['assembly'][:60], d['roofline']['launch_ms'], d['roofline']['frac'])
except Exception as e: print('ERR', e)
PY
