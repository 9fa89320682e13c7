#!/bin/bash
# bench.py --gpus N only (it verifies the assembled frames against the single-GPU render itself) + the breakdown
N=${1:-8}; TAG=${2:-r2s$N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
[ "${BREAKDOWN:-1}" == "1" ] && timeout 300 $TR --master-port 29513 tools/peer_breakdown.py 2>/dev/null | grep "^{" | tee gpurun_out/${TAG}_breakdown.json
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-30} --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err | grep -v "^\*\|OMP_NUM"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench.json') if l.startswith('{')][-1])
    print('value %.0f (%.3f ms) single_gpu %.0f (%.3f ms) compute_only %.3f ms dense_allgather %.3f ms e2e %.0f (%.3f ms)' % (d['value'], d['ms_per_step'], d['single_gpu_same_workload']['value'], d['single_gpu_same_workload']['ms_per_step'], d['compute_only']['ms_per_step'], d['dense_allgather']['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
    print(d['verified'], d['config']['assembly'][:40], d['roofline']['launch_ms'], d['roofline']['frac'], d['exchange'])
except Exception as e: print('ERR', e)
PY
