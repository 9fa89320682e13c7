#!/bin/bash
N=${1:-2}; TAG=${2:-r2bd$N}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/peer_breakdown.py 2>/dev/null | grep "^{" | tee gpurun_out/${TAG}_breakdown.json
