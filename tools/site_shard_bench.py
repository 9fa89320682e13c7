#!/usr/bin/env python
"""BASELINE.json configs[3]: the config-3 site (320 frames x 6 cams, 1600 polylines) sharded by frame over the GPUs of
one box (strong scaling), without assembly, with the dense all-gather of uint8 frames north_star names, and with the
sparse assembly (records all-gathered + cama_overlay_expand).  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 tools/site_shard_bench.py
"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from cama_b200 import shard, synth
from cama_b200.batched import Reproject

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
root = tempfile.mkdtemp()
spec = synth.config3_spec(); spec.write_cama = False
clip = synth.write_clip(spec, root)
rp = Reproject(synth.CAMA_CONFIGS, clip, device=local)
idx, w2c = rp.frame_poses("nuscenes")
F, C = len(idx), rp.renderer.n_cams
lo, hi = shard.frame_block(F, rank, world)
w2c_dev = torch.from_numpy(np.ascontiguousarray(w2c[lo:hi])).to(rp.rt.device)
res = rp.resident("nuscenes")
local_frames = rp.renderer.render(res, w2c_dev, check=True)            # settles capacities

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def timed(fn, reps=10):
    for _ in range(2):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    barrier()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt

out = {"workload": "configs[3]: site, 320 frames x 6 cams sharded by frame", "n_gpus": world, "frames": F, "frames_per_rank": hi - lo}
dt = timed(lambda: rp.renderer.render(res, w2c_dev, out=local_frames, check=False))
out["no_assembly"] = {"ms": round(dt * 1e3, 3), "cam_frames_per_s": round(F * C / dt)}
if world > 1:
    whole = None
    def dense():
        global whole
        rp.renderer.render(res, w2c_dev, out=local_frames, check=False)
        whole = shard.gather_frames(local_frames, F)
    dt = timed(dense, reps=5)
    out["dense_allgather"] = {"ms": round(dt * 1e3, 3), "cam_frames_per_s": round(F * C / dt)}
    ref_sum = int(whole[:, :, ::9, ::9].sum().item())
    del whole
    got = None
    def sparse():
        global got
        _, got = shard.render_sharded(rp, "nuscenes", gather="sparse")
    dt = timed(sparse, reps=5)
    out["sparse_allgather"] = {"ms": round(dt * 1e3, 3), "cam_frames_per_s": round(F * C / dt), "same_checksum_as_dense": int(got[:, :, ::9, ::9].sum().item()) == ref_sum,
                               "note": "includes the host pose lookup of render_sharded and the record-count read-back"}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
