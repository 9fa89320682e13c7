#!/usr/bin/env python
"""Multi-GPU parity (run under torchrun on a box with N GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

Every rank renders its frame block of the same synthetic config-2 clip, the blocks are assembled on every rank
three ways — NCCL all-gather of the dense frames, NCCL all-gather of the sparse records + device expand, and the
peer-memory exchange (the raster mirrors its records into the peers' mailboxes; cama_peer_publish / cama_peer_expand;
several steps in a row, so both mailbox parities and the slot reuse are exercised) — and every rank checks the
assembled clip bit for bit against its own single-GPU render; then the LiDAR sweeps of a clip are sharded and the
all-reduced voxel counts are checked against the single-GPU aggregation.  Rank 0 prints one JSON line (kept under
profiles/).
"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from cama_b200 import shard, synth
from cama_b200.batched import Reproject


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    with tempfile.TemporaryDirectory() as root:
        spec = synth.config2_spec(n_frames=37)            # not a multiple of the world size: padded last block
        spec.write_cama = False
        clip = synth.write_clip(spec, root)
        rp = Reproject(synth.CAMA_CONFIGS, clip, device=local)
        idx, gathered = shard.render_sharded(rp, "nuscenes", gather=True)
        whole = rp.render_device("nuscenes")
        ok = len(idx) == 37 and bool((gathered == whole).all())
        idx_s, gathered_s = shard.render_sharded(rp, "nuscenes", gather="sparse")      # records over NCCL + cama_overlay_expand
        ok = ok and idx_s == idx and bool((gathered_s == whole).all())
        lo, hi = shard.frame_block(37, rank, world)
        idx_local, block = shard.render_sharded(rp, "nuscenes", gather=False)
        ok = ok and idx_local == idx[lo:hi] and bool((block == whole[lo:hi]).all())
        report = {"world": world, "frames": 37, "dense_nccl_allgather": bool((gathered == whole).all()), "sparse_nccl_allgather": bool((gathered_s == whole).all())}
        # peer-memory assembly: five steps in a row into the same output, every one checked; then a poisoned output buffer
        peer_ok = True
        for kind in ("lists", "chunks"):
            asm = shard.SiteAssembler(rp, "nuscenes", exchange=kind)
            report[f"peer_{kind}_available"] = bool(asm.available) and asm.kind == kind
            report[f"peer_{kind}_error"] = asm.exchange.error
            kind_ok = asm.available and asm.kind == kind
            if asm.available:
                for step in range(5):
                    out = asm.step()
                    if step == 3:
                        torch.cuda.synchronize()
                        out.fill_(0x5a)                               # the next step must overwrite every byte
                        out = asm.step()
                    torch.cuda.synchronize()
                    kind_ok = kind_ok and asm.exchange.status_code() == 0 and bool((out == whole).all())
                dist.barrier()
                asm.exchange.close()
            report[f"peer_{kind}_assembly"] = bool(kind_ok)
            peer_ok = peer_ok and kind_ok
        asm = shard.SiteAssembler(rp, "nuscenes")
        report["peer_available"] = bool(asm.available)
        report["peer_error"] = asm.exchange.error
        if asm.available:
            idx_p, gathered_p = shard.render_sharded(rp, "nuscenes", gather="peer")    # the public entry
            peer_ok = peer_ok and idx_p == idx and bool((gathered_p == whole).all())
            dist.barrier()
            asm.exchange.close()
        report["peer_assembly"] = peer_ok
        ok = ok and peer_ok
        # LiDAR aggregation (configs[4]): sweeps sharded across the ranks, voxel counts summed with one all-reduce
        from cama_b200.lidar import LidarAggregator, allreduce_counts
        synth.write_lidar_sweeps(clip, n_sweeps=37, n_points=3000, seed=5, ragged=True)      # (every rank has its own copy of the clip)
        agg = LidarAggregator(device=local)
        whole_counts, whole_inside, _ = agg.aggregate_clip(rp.cm)
        mine, _, _ = agg.aggregate_clip(rp.cm, rank=rank, world_size=world)
        total = allreduce_counts(mine)
        ok = ok and whole_inside > 0 and bool((total == whole_counts).all())
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            report["lidar_allreduce"] = bool((total == whole_counts).all())
            report["all_ranks_ok"] = bool(int(flag.item()))
            print(json.dumps(report), flush=True)
            print(f"multi_gpu_check world={world}: {'OK' if int(flag.item()) else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
