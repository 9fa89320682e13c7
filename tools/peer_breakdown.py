#!/usr/bin/env python
"""Where the time of an assembled step goes (run under torchrun, N GPUs): every phase of
shard.PeerExchange.render_and_assemble on the config-3 site timed alone with CUDA events, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/peer_breakdown.py
"""
import ctypes, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from cama_b200 import shard, synth, _native as N
from cama_b200.batched import Reproject

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
root = tempfile.mkdtemp()
spec = synth.config3_spec(); spec.write_cama = False
clip = synth.write_clip(spec, root)
rp = Reproject(synth.CAMA_CONFIGS, clip, device=local)
rt, r = rp.rt, rp.renderer
asm = shard.SiteAssembler(rp, "nuscenes", exchange="chunks")
px = asm.exchange
F, C = asm.n_frames, r.n_cams
out = torch.empty((F, C, 540, 960, 3), dtype=torch.uint8, device=rt.device)
stream = torch.cuda.current_stream()


def timed(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=rt.device)
    lo = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return round(float(t.item()), 4), round(float(lo.item()), 4)


res = {"world": world, "frames_per_rank": asm.hi - asm.lo, "capacity": px.capacity}
res["clear_library"] = timed(lambda: N.check(N.lib().cama_frames_clear(rt.ctx, rt.ptr(out), out.numel(), rt.stream())))
res["clear_torch_zero_"] = timed(lambda: out.zero_())
local_frames = torch.empty((asm.hi - asm.lo, C, 540, 960, 3), dtype=torch.uint8, device=rt.device)
r.render(asm.res, asm.w2c_dev, out=local_frames, check=True)
res["render_dense_block"] = timed(lambda: r.render(asm.res, asm.w2c_dev, out=local_frames, check=False))
hdr = N.PEER_HEADER_BYTES
own = px.slot(rank, 0, rank)
ov_local = {"records_ptr": own + hdr, "count_ptr": px.count.data_ptr(), "capacity": px.capacity, "fmt": px.fmt, "image_base": asm.lo * C}
res["render_sparse_local_only"] = timed(lambda: r.enqueue_overlay(asm.res, asm.w2c_dev, ov_local))
ov_mirror = dict(ov_local, mirrors=[px.slot(q, 0, rank) + hdr for q in range(world) if q != rank])
res["render_sparse_mirrored"] = timed(lambda: r.enqueue_overlay(asm.res, asm.w2c_dev, ov_mirror))
# publish once (step 1000, parity 0), then expand the same published slots again and again
headers = (ctypes.c_void_p * world)(*[px.slot(q, 0, rank) for q in range(world)])
N.check(N.lib().cama_peer_publish(rt.ctx, px.count.data_ptr(), 1000, headers, world, rt.stream()))
torch.cuda.synchronize(); dist.barrier()
slots = (ctypes.c_void_p * world)(*[px.slot(rank, 0, q) for q in range(world)])
pal = px._palette_dev(asm.res); scratch = rt.scratch("palette32", 1024)
expand = lambda: N.check(N.lib().cama_peer_expand(rt.ctx, slots, world, rank, 1000, px.capacity, px.fmt, rt.ptr(pal), rt.ptr(scratch), rt.ptr(out), F, C, 540, 960, 0,
                                                  px.status.data_ptr(), rt.stream()))
res["expand_all_slots"] = timed(expand)
px.step = 1000
res["full_step_overlapped_clear"] = timed(lambda: asm.step(out=out))
rs, asm.render_stream = asm.render_stream, None
res["full_step_serial_clear"] = timed(lambda: asm.step(out=out))
asm.render_stream = rs
res["status"] = px.status_code()
dist.barrier()
# ---- the other exchange: centre records into everybody's per-band lists, every rank rasters every frame
asm2 = shard.SiteAssembler(rp, "nuscenes", exchange="lists")
res["lists_available"] = bool(asm2.available and asm2.kind == "lists")
if res["lists_available"]:
    lx = asm2.exchange
    res["lists_capacity"], res["lists_bytes_per_rank"] = lx.capacity, lx.total_bytes
    res["lists_full_step"] = timed(lambda: asm2.step(out=out))
    res["lists_raster_all_frames"] = timed(lambda: lx.reraster(r, asm2.res, out))
    own = lx.base[rank] + (lx.step & 1) * lx.parity_bytes
    geo = {"phases": N.PHASE_GEOMETRY, "records_ptr": own + lx.records_off, "cursor_ptr": own + lx.cursor_off, "frame_base": asm2.lo, "frames": lx.n_frames}
    res["lists_geometry_local_only"] = timed(lambda: r.enqueue_phase(asm2.res, asm2.w2c_dev, asm2.hi - asm2.lo, geo, lx.capacity))
    dist.barrier()
    # phase times INSIDE full steps (events on the stream): geometry, publish, wait for the peers, raster
    import time
    dist.barrier(); torch.cuda.synchronize()
    all_marks, t0 = [], time.perf_counter()
    for _ in range(12):
        marks = []
        lx.render_and_assemble(r, asm2.res, asm2.w2c_dev, asm2.lo, out, marks=marks)
        all_marks.append(marks)
    host_ms = (time.perf_counter() - t0) / 12 * 1e3
    torch.cuda.synchronize()
    names = ("geometry", "push+publish", "wait", "raster")
    res["lists_in_step_ms"] = {n: round(sum(m[i].elapsed_time(m[i + 1]) for m in all_marks[2:]) / len(all_marks[2:]), 4) for i, n in enumerate(names)}
    res["lists_in_step_ms"]["step_to_step"] = round(sum(a[0].elapsed_time(b[0]) for a, b in zip(all_marks[2:-1], all_marks[3:])) / (len(all_marks) - 3), 4)
    res["lists_host_enqueue_ms_per_step"] = round(host_ms, 4)
    res["lists_status"] = lx.status_code()
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
px.close()
if asm2.available:
    asm2.exchange.close()
dist.destroy_process_group()
