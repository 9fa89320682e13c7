#!/usr/bin/env python
"""Extra parity mileage on the GPU box: config-2 clips of several seeds (exact-hit and slerp pose variants) through
every output path — dense BINNED, PLANE, frame-group pipeline, sparse records + host draw, sparse records + device
expand — against the C oracle (oracle/oracle.c), frame by frame, byte for byte.

    python tools/parity_sweep.py [n_seeds]
"""
import os, sys, tempfile
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch
from cama_b200 import synth
from cama_b200.batched import Reproject
from cama_b200.reproject import render_bgr_of_class
from oracle import cama_oracle as orc, oracle_c

H, W = 540, 960
BOX6 = [-50, 50, -100, 100, -200, 200]
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bad = 0
for seed in range(1, n_seeds + 1):
    for off in (0, 25):
        with tempfile.TemporaryDirectory() as root:
            spec = synth.config2_spec(seed=seed, pose_time_offset_ms=off, name=f"sweep{seed}_{off}")
            spec.write_cama = False
            clip = synth.write_clip(spec, root)
            rp = Reproject(synth.CAMA_CONFIGS, clip, device=0)
            oc = orc.ClipOracle(synth.CAMA_CONFIGS, clip)
            flat, classes, counts = orc.flatten(oc.instance_maps["nuscenes"], 3)
            w2c = np.stack([m for _, m in oc.world_to_chassis_per_frame("nuscenes")])
            offs = np.concatenate([[0], np.cumsum(counts)])
            bgr = np.array([render_bgr_of_class(str(c)) for c in classes], np.uint8)
            want, _, _ = oracle_c.clip_render(flat, offs, bgr, w2c, np.stack(oc.chassis2cam), np.stack(oc.K), BOX6, H, W)
            _, host_w2c = rp.frame_poses("nuscenes")
            ok = {"poses": np.array_equal(host_w2c.reshape(-1, 4, 4), w2c.astype(np.float32))}
            r = rp.renderer
            ok["binned"] = np.array_equal(rp.render_device("nuscenes", mode="binned").cpu().numpy(), want)
            ok["plane"] = np.array_equal(rp.render_device("nuscenes", mode="plane").cpu().numpy(), want)
            r.pipeline_frames = 16
            ok["pipeline"] = np.array_equal(rp.render_device("nuscenes").cpu().numpy(), want)
            r.pipeline_frames = 0
            ok["sparse_host"] = np.array_equal(rp("nuscenes")[1], want)
            res = rp.resident("nuscenes")
            w2c_dev = torch.from_numpy(host_w2c).to(rp.rt.device)
            records, n, fmt = r.render_overlay(res, w2c_dev)
            ok["sparse_expand"] = np.array_equal(r.expand_overlay(records, n, fmt, res.palette, len(w2c)).cpu().numpy(), want)
            lit = int((want.reshape(-1, 3).any(axis=1)).sum())
            print(f"seed {seed} pose offset {off:2d} ms: {len(flat)} vertices, {lit} lit pixels  " + " ".join(f"{k}={'ok' if v else 'MISMATCH'}" for k, v in ok.items()), flush=True)
            bad += sum(not v for v in ok.values())
print("parity sweep:", "OK" if bad == 0 else f"{bad} MISMATCHES")
sys.exit(1 if bad else 0)
