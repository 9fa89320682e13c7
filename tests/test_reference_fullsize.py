"""Parity at the sizes BASELINE.json names, pinned to OUTPUTS OF THE UNMODIFIED REFERENCE.

tests/golden/golden_fullsize.npz holds, for config 2 (40 frames x 6 cameras: nuScenes labels with exact and
slerp poses, CAMA labels ~1.0 M vertices) and for 12 sampled frames of the config-3 site, the SHA-256 of every
frame the reference rendered (main.py:57-59 over cama/dataset.py:78-117 and cama/reproject.py:246-257, blank
backgrounds) plus its visible-point and crop counts.  tests/golden/golden_composited.npz holds the same for
ClipManager.render_vectors WITH camera JPEGs (cama/dataset.py:119-126, cama/reproject.py:228-257).
tests/golden/make_golden.py generated both by importing /root/reference; nothing here reads it.

CPU tests pin the two oracles (NumPy/OpenCV restatement, plain-C restatement) to those digests; the -m gpu
tests pin the product: the batched render, the per-frame drop-in loop and the device image path.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import load_golden
from cama_b200 import synth
from oracle import cama_oracle as orc
from oracle import oracle_c

BOX6 = [orc.CROP_BOX[k] for k in ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")]
H, W = 540, 960
CONFIG3_SAMPLE = (0, 1, 31, 64, 97, 130, 159, 160, 201, 255, 288, 319)

CASES = {
    "config2_nuscenes_exact": (lambda: synth.config2_spec(seed=0, name="g_config2_exact"), "nuscenes", None),
    "config2_nuscenes_slerp": (lambda: synth.config2_spec(seed=0, pose_time_offset_ms=25, name="g_config2_slerp"), "nuscenes", None),
    "config2_cama_exact": (lambda: synth.config2_spec(seed=0, name="g_config2_cama"), "cama", None),
    "config3_nuscenes": (lambda: synth.config3_spec(seed=1, name="g_config3"), "nuscenes", CONFIG3_SAMPLE),
}


def sha_rows(frames):
    """uint8 [..., H, W, 3] -> uint8 [..., 32]"""
    lead = frames.shape[:-3]
    flat = np.ascontiguousarray(frames).reshape((-1,) + frames.shape[-3:])
    out = np.stack([np.frombuffer(hashlib.sha256(f.tobytes()).digest(), np.uint8) for f in flat])
    return out.reshape(lead + (32,))


def write_case(case, root):
    make, dataset, sample = CASES[case]
    spec = make()
    if dataset == "nuscenes":
        spec.write_cama = False
    else:
        spec.write_nuscenes = False
    return synth.write_clip(spec, str(root)), dataset, sample


def golden_case(case):
    g = load_golden("golden_fullsize.npz")
    return {k.split(".", 1)[1]: g[k] for k in g.files if k.startswith(case + ".")}


def bgr_of(classes):
    return np.array([orc.CLASS_RGB["lane_marking" if str(c) == "lane_marking" else "Crosswalk_Line"][::-1] for c in classes], np.uint8)


# ------------------------------------------------------------------------------------------------ CPU: the oracles
@pytest.mark.parametrize("case", list(CASES))
def test_c_oracle_matches_reference_digests(case, tmp_path):
    """oracle.c on every golden frame of every case (it is what the full-size GPU tests were checked against
    in round 1; this closes the chain C oracle -> reference at full size)."""
    clip, dataset, sample = write_case(case, tmp_path)
    g = golden_case(case)
    oc = orc.ClipOracle(synth.CAMA_CONFIGS, clip)
    flat, classes, counts = orc.flatten(oc.instance_maps[dataset], 3)
    assert len(flat) == int(g["n_vertices"]) and len(counts) == int(g["n_instances"])
    poses = oc.world_to_chassis_per_frame(dataset)
    keep = list(range(len(poses))) if sample is None else list(sample)
    assert [poses[k][0] for k in keep] == g["frame_idx"].tolist()
    w2c = np.stack([poses[k][1] for k in keep])
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    frames, cc, vc = oracle_c.clip_render(flat, offs, bgr_of(classes), w2c, np.stack(oc.chassis2cam), np.stack(oc.K), BOX6, H, W)
    assert np.array_equal(cc.sum(axis=-1), g["cropped"])
    assert np.array_equal(vc.sum(axis=-1), g["visible"])
    assert np.array_equal(frames.any(-1).sum(axis=(-1, -2)), g["lit"])
    assert np.array_equal(sha_rows(frames), g["sha256"])


def test_numpy_oracle_matches_reference_digests(tmp_path):
    """The NumPy/OpenCV restatement (what bench.py's reference arm times) on four config-2 frames."""
    clip, dataset, _ = write_case("config2_nuscenes_slerp", tmp_path)
    g = golden_case("config2_nuscenes_slerp")
    oc = orc.ClipOracle(synth.CAMA_CONFIGS, clip)
    for k, (image_idx, chassis) in enumerate(oc.frames(dataset)):
        if k not in (0, 13, 26, 39):
            continue
        assert image_idx == int(g["frame_idx"][k])
        assert sum(len(i["points"]) for i in chassis) == int(g["cropped"][k])
        per_cam = oc.project_all(chassis)
        for c, cam in enumerate(oc.cameras):
            assert sum(len(i["points"]) for i in per_cam[cam]) == int(g["visible"][k, c])
            img = orc.render_instances(np.zeros((H, W, 3), np.uint8), per_cam[cam])
            assert np.array_equal(sha_rows(img), g["sha256"][k, c])


@pytest.fixture(scope="module")
def composited_clip(tmp_path_factory):
    """The clip golden_composited.npz was made from, JPEGs included; skips when this box's JPEG codec does not
    reproduce the decoded images the fixture was computed from (the digests of those are in the fixture)."""
    import cv2
    root = str(tmp_path_factory.mktemp("composited"))
    spec = synth.tiny_spec(name="tiny_composited")
    clip = synth.write_clip(spec, root)
    synth.write_background_jpegs(clip, spec.n_frames, seed=4)
    g = load_golden("golden_composited.npz")
    oc = orc.ClipOracle(synth.CAMA_CONFIGS, clip)
    for k, image_idx in enumerate(g["nuscenes.frame_idx"].tolist()):
        for c, cam in enumerate(oc.cameras):
            raw = cv2.imread(os.path.join(clip, cam, f"{oc.attribute['sync'][cam][image_idx]}.jpg"))
            if not np.array_equal(sha_rows(raw), g["nuscenes.raw_sha256"][k, c]):
                pytest.skip("this box's JPEG codec decodes the synthetic camera images differently from the build container's")
    return clip


@pytest.mark.parametrize("dataset", ["nuscenes", "cama"])
def test_oracle_render_vectors_matches_reference(composited_clip, dataset):
    """R15 on the oracle: imread -> initUndistortRectifyMap -> remap -> in-place discs == the reference's images."""
    g = load_golden("golden_composited.npz")
    oc = orc.ClipOracle(synth.CAMA_CONFIGS, composited_clip)
    got_idx = []
    for k, (image_idx, chassis) in enumerate(oc.frames(dataset)):
        images = oc.render_vectors(oc.project_all(chassis), image_idx)
        assert list(images) == synth.CAMERA_LIST
        got_idx.append(image_idx)
        assert np.array_equal(sha_rows(np.stack([images[c] for c in oc.cameras])), g[f"{dataset}.sha256"][k])
        if image_idx == 1:
            assert np.array_equal(images["camera_front"], g[f"{dataset}.image_front"])
            assert np.array_equal(images["camera_rear"], g[f"{dataset}.image_rear"])
    assert got_idx == g[f"{dataset}.frame_idx"].tolist()


# ------------------------------------------------------------------------------------------------ GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_batched_render_matches_reference_digests(case, tmp_path):
    """Reproject.render_device / cama_clip_render at full size against what the reference rendered: frame
    digests, visible counts per (frame, camera), crop counts per frame."""
    import torch
    from cama_b200.batched import Reproject
    clip, dataset, sample = write_case(case, tmp_path)
    g = golden_case(case)
    rp = Reproject(synth.CAMA_CONFIGS, clip, device=0)
    idx, w2c = rp.frame_poses(dataset)
    res = rp.resident(dataset)
    assert res.n_vertices == int(g["n_vertices"]) and res.n_instances == int(g["n_instances"])
    keep = list(range(len(idx))) if sample is None else list(sample)
    assert [idx[k] for k in keep] == g["frame_idx"].tolist()
    if sample is None:                                         # the public call, host frames (sparse transfer)
        got_idx, frames = rp(dataset)
        assert got_idx == idx
        frames = frames[keep]
    else:                                                      # the site: render all 320 frames, look at the sampled ones
        frames = rp.render_device(dataset, w2c=w2c)[keep].cpu().numpy()
    assert np.array_equal(sha_rows(frames), g["sha256"])
    w2c_dev = torch.from_numpy(np.ascontiguousarray(w2c[keep])).cuda()
    out, dbg = rp.renderer.render(res, w2c_dev, mode="binned", debug=True)
    assert np.array_equal(dbg["crop_counts"].sum(dim=-1).cpu().numpy(), g["cropped"])
    assert np.array_equal(dbg["visible_counts"].sum(dim=-1).cpu().numpy(), g["visible"])
    assert np.array_equal(sha_rows(out.cpu().numpy()), g["sha256"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["config2_nuscenes_exact", "config2_nuscenes_slerp"])
def test_drop_in_loop_matches_reference_digests(case, tmp_path):
    """The three calls unmodified main.py makes per frame (yield_frame -> project_all_camera -> render_maps),
    every frame of config 2, against the reference's digests."""
    from cama_b200.dataset import ClipManager
    clip, dataset, _ = write_case(case, tmp_path)
    g = golden_case(case)
    cm = ClipManager(synth.CAMA_CONFIGS, clip, progress=False)
    seen = []
    for k, (image_idx, instance_map) in enumerate(cm.yield_frame(dataset)):
        maps_2d = cm.project_all_camera(instance_map)
        seen.append(image_idx)
        assert sum(len(i["points"]) for i in instance_map) == int(g["cropped"][k])
        for c, cam in enumerate(cm.cm_list):
            assert sum(len(i["points"]) for i in maps_2d[cam.camera_name]) == int(g["visible"][k, c])
            if k % 5 == 0:
                img = cam.render_maps(np.zeros((H, W, 3), np.uint8), maps_2d[cam.camera_name])
                assert np.array_equal(sha_rows(img), g["sha256"][k, c]), (k, c)
    assert seen == g["frame_idx"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("dataset", ["nuscenes", "cama"])
def test_render_vectors_matches_reference_images(composited_clip, dataset):
    """R15 / D5: the composited camera images.  Three product paths against the reference's own output:
    (1) ClipManager.render_vectors per frame (drop-in), (2) Reproject(...)(dataset, backgrounds=...) drawing in
    place on the undistort-resized host images, (3) render_device(raw_backgrounds=...): device undistort-resize
    of the decoded JPEGs + in-place draw."""
    import cv2
    import torch
    from cama_b200.batched import Reproject
    from cama_b200.dataset import ClipManager
    g = load_golden("golden_composited.npz")
    want = g[f"{dataset}.sha256"]
    cm = ClipManager(synth.CAMA_CONFIGS, composited_clip, progress=False)
    # (1) the per-frame drop-in calls
    idx = []
    for k, (image_idx, instance_map) in enumerate(cm.yield_frame(dataset)):
        images = cm.render_vectors(cm.project_all_camera(instance_map), image_idx)
        assert list(images) == synth.CAMERA_LIST
        idx.append(image_idx)
        assert np.array_equal(sha_rows(np.stack([images[n] for n in synth.CAMERA_LIST])), want[k]), k
        if image_idx == 1:
            assert np.array_equal(images["camera_front"], g[f"{dataset}.image_front"])
            assert np.array_equal(images["camera_rear"], g[f"{dataset}.image_rear"])
    assert idx == g[f"{dataset}.frame_idx"].tolist()
    # (2) batched, host backgrounds drawn on in place
    rp = Reproject(synth.CAMA_CONFIGS, composited_clip, device=0, clip_manager=cm)
    raw = np.stack([np.stack([cv2.imread(cam.get_image_path(i, True)) for cam in cm.cm_list]) for i in idx])
    backgrounds = np.stack([np.stack([cam.resize_image(raw[k, c]) for c, cam in enumerate(cm.cm_list)]) for k in range(len(idx))])
    got_idx, frames = rp(dataset, backgrounds=backgrounds)
    assert got_idx == idx and frames is backgrounds
    assert np.array_equal(sha_rows(frames), want)
    # (3) batched on the device from the decoded camera images
    dev = rp.render_device(dataset, raw_backgrounds=torch.from_numpy(raw).cuda()).cpu().numpy()
    assert np.array_equal(sha_rows(dev), want)
