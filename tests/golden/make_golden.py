#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Only runnable in the build container (needs /root/reference, numpy, scipy, opencv);
the fixtures it writes are committed so that nothing at test/bench time reads
/root/reference.  Usage:  python tests/golden/make_golden.py

What is recorded (all produced by reference code, never by this repo's code):

* golden_known_answers.npz   survey section 8(c) known-answer facts
* golden_pose.npz            pose_transformer outputs on seeded inputs
* golden_lidar.npz           the reference-side inputs of the LiDAR aggregation: sweeps read by the
                             reference's DatasetReader.yield_lidar from a synth clip, lidar->chassis from its
                             get_extrinsic, chassis->world from its ClipManager trajectory + seek_by_timestamp, and
                             the world points from its MapManager.transform_3d_instance_maps.  (The voxel
                             accumulation itself does not exist in the reference: parity unpinned for that step.)
* golden_clip_<dataset>_<variant>.npz
      full per-frame outputs of ClipManager.yield_frame / project_all_camera /
      CameraManager.render_maps on synth.tiny_spec clips (exact-hit and slerp pose variants)

* golden_fullsize.npz        the reference at the sizes BASELINE.json names: per-(frame, camera) SHA-256 of the rendered
                             frame, visible-point counts and per-frame crop counts for config 2 (all 40 frames x 6
                             cameras; nuScenes labels with exact and slerp poses, CAMA labels with exact poses) and for
                             12 sampled frames of the config-3 site (320 frames, 767 k vertices)
* golden_composited.npz      ClipManager.render_vectors on a 3-frame clip WITH camera JPEGs (cv2.imread ->
                             initUndistortRectifyMap -> remap -> render_maps in place): SHA-256 per camera-frame, two
                             full images, and digests of the decoded JPEGs the result was computed from

The synthetic inputs are regenerated at test time from cama_b200.synth (seeded), so the
fixtures hold outputs plus a digest of the inputs they were computed from.
"""
import hashlib
import os
import sys
import tempfile
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)
sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))   # cama/tools.py:1 imports it, nothing here uses it

import numpy as np  # noqa: E402
from cama.dataset import ClipManager  # noqa: E402  (reference)
from cama.reproject import MapManager, CameraManager  # noqa: E402  (reference)
from cama import pose_transformer as ref_pt  # noqa: E402  (reference)
from cama_b200 import synth  # noqa: E402


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def pack_instances(instances, dtype=None):
    pts = [np.asarray(i["points"]) for i in instances]
    offsets = np.zeros(len(pts) + 1, np.int64)
    offsets[1:] = np.cumsum([len(p) for p in pts])
    flat = np.concatenate(pts, 0) if pts else np.zeros((0, 3))
    if dtype is not None:
        assert flat.dtype == dtype, flat.dtype
    classes = np.array([i["class"] for i in instances])
    return np.ascontiguousarray(flat), offsets, classes


def clip_golden(spec, dataset, out_path, root):
    clip = synth.write_clip(spec, root)
    cm = ClipManager(synth.CAMA_CONFIGS, clip)
    instances = cm.instance_maps[dataset]
    inst_points, inst_offsets, inst_classes = pack_instances(instances)
    n_inst = len(instances)
    n_cam = len(cm.cm_list)

    frame_idx, w2c = [], []
    crop_counts, crop_points = [], []
    vu_counts, vu_points = [], []
    frames = []
    # world2chassis per frame, recomputed exactly as cama/dataset.py:84-99 does
    from cama.dataset_reader import DatasetReader
    dr = DatasetReader(clip)
    pt = cm.get_pt_nuscenes(dr) if dataset == "nuscenes" else cm.get_pt_cama(dr)
    stamps = dr.get_sensor_timestamp("camera_front", sync=True)

    for image_idx, instance_map in cm.yield_frame(dataset):
        frame_idx.append(image_idx)
        chassis2world = pt.seek_by_timestamp(stamps[image_idx], t_max_diff=0.5, interpolate=True).astype(np.float32)
        world2chassis = np.linalg.inv(chassis2world)
        w2c.append(world2chassis)
        # per-instance survivors (reference API called on one-instance lists keeps the ids)
        counts = np.zeros(n_inst, np.int32)
        chassis_per_inst = []
        for i, inst in enumerate(instances):
            got = cm.mm.crop_3d_instance_maps(cm.mm.transform_3d_instance_maps([inst], world2chassis))
            counts[i] = len(got[0]["points"]) if got else 0
            chassis_per_inst.append(got[0] if got else None)
        flat, _, _ = pack_instances(instance_map)
        assert counts.sum() == len(flat)
        crop_counts.append(counts)
        crop_points.append(flat)

        maps_2d = cm.project_all_camera(instance_map)
        vcounts = np.zeros((n_cam, n_inst), np.int32)
        cam_frames = []
        for c, cam in enumerate(cm.cm_list):
            for i, inst in enumerate(chassis_per_inst):
                if inst is None:
                    continue
                got = cam.project_to_image(cm.mm.transform_3d_instance_maps([inst], cam.get_chassis2camera()))
                vcounts[c, i] = len(got[0]["points"]) if got else 0
            vu_flat = [np.asarray(i["points"]) for i in maps_2d[cam.camera_name]]
            vu_flat = np.concatenate(vu_flat, 0) if vu_flat else np.zeros((0, 2))
            assert vcounts[c].sum() == len(vu_flat)
            vu_points.append(np.ascontiguousarray(vu_flat))
            cam_frames.append(cam.render_maps(np.zeros((cam.height, cam.width, 3), np.uint8),
                                              maps_2d[cam.camera_name]))
        vu_counts.append(vcounts)
        frames.append(np.stack(cam_frames))

    np.savez_compressed(
        out_path,
        input_digest=np.array(digest(inst_points, inst_offsets)),
        inst_points=inst_points, inst_offsets=inst_offsets, inst_classes=inst_classes,
        frame_idx=np.array(frame_idx, np.int32), world2chassis=np.stack(w2c),
        crop_counts=np.stack(crop_counts), crop_points=np.concatenate(crop_points, 0),
        vu_counts=np.stack(vu_counts), vu_points=np.concatenate(vu_points, 0),
        frames=np.stack(frames),
        K=np.stack([c.K for c in cm.cm_list]),
        chassis2camera=np.stack([np.asarray(c.get_chassis2camera(), np.float64) for c in cm.cm_list]),
    )
    print(out_path, "frames", frame_idx, "N", len(inst_points), "crop", int(np.stack(crop_counts).sum()),
          "visible", int(np.stack(vu_counts).sum()), "lit", int(np.stack(frames).any(-1).sum()))


def known_answers(out_path, root):
    mm = MapManager()
    out = {}
    # fact (4): densify semantics (f32, end point dropped, sub-resolution segment dropped)
    labels = [{"attrs": {"type": "lane_marking"}, "data": [[0, 0], [1, 0], [1.05, 0], [1.35, 0]]}]
    dens = mm.load_3d_instance_maps(labels)
    out["densify_points"] = dens[0]["points"]
    # fact (1)-(3): radius-2 disc stamp incl. border clipping, painter's order
    import cv2
    img = np.zeros((9, 9, 3), np.uint8)
    cv2.circle(img, (4, 4), 2, (7, 8, 9), -1)
    out["stamp_centre"] = img
    img = np.zeros((9, 9, 3), np.uint8)
    cv2.circle(img, (0, 8), 2, (7, 8, 9), -1)
    cv2.circle(img, (9, -1), 2, (1, 2, 3), -1)
    out["stamp_clipped"] = img
    # fact (6)/(7): config 1 anchor — 50 raw points handed over without densify, identity pose
    clip = synth.write_clip(synth.config1_spec(), root)
    cam = CameraManager(clip, "camera_front")
    out["K_scaled_front"] = cam.K
    xs = np.arange(5.0, 55.0, 1.0)
    pts = np.stack([xs, np.full_like(xs, 1.8), np.zeros_like(xs)], 1).astype(np.float32)
    inst = [{"class": "lane_marking", "points": pts}]
    chassis = mm.crop_3d_instance_maps(mm.transform_3d_instance_maps(inst, np.eye(4, dtype=np.float32)))
    cam_pts = mm.transform_3d_instance_maps(chassis, cam.get_chassis2camera())
    vu = cam.project_to_image(cam_pts)
    image = cam.render_maps(np.zeros((540, 960, 3), np.uint8), vu)
    out["config1_points"] = pts
    out["config1_crop"] = chassis[0]["points"]
    out["config1_vu"] = np.ascontiguousarray(vu[0]["points"])
    out["config1_image"] = image
    print("config1 anchor: cropped", len(chassis[0]["points"]), "visible", len(vu[0]["points"]),
          "lit", int(image.any(-1).sum()))
    # CAMA-label load path (reference cama/reproject.py:72-106) on a small BEV height map
    rng = np.random.default_rng(5)
    bev = (rng.standard_normal((400, 400)) * 0.05).astype(np.float32)
    labels = [{"attrs": {"type": "Road_teeth"}, "data": (rng.uniform(-3, 402, size=(6, 2))).round(2).tolist()},
              {"attrs": {"type": "lane_marking"}, "data": [[10.0, 20.0]]},
              {"attrs": {"type": "lane_marking"}, "data": [[10.5, 20.5], [10.5, 31.25], [10.52, 31.25], [30.0, 12.0]]}]
    out["cama_labels_0"] = np.array(labels[0]["data"])
    out["cama_labels_2"] = np.array(labels[2]["data"])
    out["cama_bev"] = bev
    cal = mm.calculate_3d_instance_maps(bev, labels)
    assert len(cal) == 2
    out["cama_points_0"] = cal[0]["points"]
    out["cama_points_1"] = cal[1]["points"]
    np.savez_compressed(out_path, **out)
    print(out_path)


def pose_golden(out_path):
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(11)
    n = 9
    tum = np.zeros((n, 8))
    tum[:, 0] = 100.0 + np.cumsum(rng.uniform(0.05, 0.4, n))
    tum[:, 1:4] = rng.normal(size=(n, 3)) * 20
    q = rng.normal(size=(n, 4))
    tum[:, 4:8] = q / np.linalg.norm(q, axis=1, keepdims=True)
    ext = np.eye(4)
    ext[:3, :3] = R.from_euler("zyx", [0.3, -0.2, 0.1]).as_matrix()
    ext[:3, 3] = [0.5, -1.0, 2.0]
    out = {"tum": tum, "ext": ext}

    out["invT"] = ref_pt.invT(ext)
    pt = ref_pt.PoseTransformer()
    pt.loadarray(tum)
    out["abs"] = pt.as_transform(absolute=True)
    out["rel"] = pt.as_transform(absolute=False)
    out["quats"] = np.asarray(pt.as_quaternions())
    out["euler_abs"] = pt.as_euler(absolute=True)
    out["euler_rel"] = pt.as_euler(absolute=False)
    out["axis_abs"] = pt.as_axis_angle(absolute=True)
    out["axis_rel"] = pt.as_axis_angle(absolute=False)
    out["trans_abs"] = pt.as_translations(absolute=True)
    out["trans_rel"] = pt.as_translations(absolute=False)
    out["trans_quat"] = pt.as_trans_quat()
    out["dump_tum"] = pt.dumparray()

    queries = [float(tum[0, 0]), float(tum[3, 0]), float(tum[3, 0] + 1e-10),
               float(0.25 * tum[2, 0] + 0.75 * tum[3, 0]), float(0.5 * (tum[6, 0] + tum[7, 0])),
               float(tum[-1, 0]), float(tum[0, 0] - 0.01), float(tum[-1, 0] + 0.01), float(tum[0, 0] - 5e-10)]
    out["queries"] = np.array(queries)
    for mode, interp in (("interp", True), ("nearest", False)):
        res, ok = [], []
        for qt in queries:
            try:
                res.append(np.asarray(pt.seek_by_timestamp(qt, 0.5, interpolate=interp), np.float64))
                ok.append(1)
            except RuntimeError:
                res.append(np.full((4, 4), np.nan))
                ok.append(0)
        out[f"seek_{mode}"] = np.stack(res)
        out[f"seek_{mode}_ok"] = np.array(ok)
    # a t_max_diff small enough that some interpolations are refused
    res = []
    for qt in queries:
        try:
            pt.seek_by_timestamp(qt, 0.1, interpolate=True)
            res.append(1)
        except RuntimeError:
            res.append(0)
    out["seek_interp_tight_ok"] = np.array(res)
    out["slerp"] = ref_pt.SlerpTransform(out["abs"][1], out["abs"][2], 0.3)

    for name in ("right_rotate", "left_rotate", "transform", "normalize2center", "normalize2origin"):
        p = ref_pt.PoseTransformer()
        p.loadarray(tum)
        getattr(p, name)(*(() if name.startswith("normalize") else (ext,)))
        out[name] = p.as_transform(absolute=True)

    # kitti / asl loaders, relative-input constructors
    kitti = out["abs"][:, :3, :].reshape(n, 12)
    p = ref_pt.PoseTransformer()
    p.loadarray(kitti, style="kitti")
    out["kitti_abs"] = p.as_transform(True)
    out["kitti_rel"] = p.as_transform(False)
    asl = np.zeros((n, 17))
    asl[:, 0] = tum[:, 0] * 1e9
    asl[:, 1:4] = tum[:, 1:4]
    asl[:, 4] = tum[:, 7]
    asl[:, 5:8] = tum[:, 4:7]
    p = ref_pt.PoseTransformer()
    p.loadarray(asl, style="asl")
    out["asl"] = asl
    out["asl_abs"] = p.as_transform(True)
    out["asl_ts"] = p.get_timestamps()
    rel_eulers = rng.normal(size=(5, 3)) * 0.2
    rel_trans = rng.normal(size=(5, 3))
    p = ref_pt.PoseTransformer()
    p.from_relative_eulers(rel_eulers)
    p.from_translation(rel_trans, absolute=False)
    out["rel_eulers"] = rel_eulers
    out["rel_trans"] = rel_trans
    out["from_rel_abs"] = p.as_transform(True)
    p = ref_pt.PoseTransformer()
    p.from_axis_angle(rel_eulers, absolute=True)
    p.from_translation(rel_trans, absolute=True)
    out["from_abs_axis_abs"] = p.as_transform(True)
    p = ref_pt.PoseTransformer()
    p.from_relative_quaternion(tum[:5, 4:8])
    p.from_relative_translation(rel_trans)
    out["from_rel_quat_abs"] = p.as_transform(True)
    p = ref_pt.PoseTransformer()
    p.from_relative_axis_angle(rel_eulers)
    p.from_relative_translation(rel_trans)
    p.load_timestamp(list(tum[:5, 0][::-1]))
    p.sort_by_timestamps()
    out["sorted_rel"] = p.as_transform(False)
    out["sorted_ts"] = p.get_timestamps()
    np.savez_compressed(out_path, **out)
    print(out_path)


def lidar_golden(out_path, root):
    """Everything the reference computes on the way to a LiDAR aggregation (pose offset 25 ms: slerp branch)."""
    from cama.dataset_reader import DatasetReader  # reference
    spec = synth.tiny_spec(n_frames=5, pose_time_offset_ms=25, name="tiny_lidar")
    clip = synth.write_clip(spec, root)
    synth.write_lidar_sweeps(clip, n_sweeps=4, n_points=300, seed=3, ragged=True)
    dr = DatasetReader(clip)
    lidar2chassis = dr.get_extrinsic("lidar_top", "chassis")
    cm = ClipManager(synth.CAMA_CONFIGS, clip)
    pt = cm.get_pt_nuscenes(dr)
    mm = MapManager()
    out = {"lidar2chassis": np.asarray(lidar2chassis, dtype=np.float64)}
    stamps, sizes, pts_all, world_all, mats = [], [], [], [], []
    for stamp, cloud in dr.yield_lidar():
        chassis2world = pt.seek_by_timestamp(float(stamp), t_max_diff=0.5, interpolate=True)
        T = chassis2world @ lidar2chassis
        world = mm.transform_3d_instance_maps([{"class": "lidar", "points": cloud[:, :3]}], T)[0]["points"] if len(cloud) else np.zeros((0, 3))
        stamps.append(stamp); sizes.append(len(cloud)); pts_all.append(cloud); world_all.append(np.ascontiguousarray(world)); mats.append(T)
    out.update(stamps=np.array(stamps), sizes=np.array(sizes, np.int64), points=np.concatenate(pts_all, 0), world=np.concatenate(world_all, 0),
               transforms=np.stack(mats), inputs_digest=np.array(digest(np.concatenate(pts_all, 0))))
    np.savez_compressed(out_path, **out)
    print(out_path)


def frame_digests(frames):
    """uint8 [n, 32]: SHA-256 of each image of `frames` (C-contiguous bytes)."""
    return np.stack([np.frombuffer(hashlib.sha256(np.ascontiguousarray(f).tobytes()).digest(), np.uint8) for f in frames])


def fullsize_case(spec, dataset, root, sample=None):
    """The unmodified reference loop (main.py:57-59 with blank frames) on a full-size clip ->
    dict of per-frame / per-camera-frame digests and counts.  sample: positions (in yield order) to keep."""
    clip = synth.write_clip(spec, root)
    cm = ClipManager(synth.CAMA_CONFIGS, clip)
    n_cam = len(cm.cm_list)
    idx, digests, visible, cropped, lit = [], [], [], [], []
    for k, (image_idx, instance_map) in enumerate(cm.yield_frame(dataset)):
        if sample is not None and k not in sample:
            continue
        maps_2d = cm.project_all_camera(instance_map)
        images = [cam.render_maps(np.zeros((cam.height, cam.width, 3), np.uint8), maps_2d[cam.camera_name]) for cam in cm.cm_list]
        idx.append(image_idx)
        cropped.append(sum(len(i["points"]) for i in instance_map))
        visible.append([sum(len(i["points"]) for i in maps_2d[cam.camera_name]) for cam in cm.cm_list])
        lit.append([int(im.any(-1).sum()) for im in images])
        digests.append(frame_digests(images))
    n_vertices = sum(len(i["points"]) for i in cm.instance_maps[dataset])
    print(spec.name, dataset, "frames", len(idx), "N", n_vertices, "visible", int(np.sum(visible)), "lit", int(np.sum(lit)))
    return {"frame_idx": np.array(idx, np.int32), "sha256": np.stack(digests).reshape(len(idx), n_cam, 32),
            "visible": np.array(visible, np.int64), "cropped": np.array(cropped, np.int64), "lit": np.array(lit, np.int64),
            "n_vertices": np.array(n_vertices, np.int64), "n_instances": np.array(len(cm.instance_maps[dataset]), np.int64)}


CONFIG3_SAMPLE = (0, 1, 31, 64, 97, 130, 159, 160, 201, 255, 288, 319)      # positions in yield order (image_idx - 1)


def fullsize_golden(out_path, root):
    out = {}
    cases = {"config2_nuscenes_exact": (synth.config2_spec(seed=0, name="g_config2_exact"), "nuscenes", None),
             "config2_nuscenes_slerp": (synth.config2_spec(seed=0, pose_time_offset_ms=25, name="g_config2_slerp"), "nuscenes", None),
             "config2_cama_exact": (synth.config2_spec(seed=0, name="g_config2_cama"), "cama", None),
             "config3_nuscenes": (synth.config3_spec(seed=1, name="g_config3"), "nuscenes", set(CONFIG3_SAMPLE))}
    for key, (spec, dataset, sample) in cases.items():
        if dataset == "nuscenes":
            spec.write_cama = False
        else:
            spec.write_nuscenes = False
        for name, value in fullsize_case(spec, dataset, root, sample).items():
            out[f"{key}.{name}"] = value
    np.savez_compressed(out_path, **out)
    print(out_path)


def composited_golden(out_path, root):
    """render_vectors with real camera images (D5: the reference draws in place on the undistort-resized frame)."""
    import cv2
    spec = synth.tiny_spec(name="tiny_composited")
    clip = synth.write_clip(spec, root)
    synth.write_background_jpegs(clip, spec.n_frames, seed=4)
    cm = ClipManager(synth.CAMA_CONFIGS, clip)
    out = {}
    for dataset in ("nuscenes", "cama"):
        idx, digests, raw_digests, keep = [], [], [], {}
        for image_idx, instance_map in cm.yield_frame(dataset):
            maps_2d = cm.project_all_camera(instance_map)
            image_dict = cm.render_vectors(maps_2d, image_idx)
            assert list(image_dict) == synth.CAMERA_LIST
            images = [image_dict[name] for name in synth.CAMERA_LIST]
            idx.append(image_idx)
            digests.append(frame_digests(images))
            raw_digests.append(frame_digests([cv2.imread(cam.get_image_path(image_idx, True)) for cam in cm.cm_list]))
            if image_idx == 1:
                keep = {"image_front": images[1], "image_rear": images[4]}
        out[f"{dataset}.frame_idx"] = np.array(idx, np.int32)
        out[f"{dataset}.sha256"] = np.stack(digests)
        out[f"{dataset}.raw_sha256"] = np.stack(raw_digests)
        for name, image in keep.items():
            out[f"{dataset}.{name}"] = image
        print("composited", dataset, "frames", idx)
    np.savez_compressed(out_path, **out)
    print(out_path)


def main():
    if "--fullsize-only" in sys.argv:
        with tempfile.TemporaryDirectory() as root:
            fullsize_golden(os.path.join(HERE, "golden_fullsize.npz"), root)
            composited_golden(os.path.join(HERE, "golden_composited.npz"), root)
        return
    with tempfile.TemporaryDirectory() as root:
        fullsize_golden(os.path.join(HERE, "golden_fullsize.npz"), root)
        composited_golden(os.path.join(HERE, "golden_composited.npz"), root)
        lidar_golden(os.path.join(HERE, "golden_lidar.npz"), root)
        known_answers(os.path.join(HERE, "golden_known_answers.npz"), root)
        pose_golden(os.path.join(HERE, "golden_pose.npz"))
        for variant, off in (("exact", 0), ("slerp", 25)):
            for dataset in ("nuscenes", "cama"):
                spec = synth.tiny_spec(pose_time_offset_ms=off, name=f"tiny_{variant}")
                clip_golden(spec, dataset, os.path.join(HERE, f"golden_clip_{dataset}_{variant}.npz"), root)


if __name__ == "__main__":
    main()
