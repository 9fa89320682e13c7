"""Host-side logic of the package (no GPU): pose algebra, clip reading, load-time densify and the
per-frame pose chain, all against golden vectors produced by the unmodified reference."""
import sys

import numpy as np
import pytest

from conftest import load_golden
from cama_b200 import synth
from cama_b200.dataset import ClipManager
from cama_b200.dataset_reader import DatasetReader
from cama_b200.pose_transformer import PoseTransformer, SlerpTransform, invT
from cama_b200.reproject import CameraManager, MapManager, densify_polyline, pack_instances, unpack_instances


def test_pose_transformer_full_surface():
    g = load_golden("golden_pose.npz")
    tum, ext = g["tum"], g["ext"]
    assert np.array_equal(invT(ext), g["invT"])
    pt = PoseTransformer()
    pt.loadarray(tum)
    assert np.array_equal(pt.as_transform(True), g["abs"])
    assert np.array_equal(pt.as_transform(False), g["rel"])
    assert np.array_equal(np.asarray(pt.as_quaternions()), g["quats"])
    assert np.array_equal(pt.as_euler(True), g["euler_abs"])
    assert np.array_equal(pt.as_euler(False), g["euler_rel"])
    assert np.array_equal(pt.as_axis_angle(True), g["axis_abs"])
    assert np.array_equal(pt.as_axis_angle(False), g["axis_rel"])
    with pytest.warns(UserWarning):
        assert np.array_equal(pt.as_axisangle(True), g["axis_abs"])
    assert np.array_equal(pt.as_translations(True), g["trans_abs"])
    assert np.array_equal(pt.as_translations(False), g["trans_rel"])
    assert np.array_equal(pt.as_trans_quat(), g["trans_quat"])
    assert np.array_equal(pt.dumparray(), g["dump_tum"])
    assert np.array_equal(pt.get_timestamps(), tum[:, 0:1])
    assert np.array_equal(SlerpTransform(g["abs"][1], g["abs"][2], 0.3), g["slerp"])
    with pytest.raises(NotImplementedError):
        pt.dumparray(style="kitti")
    with pytest.raises(NotImplementedError):
        pt.as_quaternions(absolute=False)


def test_seek_by_timestamp_matches_reference():
    g = load_golden("golden_pose.npz")
    pt = PoseTransformer()
    pt.loadarray(g["tum"])
    for mode, interp in (("interp", True), ("nearest", False)):
        for q, want, ok in zip(g["queries"], g[f"seek_{mode}"], g[f"seek_{mode}_ok"]):
            if ok:
                assert np.array_equal(np.asarray(pt.seek_by_timestamp(float(q), 0.5, interpolate=interp)), want)
            else:
                with pytest.raises(RuntimeError):
                    pt.seek_by_timestamp(float(q), 0.5, interpolate=interp)
    for q, ok in zip(g["queries"], g["seek_interp_tight_ok"]):
        if ok:
            pt.seek_by_timestamp(float(q), 0.1, interpolate=True)
        else:
            with pytest.raises(RuntimeError):
                pt.seek_by_timestamp(float(q), 0.1, interpolate=True)
    with pytest.raises(AssertionError):
        pt.seek_by_timestamp(np.float32(1.0), 0.5)          # must be a python float
    with pytest.raises(AssertionError):
        pt.seek_by_timestamp(1.0, 1)
    with pytest.raises(RuntimeError):
        PoseTransformer().seek_by_timestamp(1.0, 0.5)


@pytest.mark.parametrize("name", ["right_rotate", "left_rotate", "transform", "normalize2center", "normalize2origin"])
def test_trajectory_edits(name):
    g = load_golden("golden_pose.npz")
    pt = PoseTransformer()
    pt.loadarray(g["tum"])
    getattr(pt, name)(*(() if name.startswith("normalize") else (g["ext"],)))
    assert np.array_equal(pt.as_transform(True), g[name])
    if name == "right_rotate":
        p2 = PoseTransformer()
        p2.loadarray(g["tum"])
        with pytest.warns(UserWarning):
            p2.rotate(g["ext"])
        assert np.array_equal(p2.as_transform(True), g[name])


def test_other_loaders():
    g = load_golden("golden_pose.npz")
    n = len(g["tum"])
    p = PoseTransformer()
    p.loadarray(g["abs"][:, :3, :].reshape(n, 12), style="kitti")
    assert np.array_equal(p.as_transform(True), g["kitti_abs"]) and np.array_equal(p.as_transform(False), g["kitti_rel"])
    p = PoseTransformer()
    p.loadarray(g["asl"], style="asl")
    assert np.array_equal(p.as_transform(True), g["asl_abs"]) and np.array_equal(p.get_timestamps(), g["asl_ts"])
    with pytest.raises(NotImplementedError):
        p.loadarray(g["asl"], style="euroc")
    p = PoseTransformer()
    p.from_relative_eulers(g["rel_eulers"])
    p.from_translation(g["rel_trans"], absolute=False)
    assert np.array_equal(p.as_transform(True), g["from_rel_abs"])
    p = PoseTransformer()
    p.from_axis_angle(g["rel_eulers"], absolute=True)
    p.from_translation(g["rel_trans"], absolute=True)
    assert np.array_equal(p.as_transform(True), g["from_abs_axis_abs"])
    p = PoseTransformer()
    p.from_relative_quaternion(g["tum"][:5, 4:8])
    p.from_relative_translation(g["rel_trans"])
    assert np.array_equal(p.as_transform(True), g["from_rel_quat_abs"])
    p = PoseTransformer()
    p.from_relative_axis_angle(g["rel_eulers"])
    p.from_relative_translation(g["rel_trans"])
    p.load_timestamp(list(g["tum"][:5, 0][::-1]))
    p.sort_by_timestamps()
    assert np.array_equal(p.as_transform(False), g["sorted_rel"]) and np.array_equal(p.get_timestamps(), g["sorted_ts"])
    with pytest.raises(AssertionError):
        PoseTransformer().from_relative_axis_angle(np.zeros((3, 4)))


def test_densify_and_label_loading_known_answers():
    g = load_golden("golden_known_answers.npz")
    mm = MapManager()
    labels = [{"attrs": {"type": "lane_marking"}, "data": [[0, 0], [1, 0], [1.05, 0], [1.35, 0]]},
              {"attrs": {"type": "lane_marking"}, "data": [[3, 3]]}]
    out = mm.load_3d_instance_maps(labels)
    assert len(out) == 1 and out[0]["points"].dtype == np.float32
    assert np.array_equal(out[0]["points"], g["densify_points"])
    with pytest.raises(IndexError):       # every segment shorter than one step: the reference raises IndexError too
        mm.load_3d_instance_maps([{"attrs": {"type": "x"}, "data": [[0, 0], [0.01, 0]]}])
    labels = [{"attrs": {"type": "Road_teeth"}, "data": g["cama_labels_0"].tolist()},
              {"attrs": {"type": "lane_marking"}, "data": [[10.0, 20.0]]},
              {"attrs": {"type": "lane_marking"}, "data": g["cama_labels_2"].tolist()}]
    got = mm.calculate_3d_instance_maps(g["cama_bev"], labels)
    assert [i["class"] for i in got] == ["Road_teeth", "lane_marking"]
    assert np.array_equal(got[0]["points"], g["cama_points_0"]) and np.array_equal(got[1]["points"], g["cama_points_1"])
    assert got[0]["points"].dtype == np.float32
    # float64 height map promotes the instance to float64, as np.concatenate does in the reference
    got64 = mm.calculate_3d_instance_maps(g["cama_bev"].astype(np.float64), labels)
    assert got64[0]["points"].dtype == np.float64 and np.array_equal(got64[0]["points"], g["cama_points_0"].astype(np.float64))


def test_pack_unpack_roundtrip():
    inst = [{"class": "a", "points": np.ones((3, 3), np.float32)}, {"class": "b", "points": np.zeros((0, 3), np.float32)},
            {"class": "c", "points": np.full((2, 3), 2, np.float32)}]
    flat, off, cls = pack_instances(inst)
    assert flat.shape == (5, 3) and list(off) == [0, 3, 3, 5] and cls == ["a", "b", "c"]
    assert [i["class"] for i in unpack_instances(flat, off, cls)] == ["a", "c"]
    assert len(unpack_instances(flat, off, cls, drop_empty=False)) == 3
    dense = densify_polyline([[0, 0], [0.35, 0]], 0.1)
    assert dense.dtype == np.float32 and dense.shape == (3, 2) and dense[0, 0] == 0


@pytest.mark.parametrize("variant,offset", [("exact", 0), ("slerp", 25)])
@pytest.mark.parametrize("dataset", ["nuscenes", "cama"])
def test_clip_manager_host_side(clip_root, dataset, variant, offset):
    """Loading, calibration and the per-frame pose chain of ClipManager vs the reference's."""
    g = load_golden(f"golden_clip_{dataset}_{variant}.npz")
    clip = synth.write_clip(synth.tiny_spec(pose_time_offset_ms=offset, name=f"tiny_{variant}"), clip_root)
    cm = ClipManager(synth.CAMA_CONFIGS, clip)
    assert set(cm.instance_maps) == {"cama", "nuscenes"}
    flat, offs, classes = pack_instances(cm.instance_maps[dataset])
    assert flat.dtype == np.float32 and np.array_equal(flat, g["inst_points"])
    assert np.array_equal(offs, g["inst_offsets"]) and classes == [str(c) for c in g["inst_classes"]]
    assert [c.camera_name for c in cm.cm_list] == synth.CAMERA_LIST
    assert np.array_equal(np.stack([c.K for c in cm.cm_list]), g["K"])
    assert np.array_equal(np.stack([c.get_chassis2camera() for c in cm.cm_list]), g["chassis2camera"])
    poses = cm.frame_poses(dataset)
    assert [i for i, _ in poses] == list(g["frame_idx"])
    w2c = np.stack([m for _, m in poses])
    assert w2c.dtype == np.float32 and np.array_equal(w2c, g["world2chassis"])


def test_frames_without_pose_are_skipped(clip_root):
    """cama/dataset.py:90-96: a RuntimeError from the pose lookup silently drops the frame."""
    spec = synth.tiny_spec(n_frames=6, name="tiny_gap")
    clip = synth.write_clip(spec, clip_root)
    import os
    path = os.path.join(clip, "odometry", "wigo_offset_clip.txt")
    rows = np.loadtxt(path)
    np.savetxt(path, rows[:5], fmt="%.12f")         # poses stop before the last two frames
    cm = ClipManager(synth.CAMA_CONFIGS, clip)
    assert [i for i, _ in cm.frame_poses("nuscenes")] == [1, 2, 3, 4]
    assert [i for i, _ in cm.frame_poses("cama")] == [1, 2, 3, 4, 5, 6]


def test_missing_labels_and_missing_clip(clip_root, tmp_path):
    spec = synth.tiny_spec(name="tiny_nocama")
    spec.write_cama = False
    clip = synth.write_clip(spec, str(tmp_path))
    cm = ClipManager(synth.CAMA_CONFIGS, clip)
    assert set(cm.instance_maps) == {"nuscenes"}
    with pytest.raises(KeyError):
        next(cm.yield_frame("cama"))
    with pytest.raises(FileNotFoundError):
        DatasetReader(str(tmp_path / "nope"))


def test_dataset_reader_extrinsic_chain(tmp_path):
    import json
    a2b = np.eye(4); a2b[:3, 3] = [1, 2, 3]
    b2c = np.eye(4); b2c[:3, :3] = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]; b2c[:3, 3] = [0.5, 0, 0]
    attr = {"calibration": {"a_2_b": a2b.tolist(), "b_2_c": b2c.tolist(), "cam": {"K": np.eye(3).tolist(), "d": [0.0] * 8,
                                                                                  "image_width": 10, "image_height": 5}},
            "sync": {"cam": [1000, 1500]}, "unsync": {"cam": [1000, 1250, 1500]}}
    (tmp_path / "attribute.json").write_text(json.dumps(attr))
    dr = DatasetReader(str(tmp_path))
    assert dr.get_extrinsic("a", "a").dtype == np.float32
    assert np.array_equal(dr.get_extrinsic("a", "b"), a2b)
    assert np.array_equal(dr.get_extrinsic("b", "a"), invT(a2b))
    assert dr.get_extrinsic_path("a", "c") == ["a", "b", "c"]
    assert np.array_equal(dr.get_extrinsic("a", "c"), b2c @ a2b @ np.eye(4, dtype=np.float32))
    assert np.array_equal(dr.get_extrinsic("c", "a"), invT(a2b) @ invT(b2c) @ np.eye(4, dtype=np.float32))
    assert dr.get_extrinsic("a", "zzz") is None
    assert dr.get_sensor_timestamp("cam") == [1.0, 1.5] and dr.get_sensor_timestamp("cam", sync=False) == [1.0, 1.25, 1.5]
    info = dr.get_intrinsics("cam")
    assert info["width"] == 10 and info["height"] == 5 and info["hfov"] is None
    assert sorted(dr.get_all_sensors()) == ["a", "b", "c", "cam"]
    assert list(dr.yield_sensor_filepath("cam", "jpg"))[0].endswith("cam/1000.jpg")


def test_install_as_cama_alias():
    import cama_b200
    saved = {k: v for k, v in sys.modules.items() if k == "cama" or k.startswith("cama.")}
    for k in saved:
        del sys.modules[k]
    try:
        cama_b200.install_as_cama()
        from cama.dataset import ClipManager as CM          # the imports main.py does
        from cama.tools import load_json                   # noqa: F401
        from cama.pose_transformer import invT as inv2
        assert CM is ClipManager and inv2 is invT
    finally:
        for k in [k for k in sys.modules if k == "cama" or k.startswith("cama.")]:
            del sys.modules[k]
        sys.modules.update(saved)
