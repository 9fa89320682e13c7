"""Pins the oracle (NumPy restatement and plain-C restatement) against golden vectors produced by
running the unmodified reference (tests/golden/make_golden.py).  CPU-only."""
import numpy as np
import pytest

from conftest import load_golden, split_instances
from cama_b200 import synth
from oracle import cama_oracle as orc
from oracle import oracle_c

BOX6 = [orc.CROP_BOX[k] for k in ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")]


def bgr_of(classes):
    out = np.zeros((len(classes), 3), np.uint8)
    for i, c in enumerate(classes):
        rgb = orc.CLASS_RGB["lane_marking" if str(c) == "lane_marking" else "Crosswalk_Line"]
        out[i] = rgb[::-1]
    return out


# ------------------------------------------------------------------ known answers
def test_densify_known_answer():
    g = load_golden("golden_known_answers.npz")
    labels = [{"attrs": {"type": "lane_marking"}, "data": [[0, 0], [1, 0], [1.05, 0], [1.35, 0]]}]
    got = orc.instances_from_metric_labels(labels)[0]["points"]
    assert got.dtype == np.float32 and got.shape == (13, 3)
    assert np.array_equal(got, g["densify_points"])
    c = oracle_c.densify(np.array(labels[0]["data"]))
    assert np.array_equal(c, g["densify_points"][:, :2])
    assert int(np.float32(0.3) / 0.1) == 3      # NEP-50 float32 arithmetic, the reason N differs from a float64 port


def test_stamp_is_clipped_l1_ball():
    g = load_golden("golden_known_answers.npz")
    img = np.zeros((9, 9, 3), np.uint8)
    oracle_c.stamp(img, np.array([[4.0, 4.0]]), [7, 8, 9])
    assert np.array_equal(img, g["stamp_centre"])
    assert int(img.any(-1).sum()) == 13
    img = np.zeros((9, 9, 3), np.uint8)
    oracle_c.stamp(img, np.array([[8.0, 0.0]]), [7, 8, 9])
    oracle_c.stamp(img, np.array([[-1.0, 9.0]]), [1, 2, 3])
    assert np.array_equal(img, g["stamp_clipped"])


def test_config1_anchor():
    g = load_golden("golden_known_answers.npz")
    pts = g["config1_points"]
    K = g["K_scaled_front"]
    assert np.allclose(K, [[759.84, 0, 489.78], [0, 759.84, 294.9], [0, 0, 1]])
    inst = [{"class": "lane_marking", "points": pts}]
    chassis = orc.crop_instances(orc.transform_instances(inst, np.eye(4, dtype=np.float32)))
    assert np.array_equal(chassis[0]["points"], g["config1_crop"]) and len(chassis[0]["points"]) == 46
    E = orc.inv_rigid(synth.camera_to_chassis("camera_front"))
    vu = orc.project_instances(orc.transform_instances(chassis, E), K, 960, 540)
    assert len(vu[0]["points"]) == 44
    assert np.array_equal(vu[0]["points"], g["config1_vu"])
    img = orc.render_instances(np.zeros((540, 960, 3), np.uint8), vu)
    assert np.array_equal(img, g["config1_image"]) and int(img.any(-1).sum()) == 401
    # plain C, same numbers
    c_ch = oracle_c.transform(np.eye(4), pts)
    keep = oracle_c.crop_mask(c_ch, BOX6)
    assert np.array_equal(c_ch[keep], g["config1_crop"])
    c_vu, vis = oracle_c.project(K, oracle_c.transform(E, c_ch[keep]), 960, 540)
    assert np.array_equal(c_vu[vis], g["config1_vu"])
    img = oracle_c.stamp(np.zeros((540, 960, 3), np.uint8), c_vu[vis], [211, 211, 211])
    assert np.array_equal(img, g["config1_image"])


def test_cama_label_load_known_answer():
    g = load_golden("golden_known_answers.npz")
    labels = [{"attrs": {"type": "Road_teeth"}, "data": g["cama_labels_0"].tolist()},
              {"attrs": {"type": "lane_marking"}, "data": [[10.0, 20.0]]},
              {"attrs": {"type": "lane_marking"}, "data": g["cama_labels_2"].tolist()}]
    got = orc.instances_from_pixel_labels(g["cama_bev"], labels)
    assert len(got) == 2 and got[0]["points"].dtype == np.float32
    assert np.array_equal(got[0]["points"], g["cama_points_0"])
    assert np.array_equal(got[1]["points"], g["cama_points_1"])
    for key, ref in (("cama_labels_0", "cama_points_0"), ("cama_labels_2", "cama_points_1")):
        dense = oracle_c.densify(g[key])
        assert np.array_equal(oracle_c.pixel_to_world(dense, g["cama_bev"]), g[ref])


# ------------------------------------------------------------------ pose algebra
def test_pose_golden():
    g = load_golden("golden_pose.npz")
    assert np.array_equal(orc.inv_rigid(g["ext"]), g["invT"])
    stamps, poses = orc.tum_to_poses(g["tum"])
    assert np.array_equal(np.stack(poses), g["abs"])
    assert np.array_equal(orc.slerp_rigid(poses[1], poses[2], 0.3), g["slerp"])
    for mode, interp in (("interp", True), ("nearest", False)):
        for q, want, ok in zip(g["queries"], g[f"seek_{mode}"], g[f"seek_{mode}_ok"]):
            if ok:
                assert np.array_equal(orc.seek_pose(stamps, poses, float(q), 0.5, interp), want)
            else:
                with pytest.raises(RuntimeError):
                    orc.seek_pose(stamps, poses, float(q), 0.5, interp)
    for q, ok in zip(g["queries"], g["seek_interp_tight_ok"]):
        if ok:
            orc.seek_pose(stamps, poses, float(q), 0.1, True)
        else:
            with pytest.raises(RuntimeError):
                orc.seek_pose(stamps, poses, float(q), 0.1, True)
    assert np.array_equal(np.stack([p @ g["ext"] for p in poses]), g["right_rotate"])
    c = orc.inv_rigid(poses[len(poses) // 2])
    assert np.array_equal(np.stack([c @ p for p in poses]), g["normalize2center"])


# ------------------------------------------------------------------ whole-clip golden
@pytest.mark.parametrize("variant,offset", [("exact", 0), ("slerp", 25)])
@pytest.mark.parametrize("dataset", ["nuscenes", "cama"])
def test_clip_golden_numpy_oracle(clip_root, dataset, variant, offset):
    g = load_golden(f"golden_clip_{dataset}_{variant}.npz")
    clip = synth.write_clip(synth.tiny_spec(pose_time_offset_ms=offset, name=f"tiny_{variant}"), clip_root)
    oc = orc.ClipOracle(synth.CAMA_CONFIGS, clip)
    flat, classes, counts = orc.flatten(oc.instance_maps[dataset], 3)
    assert np.array_equal(flat, g["inst_points"]) and flat.dtype == g["inst_points"].dtype
    assert np.array_equal(np.cumsum([0] + counts), g["inst_offsets"])
    assert list(classes) == [str(c) for c in g["inst_classes"]]
    assert np.array_equal(np.stack(oc.K), g["K"])
    assert np.array_equal(np.stack(oc.chassis2cam), g["chassis2camera"])

    per_frame = oc.world_to_chassis_per_frame(dataset)
    assert [i for i, _ in per_frame] == list(g["frame_idx"])
    assert np.array_equal(np.stack([m for _, m in per_frame]), g["world2chassis"])
    assert per_frame[0][1].dtype == np.float32

    crop_pos = vu_pos = 0
    for f, (image_idx, chassis) in enumerate(oc.frames(dataset)):
        cflat, _, ccounts = orc.flatten(chassis, 3)
        assert ccounts == [int(c) for c in g["crop_counts"][f] if c > 0]
        assert np.array_equal(cflat, g["crop_points"][crop_pos:crop_pos + len(cflat)])
        crop_pos += len(cflat)
        per_cam = oc.project_all(chassis)
        for c, cam in enumerate(oc.cameras):
            vflat, _, vcounts = orc.flatten(per_cam[cam], 2)
            assert vcounts == [int(k) for k in g["vu_counts"][f, c] if k > 0]
            assert np.array_equal(vflat, g["vu_points"][vu_pos:vu_pos + len(vflat)])
            vu_pos += len(vflat)
            img = orc.render_instances(np.zeros((540, 960, 3), np.uint8), per_cam[cam])
            assert np.array_equal(img, g["frames"][f, c])
    assert crop_pos == len(g["crop_points"]) and vu_pos == len(g["vu_points"])


@pytest.mark.parametrize("variant", ["exact", "slerp"])
@pytest.mark.parametrize("dataset", ["nuscenes", "cama"])
def test_clip_golden_c_oracle(dataset, variant):
    """Plain-C restatement from the fixture's own inputs: frames and counts bit-exact; the float
    stages bit-exact wherever an instance has >= 2 points (see the gemv note in cama_oracle.py)."""
    g = load_golden(f"golden_clip_{dataset}_{variant}.npz")
    frames, crop_counts, vis_counts = oracle_c.clip_render(
        g["inst_points"], g["inst_offsets"], bgr_of(g["inst_classes"]), g["world2chassis"],
        g["chassis2camera"], g["K"], BOX6, 540, 960)
    assert np.array_equal(crop_counts, g["crop_counts"])
    assert np.array_equal(vis_counts, g["vu_counts"])
    assert np.array_equal(frames, g["frames"])
    # stage-by-stage floats
    crop_pos = vu_pos = 0
    offs = g["inst_offsets"]
    for f in range(len(g["frame_idx"])):
        T = g["world2chassis"][f].astype(np.float64)
        chassis = oracle_c.transform(T, g["inst_points"])
        keep = oracle_c.crop_mask(chassis, BOX6)
        n_keep = int(keep.sum())
        want = g["crop_points"][crop_pos:crop_pos + n_keep]
        got = chassis[keep]
        single = np.repeat(np.diff(offs) == 1, np.diff(offs))[keep]
        assert np.array_equal(got[~single], want[~single])
        assert np.allclose(got, want, rtol=0, atol=1e-9)
        crop_pos += n_keep
        for c in range(g["K"].shape[0]):
            vu, vis = oracle_c.project(g["K"][c], oracle_c.transform(g["chassis2camera"][c], want), 960, 540)
            n_vis = int(vis.sum())
            ref = g["vu_points"][vu_pos:vu_pos + n_vis]
            multi = np.repeat(g["crop_counts"][f] > 1, g["crop_counts"][f])[vis]
            assert np.array_equal(vu[vis][multi], ref[multi])
            assert np.allclose(vu[vis], ref, rtol=0, atol=1e-9)
            vu_pos += n_vis
    assert crop_pos == len(g["crop_points"]) and vu_pos == len(g["vu_points"])
