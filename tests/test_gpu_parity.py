"""GPU parity: the sm_100a kernels, called through the C ABI (ctypes -> libcama_b200.so), against
the reference-generated golden fixtures and the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): rendered frames, counts and pixel centres bit-exact; projected
(v,u) float coordinates within 1e-5 (they are in fact bit-exact except where NumPy's one-column
gemv path differs from dgemm by an ulp, see oracle/cama_oracle.py).
Nothing here reads /root/reference.
"""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, split_instances
from cama_b200 import synth
from oracle import cama_oracle as orc
from oracle import oracle_c

pytestmark = pytest.mark.gpu

COORD_TOL = 1e-5
BOX6 = [orc.CROP_BOX[k] for k in ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")]
H, W = 540, 960


@pytest.fixture(scope="module")
def rt():
    import torch
    assert torch.cuda.is_available(), "the -m gpu tests need a CUDA device"
    from cama_b200.runtime import get_runtime
    runtime = get_runtime(0)
    from cama_b200 import _native as N
    assert N.lib().cama_abi_version() == N.ABI_VERSION
    return runtime


def bgr_of(classes):
    out = np.zeros((len(classes), 3), np.uint8)
    for i, c in enumerate(classes):
        out[i] = orc.CLASS_RGB["lane_marking" if str(c) == "lane_marking" else "Crosswalk_Line"][::-1]
    return out


def golden_instances(g):
    offs = g["inst_offsets"]
    return [{"class": str(c), "points": g["inst_points"][offs[i]:offs[i + 1]]} for i, c in enumerate(g["inst_classes"])]


def renderer_for(g, device=0):
    from cama_b200.batched import ClipRenderer
    return ClipRenderer(g["chassis2camera"], g["K"], H, W, BOX6, device=device)


def to_dev(a, dtype=None):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


# ------------------------------------------------------------------ per-call operators vs golden
@pytest.mark.parametrize("variant", ["exact", "slerp"])
@pytest.mark.parametrize("dataset", ["nuscenes", "cama"])
def test_operators_match_reference_outputs(rt, dataset, variant):
    """MapManager / CameraManager methods (one C-ABI operator each) on the golden clip."""
    from cama_b200.reproject import CameraManager, MapManager
    g = load_golden(f"golden_clip_{dataset}_{variant}.npz")
    instances = golden_instances(g)
    mm = MapManager(device=0)
    cams = []
    for c in range(g["K"].shape[0]):
        cm = CameraManager.__new__(CameraManager)          # calibration comes from the fixture, not from a clip dir
        cm.K, cm.chassis2camera, cm.width, cm.height, cm._device = g["K"][c], g["chassis2camera"][c], W, H, 0
        cm.camera_name = f"cam{c}"
        cams.append(cm)
    crop_pos = vu_pos = 0
    for f in range(len(g["frame_idx"])):
        w2c = g["world2chassis"][f]
        # R9 then R10, as two calls (reference cama/dataset.py:99-105) ...
        moved = mm.transform_3d_instance_maps(instances, w2c)
        assert len(moved) == len(instances) and moved[0]["points"].dtype == np.float64
        chassis = mm.crop_3d_instance_maps(moved)
        # ... and fused
        fused = mm.transform_crop_3d_instance_maps(instances, w2c)
        assert [i["class"] for i in chassis] == [i["class"] for i in fused]
        for a, b in zip(chassis, fused):
            assert np.array_equal(a["points"], b["points"])
        flat, _, counts = orc.flatten(chassis, 3)
        assert counts == [int(c) for c in g["crop_counts"][f] if c > 0]
        want = g["crop_points"][crop_pos:crop_pos + len(flat)]
        multi = np.repeat(np.array(counts) > 1, counts)
        assert np.array_equal(flat[multi], want[multi])
        assert np.abs(flat - want).max(initial=0.0) <= COORD_TOL
        crop_pos += len(flat)
        # feed the REFERENCE's chassis points onward so stages are checked independently
        ref_chassis = split_instances(want, g["crop_counts"][f], g["inst_classes"])
        for c, cm in enumerate(cams):
            in_cam = mm.transform_3d_instance_maps(ref_chassis, cm.get_chassis2camera())
            vu = cm.project_to_image(in_cam)
            vu_fused = cm.transform_project_to_image(ref_chassis)
            vflat, vclasses, vcounts = orc.flatten(vu, 2)
            fflat, _, fcounts = orc.flatten(vu_fused, 2)
            assert vcounts == fcounts == [int(k) for k in g["vu_counts"][f, c] if k > 0]
            assert np.array_equal(vflat, fflat)
            ref = g["vu_points"][vu_pos:vu_pos + len(vflat)]
            assert np.abs(vflat - ref).max(initial=0.0) <= COORD_TOL
            assert np.array_equal(vflat.astype(np.int32), ref.astype(np.int32))
            vu_pos += len(vflat)
            img = cm.render_maps(np.zeros((H, W, 3), np.uint8), split_instances(ref, g["vu_counts"][f, c], g["inst_classes"]))
            assert np.array_equal(img, g["frames"][f, c])
    assert crop_pos == len(g["crop_points"]) and vu_pos == len(g["vu_points"])


def test_config1_anchor(rt):
    """50-vertex lane polyline, CAM_FRONT: 46 in the box, 44 visible, 401 px lit."""
    from cama_b200.batched import ClipRenderer
    g = load_golden("golden_known_answers.npz")
    E = orc.inv_rigid(synth.camera_to_chassis("camera_front"))
    r = ClipRenderer(E[None], g["K_scaled_front"][None], H, W, BOX6, device=0)
    res = r.resident([{"class": "lane_marking", "points": g["config1_points"].astype(np.float32)}])
    frames, dbg = r.render(res, to_dev(np.eye(4, dtype=np.float32).reshape(1, 16)), debug=True, want_vu=True)
    assert int(dbg["crop_counts"].sum()) == 46 and int(dbg["visible_counts"].sum()) == 44
    img = frames[0, 0].cpu().numpy()
    assert np.array_equal(img, g["config1_image"]) and int(img.any(-1).sum()) == 401
    vu = dbg["vu_dense"][0, 0].cpu().numpy()
    vu = vu[~np.isnan(vu[:, 0])]
    assert np.array_equal(vu, g["config1_vu"])


def test_render_points_clips_like_cv2_circle(rt):
    g = load_golden("golden_known_answers.npz")
    img = rt.render_points(np.zeros((9, 9, 3), np.uint8), np.array([[4.0, 4.0]]), np.array([0, 1]), np.array([[7, 8, 9]], np.uint8))
    assert np.array_equal(img, g["stamp_centre"])
    img = np.zeros((9, 9, 3), np.uint8)
    rt.render_points(img, np.array([[8.0, 0.0], [-1.0, 9.0]]), np.array([0, 1, 2]), np.array([[7, 8, 9], [1, 2, 3]], np.uint8))
    assert np.array_equal(img, g["stamp_clipped"])
    # NaN / huge centres are ignored like INT32_MIN centres are by cv2
    img = rt.render_points(np.zeros((9, 9, 3), np.uint8), np.array([[np.nan, 1.0], [1e300, 2.0], [3.9, 3.9]]), np.array([0, 3]),
                           np.array([[5, 5, 5]], np.uint8))
    want = oracle_c.stamp(np.zeros((9, 9, 3), np.uint8), np.array([[3.9, 3.9]]), [5, 5, 5])
    assert np.array_equal(img, want)


# ------------------------------------------------------------------ batched clip path vs golden
@pytest.mark.parametrize("mode", ["binned", "plane"])
@pytest.mark.parametrize("variant", ["exact", "slerp"])
@pytest.mark.parametrize("dataset", ["nuscenes", "cama"])
def test_clip_render_matches_reference_outputs(rt, dataset, variant, mode):
    g = load_golden(f"golden_clip_{dataset}_{variant}.npz")
    r = renderer_for(g)
    res = r.resident(golden_instances(g))
    frames, dbg = r.render(res, to_dev(g["world2chassis"].reshape(-1, 16)), mode=mode, debug=True, want_vu=True)
    assert np.array_equal(dbg["crop_counts"].cpu().numpy(), g["crop_counts"])
    assert np.array_equal(dbg["visible_counts"].cpu().numpy(), g["vu_counts"])
    assert np.array_equal(frames.cpu().numpy(), g["frames"])
    vu = dbg["vu_dense"].cpu().numpy().reshape(-1, 2)      # (frame, camera, vertex) order == the reference's loop order
    vu = vu[~np.isnan(vu[:, 0])]
    assert vu.shape == g["vu_points"].shape
    assert np.abs(vu - g["vu_points"]).max() <= COORD_TOL
    assert np.array_equal(vu.astype(np.int32), g["vu_points"].astype(np.int32))
    single = np.repeat(g["crop_counts"].reshape(-1) == 1, g["vu_counts"].sum(axis=1).reshape(-1))
    assert np.array_equal(vu[~single], g["vu_points"][~single])
    if mode == "binned":
        assert r.last_stats["mode"] == 2 and r.last_stats["overflow"] == 0


@pytest.mark.parametrize("variant,offset", [("exact", 0), ("slerp", 25)])
@pytest.mark.parametrize("dataset", ["nuscenes", "cama"])
def test_drop_in_frame_loop(rt, clip_root, dataset, variant, offset):
    """The three calls of the reference's main.py:57-59 on a clip directory, then the batched
    Reproject façade on the same clip; both must equal what the reference produced."""
    from cama_b200.batched import Reproject
    from cama_b200.dataset import ClipManager
    g = load_golden(f"golden_clip_{dataset}_{variant}.npz")
    clip = synth.write_clip(synth.tiny_spec(pose_time_offset_ms=offset, name=f"tiny_gpu_{variant}"), clip_root)
    cm = ClipManager(synth.CAMA_CONFIGS, clip, device=0, progress=False)
    seen = []
    for f, (image_idx, instance_map) in enumerate(cm.yield_frame(dataset)):
        maps_2d = cm.project_all_camera(instance_map)
        assert list(maps_2d) == synth.CAMERA_LIST
        for c, cam in enumerate(cm.cm_list):
            img = np.zeros((H, W, 3), np.uint8)
            out = cam.render_maps(img, maps_2d[cam.camera_name])
            assert out is img                               # drawn in place, like cv2.circle
            assert np.array_equal(img, g["frames"][f, c])
        seen.append(image_idx)
    assert seen == list(g["frame_idx"])
    idx, frames = Reproject(synth.CAMA_CONFIGS, clip_manager=cm, device=0)(dataset)
    assert idx == list(g["frame_idx"]) and np.array_equal(frames, g["frames"])
    idx2, frames2 = cm.render_clip(dataset, mode="plane")
    assert idx2 == idx and np.array_equal(frames2, g["frames"])


def test_backgrounds_are_composited_in_place(rt):
    g = load_golden("golden_clip_nuscenes_exact.npz")
    r = renderer_for(g)
    res = r.resident(golden_instances(g))
    rng = np.random.default_rng(3)
    bg = rng.integers(0, 256, size=g["frames"].shape, dtype=np.uint8)
    want, _, _ = oracle_c.clip_render(g["inst_points"], g["inst_offsets"], bgr_of(g["inst_classes"]), g["world2chassis"],
                                      g["chassis2camera"], g["K"], BOX6, H, W, frames=bg.copy())
    w2c = to_dev(g["world2chassis"].reshape(-1, 16))
    for mode in ("binned", "plane"):
        d_bg = to_dev(bg)
        out = r.render(res, w2c, background=d_bg, mode=mode)                  # separate output
        assert np.array_equal(out.cpu().numpy(), want)
        out = r.render(res, w2c, out=d_bg, background=d_bg, mode=mode)          # in place
        assert out.data_ptr() == d_bg.data_ptr() and np.array_equal(d_bg.cpu().numpy(), want)


# ------------------------------------------------------------------ raster stress: painter's order, borders, band seams
def _pixel_clip(rng, n_inst, n_pts, width, height, margin=4.0):
    """Vertices that project (identity poses, K = I) to uniformly random pixel positions,
    including a margin outside the image, on band seams and on the exact borders."""
    counts = rng.multinomial(n_pts, np.ones(n_inst) / n_inst)
    u = rng.uniform(-margin, width + margin, n_pts)
    v = rng.uniform(-margin, height + margin, n_pts)
    special = rng.random(n_pts) < 0.05
    u[special] = rng.choice([0.0, 0.999, width - 1.0, width - 0.001, float(width), -0.0], special.sum())
    v[special] = rng.choice([0.0, 1.0, 2.0, height - 1.0, float(height), height - 0.5], special.sum())
    z = rng.uniform(0.5, 40.0, n_pts)
    pts = np.stack([u * z, v * z, z], axis=1).astype(np.float32)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return pts, offs


@pytest.mark.parametrize("width,height", [(960, 540), (64, 48), (1024, 96), (100, 37)])
def test_raster_stress_vs_oracle(rt, width, height):
    from cama_b200.batched import ClipRenderer
    rng = np.random.default_rng(width * 7 + height)
    n_inst, n_pts, n_frames = 37, 60000 if width * height > 10000 else 3000, 2
    pts, offs = _pixel_clip(rng, n_inst, n_pts, width, height)
    classes = [synth.MAP_CLASSES[i % 3] for i in range(n_inst)]
    eye = np.eye(4)
    box = [-1e9, 1e9, -1e9, 1e9, -1e9, 1e9]
    K2 = np.eye(3)
    r = ClipRenderer(np.stack([eye, eye]), np.stack([K2, K2]), height, width, box, device=0)
    w2c = np.stack([np.eye(4, dtype=np.float32)] * n_frames)
    w2c[1, 0, 3] = 0.25                                     # second frame: shifted a little
    want, cc, vc = oracle_c.clip_render(pts, offs, bgr_of(classes), w2c, np.stack([eye, eye]), np.stack([K2, K2]), box, height, width)
    res = r.resident([{"class": classes[i], "points": pts[offs[i]:offs[i + 1]]} for i in range(n_inst)])
    modes = ["plane"] + (["binned"] if width % 16 == 0 else [])
    for mode in modes:
        frames, dbg = r.render(res, to_dev(w2c.reshape(-1, 16)), mode=mode, debug=True)
        assert np.array_equal(dbg["crop_counts"].cpu().numpy(), cc)
        assert np.array_equal(dbg["visible_counts"].cpu().numpy(), vc)
        got = frames.cpu().numpy()
        assert np.array_equal(got, want), f"{mode}: {(got != want).any(-1).sum()} pixels differ"
    if width % 16:
        from cama_b200 import _native as N
        with pytest.raises(N.CamaError):
            r.render(res, to_dev(w2c.reshape(-1, 16)), mode="binned")
        auto = r.render(res, to_dev(w2c.reshape(-1, 16)), mode="auto")       # AUTO falls back to PLANE
        assert np.array_equal(auto.cpu().numpy(), want) and r.last_stats["mode"] == 1


def test_degenerate_points_are_masked_like_numpy(rt):
    """z == 0 (inf/nan after the divide), negative z, NaN/inf vertices, points exactly on the crop
    box faces (inclusive) and on the image borders (u == W excluded, u == 0 included)."""
    from cama_b200.batched import ClipRenderer
    pts = np.array([[10, 10, 0], [0, 0, 0], [10, 10, -1], [np.nan, 1, 1], [np.inf, 1, 1], [1, 1, np.inf],
                    [0, 0, 1], [64, 10, 1], [63.999, 47.999, 1], [10, 48, 1], [50, 20, 1], [50.001, 20, 1],
                    [-0.0, 5, 1], [-1e-300, 5, 1], [30, 30, 1e-300], [5, 5, 200], [5, 5, 200.1]], dtype=np.float32)
    offs = np.array([0, 5, 12, len(pts)], np.int64)
    classes = ["lane_marking", "Road_teeth", "lane_marking"]
    box = [-50, 50, -100, 100, -200, 200]
    eye, K = np.eye(4), np.eye(3)
    r = ClipRenderer(eye[None], K[None], 48, 64, box, device=0)
    w2c = np.eye(4, dtype=np.float32).reshape(1, 16)
    with np.errstate(all="ignore"):
        want, cc, vc = oracle_c.clip_render(pts, offs, bgr_of(classes), w2c, eye[None], K[None], box, 48, 64)
        # the NumPy restatement agrees with the C one on every mask
        chassis = orc.crop_instances(orc.transform_instances([{"class": c, "points": pts[offs[i]:offs[i + 1]]} for i, c in enumerate(classes)],
                                                             np.eye(4, dtype=np.float32)))
        vu = orc.project_instances(orc.transform_instances(chassis, eye), K, 64, 48)
    assert sum(len(i["points"]) for i in vu) == int(vc.sum())
    res = r.resident([{"class": c, "points": pts[offs[i]:offs[i + 1]]} for i, c in enumerate(classes)])
    for mode in ("binned", "plane"):
        frames, dbg = r.render(res, to_dev(w2c), mode=mode, debug=True)
        assert np.array_equal(dbg["crop_counts"].cpu().numpy(), cc)
        assert np.array_equal(dbg["visible_counts"].cpu().numpy(), vc)
        assert np.array_equal(frames.cpu().numpy(), want)


def test_empty_and_ragged_inputs(rt):
    import torch
    from cama_b200.batched import ClipRenderer
    from cama_b200.reproject import CameraManager, MapManager
    g = load_golden("golden_clip_nuscenes_exact.npz")
    r = renderer_for(g)
    w2c = to_dev(g["world2chassis"].reshape(-1, 16))
    # no instances at all: black frames (or the untouched background)
    res = r.resident([])
    for mode in ("binned", "plane"):
        out = r.render(res, w2c, mode=mode)
        assert out.shape == (3, 6, H, W, 3) and int(out.max()) == 0
    # zero frames
    out = r.render(r.resident(golden_instances(g)), torch.empty((0, 16), dtype=torch.float32, device="cuda"))
    assert tuple(out.shape) == (0, 6, H, W, 3)
    # instances with zero points in the middle of the list keep ordinals aligned
    inst = golden_instances(g)
    ragged = [inst[0], {"class": "Road_teeth", "points": np.zeros((0, 3), np.float32)}] + inst[1:]
    out = r.render(r.resident(ragged), w2c)
    assert np.array_equal(out.cpu().numpy(), g["frames"])
    # per-call operators on empty lists / all-cropped input
    mm = MapManager(device=0)
    assert mm.transform_3d_instance_maps([], np.eye(4)) == [] and mm.crop_3d_instance_maps([]) == []
    far = [{"class": "lane_marking", "points": np.full((5, 3), 1e4)}]
    assert mm.crop_3d_instance_maps(far) == []
    cm = CameraManager.__new__(CameraManager)
    cm.K, cm.chassis2camera, cm.width, cm.height, cm._device = g["K"][0], g["chassis2camera"][0], W, H, 0
    assert cm.project_to_image([]) == [] and cm.project_to_image(far[:0]) == []
    behind = [{"class": "lane_marking", "points": np.array([[0.0, 0.0, -5.0], [1.0, 1.0, 0.0]])}]
    assert cm.project_to_image(behind) == []
    img = np.full((H, W, 3), 9, np.uint8)
    assert cm.render_maps(img, []) is img and int(img.min()) == 9


def test_record_pool_overflow_is_detected_and_retried(rt):
    from cama_b200 import _native as N
    g = load_golden("golden_clip_cama_exact.npz")
    r = renderer_for(g)
    res = r.resident(golden_instances(g))
    w2c = to_dev(g["world2chassis"].reshape(-1, 16))
    # 1. a deliberately tiny pool: the raw ABI call reports CAMA_E_CAPACITY with the needed size
    out = to_dev(np.zeros(g["frames"].shape, np.uint8))
    desc = r._desc(res, w2c, 3, out, None, "binned", 64, None)
    need = ctypes.c_size_t()
    N.check(N.lib().cama_clip_workspace_bytes(ctypes.byref(desc), ctypes.byref(need)))
    ws = r.rt.scratch("overflow-test", need.value)
    N.check(N.lib().cama_clip_render(r.rt.ctx, ctypes.byref(desc), r.rt.ptr(ws), ws.numel(), r.rt.stream()))
    stats = N.ClipStats()
    code = N.lib().cama_clip_stats_read(r.rt.ctx, ctypes.byref(desc), r.rt.ptr(ws), r.rt.stream(), ctypes.byref(stats))
    assert code == N.CAMA_E_CAPACITY and stats.overflow == 1 and stats.record_capacity_needed > 64
    # 2. the Python wrapper reruns with the reported capacity and gets the right frames
    r.capacity[(id(res), 3)] = 64
    frames = r.render(res, w2c, mode="binned")
    assert np.array_equal(frames.cpu().numpy(), g["frames"])
    assert r.last_stats["overflow"] == 0 and r.last_stats["record_capacity"] >= stats.record_capacity_needed


def test_abi_argument_errors(rt):
    from cama_b200 import _native as N
    L = N.lib()
    assert L.cama_transform_points(rt.ctx, None, 1, 5, None, None, None) == N.CAMA_E_INVALID
    need = ctypes.c_size_t()
    N.check(L.cama_render_workspace_bytes(H, W, ctypes.byref(need)))
    g = load_golden("golden_clip_nuscenes_exact.npz")
    r = renderer_for(g)
    res = r.resident(golden_instances(g))
    w2c = to_dev(g["world2chassis"].reshape(-1, 16))
    out = to_dev(np.zeros(g["frames"].shape, np.uint8))
    desc = r._desc(res, w2c, 3, out, None, "auto", 0, None)
    N.check(L.cama_clip_workspace_bytes(ctypes.byref(desc), ctypes.byref(need)))
    ws = rt.scratch("abi-err", need.value)
    assert L.cama_clip_render(rt.ctx, ctypes.byref(desc), rt.ptr(ws), 1024, rt.stream()) == N.CAMA_E_WORKSPACE
    assert b"workspace" in L.cama_last_error()
    desc.n_cams = 9
    assert L.cama_clip_render(rt.ctx, ctypes.byref(desc), rt.ptr(ws), ws.numel(), rt.stream()) == N.CAMA_E_UNSUPPORTED
    launched = rt.launches()
    assert launched > 0


# ------------------------------------------------------------------ BASELINE.json full-size configurations
@pytest.fixture(scope="module")
def config2_clip(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("config2"))
    spec = synth.config2_spec()
    spec.write_cama = False                                  # nuScenes-style labels: N ~ 96 k
    return synth.write_clip(spec, root)


def test_config2_full_clip_vs_oracle(rt, config2_clip):
    """BASELINE.json configs[1]: 40 frames x 6 cameras x ~200 polylines, every frame against the
    C oracle, two frames against the NumPy/OpenCV restatement of the reference loop."""
    from cama_b200.batched import Reproject
    rp = Reproject(synth.CAMA_CONFIGS, config2_clip, device=0)
    idx, frames = rp("nuscenes")
    assert idx == list(range(1, 41)) and frames.shape == (40, 6, H, W, 3)
    oc = orc.ClipOracle(synth.CAMA_CONFIGS, config2_clip)
    flat, classes, counts = orc.flatten(oc.instance_maps["nuscenes"], 3)
    assert len(counts) == 200 and 90000 < len(flat) < 100000
    per_frame = oc.world_to_chassis_per_frame("nuscenes")
    w2c = np.stack([m for _, m in per_frame])
    offs = np.concatenate([[0], np.cumsum(counts)])
    want, cc, vc = oracle_c.clip_render(flat, offs, bgr_of(classes), w2c, np.stack(oc.chassis2cam), np.stack(oc.K), BOX6, H, W)
    assert np.array_equal(frames, want)
    assert int(vc.sum()) > 1_000_000                          # ~42 k visible points per frame
    # counts through the debug outputs, both raster modes identical
    res = rp.resident("nuscenes")
    out_b, dbg = rp.renderer.render(res, to_dev(w2c.reshape(-1, 16)), mode="binned", debug=True)
    out_p = rp.renderer.render(res, to_dev(w2c.reshape(-1, 16)), mode="plane")
    assert np.array_equal(dbg["crop_counts"].cpu().numpy(), cc) and np.array_equal(dbg["visible_counts"].cpu().numpy(), vc)
    assert bool((out_b == out_p).all())
    # NumPy/OpenCV oracle (the reference's own loop structure) on the first and last frame
    for k, (image_idx, chassis) in enumerate(oc.frames("nuscenes")):
        if k not in (0, 39):
            continue
        per_cam = oc.project_all(chassis)
        for c, cam in enumerate(oc.cameras):
            img = orc.render_instances(np.zeros((H, W, 3), np.uint8), per_cam[cam])
            assert np.array_equal(frames[k, c], img)


def test_frame_sharding_is_exact(rt, config2_clip):
    """Frames are independent: rendering blocks of frames separately (what each rank of a
    multi-GPU run does) reproduces the whole-clip result bit for bit."""
    from cama_b200.batched import Reproject
    import torch
    rp = Reproject(synth.CAMA_CONFIGS, config2_clip, device=0)
    _, w2c = rp.frame_poses("nuscenes")
    whole = rp.render_device("nuscenes", w2c=w2c)
    parts = [rp.render_device("nuscenes", w2c=w2c[lo:lo + 13]) for lo in range(0, 40, 13)]
    assert bool((torch.cat(parts) == whole).all())
    # idempotence: compositing the overlay over its own output changes nothing
    again = rp.render_device("nuscenes", w2c=w2c, background=whole.clone())
    assert bool((again == whole).all())


def test_frame_group_pipeline_is_exact(rt, config2_clip):
    """cama_clip_desc.pipeline_frames: the clip rendered as frame groups on three stream lanes (geometry of group
    g+1 under the sort and raster of group g) equals the one-pass render bit for bit — dense frames, composite
    over a background, sparse records (chunk indices carry the group's frame offset) — with a ragged last group,
    repeated calls on one workspace, and the counters summed over the groups."""
    from cama_b200.batched import Reproject
    import torch
    rp = Reproject(synth.CAMA_CONFIGS, config2_clip, device=0)
    r = rp.renderer
    _, w2c = rp.frame_poses("nuscenes")
    r.pipeline_frames = -1
    whole = rp.render_device("nuscenes", w2c=w2c).clone()
    stats_whole = dict(r.last_stats)
    bg = torch.randint(0, 256, whole.shape, dtype=torch.uint8, device=whole.device)
    comp_whole = rp.render_device("nuscenes", w2c=w2c, background=bg.clone()).clone()
    _, sparse_whole = rp("nuscenes")
    sparse_whole = sparse_whole.copy()
    for group in (8, 16, 24):                        # 40 frames: 5 groups; 16+16+8; 24+16
        r.pipeline_frames = group
        for _ in range(2):
            got = rp.render_device("nuscenes", w2c=w2c)
            assert bool((got == whole).all()), group
        assert r.last_stats["records_total"] == stats_whole["records_total"] and not r.last_stats["overflow"]
        comp = rp.render_device("nuscenes", w2c=w2c, background=bg.clone())
        assert bool((comp == comp_whole).all()), group
        _, sparse = rp("nuscenes")
        assert np.array_equal(sparse, sparse_whole), group
    r.pipeline_frames = 0


def test_overlay_expand_equals_dense(rt, config2_clip):
    """cama_overlay_expand (the receiving end of the sparse all-gather): the records of the sparse output, expanded
    on the device, are the dense render byte for byte — both record formats, whole clip and a block of frames
    expanded into the middle of a larger tensor (what every rank does with its peers' records)."""
    from cama_b200 import _native as N
    from cama_b200.batched import Reproject
    import torch
    rp = Reproject(synth.CAMA_CONFIGS, config2_clip, device=0)
    r, res = rp.renderer, rp.resident("nuscenes")
    _, w2c = rp.frame_poses("nuscenes")
    dense = rp.render_device("nuscenes", w2c=w2c).clone()
    w2c_dev = torch.from_numpy(w2c).to(rp.rt.device)
    for fmt in (N.OVERLAY_PALETTE, N.OVERLAY_BGR):
        records, n, got_fmt = r.render_overlay(res, w2c_dev, fmt=fmt)
        assert got_fmt == fmt and n > 0
        out = torch.full(dense.shape, 7, dtype=torch.uint8, device=dense.device)
        r.expand_overlay(records, n, fmt, res.palette, dense.shape[0], out=out)
        assert bool((out == dense).all()), fmt
    # a block of frames (13..26) rendered on its own and expanded into its place
    records, n, fmt = r.render_overlay(res, w2c_dev[13:26].contiguous())
    big = torch.zeros_like(dense)
    r.expand_overlay(records[:n].clone(), n, fmt, res.palette, 13, out=big[13:26])
    assert bool((big[13:26] == dense[13:26]).all()) and int(big[:13].sum()) == 0 and int(big[26:].sum()) == 0


def test_render_sharded_single_process(rt, config2_clip):
    """cama_b200.shard with no process group = the whole clip; the multi-rank path is covered by
    tests/test_shard_gloo.py (CPU) and tools/multi_gpu_check.py (gpurun --gpus N)."""
    from cama_b200 import shard
    from cama_b200.batched import Reproject
    rp = Reproject(synth.CAMA_CONFIGS, config2_clip, device=0)
    idx, frames = shard.render_sharded(rp, "nuscenes")
    assert idx == list(range(1, 41))
    assert bool((frames == rp.render_device("nuscenes")).all())


def test_sparse_transfer_equals_dense(rt, config2_clip, clip_root):
    """Reproject.__call__ default (sparse overlay records over PCIe + host draw) gives the same bytes
    as the dense render, on blank frames (buffer reused across calls: previous overlay blanked), on
    host backgrounds drawn in place, and on the golden clip."""
    from cama_b200.batched import Reproject
    rp = Reproject(synth.CAMA_CONFIGS, config2_clip, device=0)
    idx, sparse = rp("nuscenes")
    assert rp.last_transfer["mode"] == "sparse" and 0 < rp.last_transfer["d2h_bytes"] < sparse.size // 5
    dense = rp("nuscenes", transfer="dense")[1].copy()           # (the dense path returns its reused pinned buffer)
    assert rp.last_transfer["mode"] == "dense"
    assert np.array_equal(sparse, dense)
    # second call with other poses into the same host buffer: the old overlay must be gone
    _, w2c = rp.frame_poses("nuscenes")
    rp.cm._trajectory_cache["nuscenes"][0].transform(np.array([[1, 0, 0, 0.5], [0, 1, 0, -0.25], [0, 0, 1, 0], [0, 0, 0, 1.0]]))
    _, sparse2 = rp("nuscenes")
    dense2 = rp("nuscenes", transfer="dense")[1].copy()
    assert sparse2 is sparse and np.array_equal(sparse2, dense2) and not np.array_equal(dense2, dense)
    # host backgrounds: drawn in place like the reference draws on the camera images
    rng = np.random.default_rng(11)
    bg = rng.integers(0, 256, size=dense.shape, dtype=np.uint8)
    keep = bg.copy()
    _, out = rp("nuscenes", backgrounds=bg)
    assert out is bg
    want = rp("nuscenes", backgrounds=keep.copy(), transfer="dense")[1]
    assert np.array_equal(bg, want) and not np.array_equal(bg, keep)
    lit = (dense2 != 0).any(-1)
    assert np.array_equal(bg[~lit], keep[~lit])
    # N4: straight into the 2x3 mosaic of cama/tools.py:22-25 (what VideoGenerator.add_frame takes)
    from cama_b200.tools import concate_image
    _, mosaic = rp("nuscenes", layout="mosaic")
    assert mosaic.shape == (40, 1080, 2880, 3)
    dicts = rp.as_image_dicts(dense2)
    for f in (0, 17, 39):
        assert np.array_equal(mosaic[f], concate_image(dicts[f]))
    # golden clip through the sparse path
    g = load_golden("golden_clip_cama_exact.npz")
    clip = synth.write_clip(synth.tiny_spec(name="tiny_sparse"), clip_root)
    idx, frames = Reproject(synth.CAMA_CONFIGS, clip, device=0)("cama")
    assert idx == list(g["frame_idx"]) and np.array_equal(frames, g["frames"])


def test_overlay_records_are_unique_chunks(rt):
    """Raw C-ABI sparse output, both record formats: every lit chunk exactly once, masks/colours/palette
    entries consistent with the dense frames."""
    from cama_b200 import _native as N
    g = load_golden("golden_clip_nuscenes_slerp.npz")
    r = renderer_for(g)
    res = r.resident(golden_instances(g))
    assert res.palette is not None and int((res.palette.any(1)).sum()) == 2          # lane grey + crosswalk yellow
    dense = g["frames"].reshape(-1, 8, 3)
    lit_chunks = set(np.flatnonzero(dense.reshape(len(dense), -1).any(1)).tolist())
    w2c = to_dev(g["world2chassis"].reshape(-1, 16))
    # 32-byte BGR records
    records, n, fmt = r.render_overlay(res, w2c, fmt=N.OVERLAY_BGR)
    assert fmt == N.OVERLAY_BGR and tuple(records.shape[1:]) == (8,)
    rec = records[:n].cpu().numpy().view(np.uint8).reshape(n, 32)
    chunk = rec[:, :4].copy().view("<u4")[:, 0]
    mask = rec[:, 4:8].copy().view("<u4")[:, 0]
    assert len(np.unique(chunk)) == n and mask.max() <= 0xFF and mask.min() >= 1
    assert lit_chunks <= set(chunk.tolist())                   # (a chunk painted pure black would be a record but not lit)
    bits = (mask[:, None] >> np.arange(8)) & 1
    bgr = rec[:, 8:].reshape(n, 8, 3)
    assert np.array_equal(bgr * bits[:, :, None], dense[chunk] * bits[:, :, None])
    assert not (bgr * (1 - bits)[:, :, None]).any() and not (dense[chunk] * (1 - bits)[:, :, None]).any()
    # 12-byte palette records (the default when the instances have <= 255 colours)
    records, n2, fmt = r.render_overlay(res, w2c)
    assert fmt == N.OVERLAY_PALETTE and n2 == n and tuple(records.shape[1:]) == (3,)
    rec = records[:n2].cpu().numpy().view(np.uint8).reshape(n2, 12)
    chunk2 = rec[:, :4].copy().view("<u4")[:, 0]
    index = rec[:, 4:]
    assert sorted(chunk2.tolist()) == sorted(chunk.tolist())
    assert np.array_equal(res.palette[index], dense[chunk2]) and index.max() <= 2


def test_device_densify_matches_host_and_oracle(rt):
    """Scope row N2: cama_densify_plan/fill against the NumPy restatement of cama/reproject.py:42-106 and the
    C oracle, bit for bit: segment lengths on both sides of multiples of the resolution, sub-resolution
    segments, negative / large coordinates, the BEV height lookup with clipping, a polyline without points."""
    from cama_b200.reproject import MapManager
    rng = np.random.default_rng(21)
    labels = []
    for k in range(60):
        n = int(rng.integers(2, 9))
        steps = rng.choice([0.05, 0.0999999, 0.1, 0.1000001, 0.3, 0.7, 2.35, 17.0], size=n - 1) * rng.choice([1.0, 1.0, 10.0])
        ang = rng.uniform(0, 2 * np.pi, n - 1)
        pts = np.cumsum(np.concatenate([rng.uniform(-300, 6300, (1, 2)), np.stack([steps * np.cos(ang), steps * np.sin(ang)], 1)]), axis=0)
        if not (np.linalg.norm(np.diff(pts.astype(np.float32), axis=0), axis=1) / np.float32(0.1)).astype(np.int64).any():
            pts[-1] += 3.0                      # keep every polyline drawable here; the empty case is checked below
        labels.append({"attrs": {"type": synth.MAP_CLASSES[k % 3]}, "data": pts.tolist()})
    labels.insert(7, {"attrs": {"type": "lane_marking"}, "data": [[5.0, 5.0]]})            # skipped by the reference
    labels.append({"attrs": {"type": "lane_marking"}, "data": [[0, 0], [1, 0], [1.05, 0], [1.35, 0]]})
    host, dev = MapManager(device=0, densify="host"), MapManager(device=0, densify="device")
    a, b = host.load_3d_instance_maps(labels), dev.load_3d_instance_maps(labels)
    assert [i["class"] for i in a] == [i["class"] for i in b] and len(a) == 61
    for x, y in zip(a, b):
        assert y["points"].dtype == np.float32 and np.array_equal(x["points"], y["points"])
    assert np.array_equal(b[-1]["points"][:, :2], oracle_c.densify(np.array(labels[-1]["data"])))
    assert dev.device_vertices(b) is not None and dev.device_vertices(a) is None
    verts = dev.device_vertices(b).cpu().numpy()
    assert np.array_equal(verts[:, :3], np.concatenate([i["points"] for i in b]))
    assert np.array_equal(verts[:, 3].view(np.int32), np.repeat(np.arange(61, dtype=np.int32), [len(i["points"]) for i in b]))
    # pixel labels + height map (indices beyond the map are clipped, both with shape[0]-1)
    bev = rng.standard_normal((500, 700)).astype(np.float32)
    a, b = host.calculate_3d_instance_maps(bev, labels), dev.calculate_3d_instance_maps(bev, labels)
    for x, y in zip(a, b):
        assert np.array_equal(x["points"], y["points"])
    want = orc.instances_from_pixel_labels(bev, labels)
    assert all(np.array_equal(x["points"], y["points"]) for x, y in zip(want, b))
    # a float64 height map promotes the points to float64: that branch stays on the host
    c = dev.calculate_3d_instance_maps(bev.astype(np.float64), labels)
    assert c[0]["points"].dtype == np.float64 and dev.device_vertices(c) is None
    # a polyline whose segments are all shorter than the resolution: the reference raises IndexError
    for mm in (host, dev):
        with pytest.raises(IndexError):
            mm.load_3d_instance_maps([{"attrs": {"type": "lane_marking"}, "data": [[0, 0], [0.05, 0]]}])
    assert dev.load_3d_instance_maps([]) == []


def test_remap_matches_cv2(rt):
    """Scope row N1: cama_remap_bilinear == cv2.remap(INTER_LINEAR) on uint8 BGR, including taps that
    leave the source (border constant 0) and maps cycling per camera."""
    import cv2
    import torch
    from cama_b200.batched import ClipRenderer
    rng = np.random.default_rng(8)
    r = ClipRenderer(np.eye(4)[None], np.eye(3)[None], H, W, BOX6, device=0)
    src = rng.integers(0, 256, size=(4, 900, 1600, 3), dtype=np.uint8)
    maps = []
    for cam in ("camera_front", "camera_rear"):
        K = synth.camera_intrinsics(cam)
        Kn = K.copy()
        Kn[0] *= W / 1600
        Kn[1] *= H / 900
        maps.append(cv2.initUndistortRectifyMap(K, np.zeros(8), None, Kn, (W, H), cv2.CV_32FC1))
    maps[1] = (maps[1][0] * np.float32(1.1) - np.float32(40.3), maps[1][1] * np.float32(1.2) - np.float32(60.7))   # leaves the image on all sides
    mx = to_dev(np.stack([m[0] for m in maps]))
    my = to_dev(np.stack([m[1] for m in maps]))
    got = r.remap(to_dev(src), mx, my).cpu().numpy()
    for i in range(4):
        want = cv2.remap(src[i], maps[i % 2][0], maps[i % 2][1], interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(got[i], want), i
    assert (got[1] == 0).all(-1).any()                         # the border value really occurs
    # exact half-way fractions: cvRound is round-half-to-even
    half = np.tile((np.arange(W, dtype=np.float32) + np.float32(0.5)) / np.float32(32), (H, 1))
    flipped = np.ascontiguousarray(half[:, ::-1])
    want = cv2.remap(src[0], half, flipped, interpolation=cv2.INTER_LINEAR)
    got = r.remap(to_dev(src[:1]), to_dev(half[None]), to_dev(flipped[None])).cpu().numpy()[0]
    assert np.array_equal(got, want)


def test_render_vectors_with_raw_camera_images(rt, clip_root):
    """ClipManager.render_vectors after the JPEG decode, batched on the device: undistort-resize + draw in
    place == cv2.remap of each image, then the reference's render_maps (here: the golden overlay on top)."""
    import cv2
    import torch
    from cama_b200.batched import Reproject
    g = load_golden("golden_clip_nuscenes_exact.npz")
    clip = synth.write_clip(synth.tiny_spec(name="tiny_raw"), clip_root)
    rp = Reproject(synth.CAMA_CONFIGS, clip, device=0)
    rng = np.random.default_rng(2)
    raw = rng.integers(0, 256, size=(3, 6, 900, 1600, 3), dtype=np.uint8)
    frames = rp.render_device("nuscenes", raw_backgrounds=to_dev(raw)).cpu().numpy()
    lit = g["frames"].any(-1)
    for f in range(3):
        for c, cam in enumerate(rp.cm.cm_list):
            bg = cam.resize_image(raw[f, c])                  # cv2.remap with the camera's maps (reference :232-240)
            want = np.where(lit[f, c][..., None], g["frames"][f, c], bg)
            assert np.array_equal(frames[f, c], want), (f, c)


def test_cama_dense_labels_vs_oracle(rt, tmp_path):
    """CAMA-label branch (0.1 px densify, BEV height lookup, N ~ 1.0 M vertices), 4 frames."""
    from cama_b200.batched import Reproject
    spec = synth.config2_spec(n_frames=4, name="config2_cama")
    spec.write_nuscenes = False
    clip = synth.write_clip(spec, str(tmp_path))
    rp = Reproject(synth.CAMA_CONFIGS, clip, device=0)
    idx, frames = rp("cama")
    oc = orc.ClipOracle(synth.CAMA_CONFIGS, clip)
    flat, classes, counts = orc.flatten(oc.instance_maps["cama"], 3)
    assert len(flat) > 900_000 and flat.dtype == np.float32
    got_flat, _, _ = orc.flatten(rp.cm.instance_maps["cama"], 3)
    assert np.array_equal(got_flat, flat)                     # load-time densify + height lookup, bit-exact
    w2c = np.stack([m for _, m in oc.world_to_chassis_per_frame("cama")])
    offs = np.concatenate([[0], np.cumsum(counts)])
    want, _, _ = oracle_c.clip_render(flat, offs, bgr_of(classes), w2c, np.stack(oc.chassis2cam), np.stack(oc.K), BOX6, H, W)
    assert idx == [1, 2, 3, 4] and np.array_equal(frames, want)


def test_config3_site_properties(rt, tmp_path):
    """BASELINE.json configs[2] (320 frames x 6 cams, N ~ 767 k): too slow for a per-pixel CPU
    check of every frame, so: 6 sampled frames against the C oracle, and the size-independent
    properties — BINNED == PLANE on a frame block, sharded == whole."""
    from cama_b200.batched import Reproject
    import torch
    spec = synth.config3_spec()
    spec.write_cama = False
    clip = synth.write_clip(spec, str(tmp_path))
    rp = Reproject(synth.CAMA_CONFIGS, clip, device=0)
    idx, w2c = rp.frame_poses("nuscenes")
    assert len(idx) == 320
    whole = rp.render_device("nuscenes", w2c=w2c)
    assert tuple(whole.shape) == (320, 6, H, W, 3)
    inst = rp.cm.instance_maps["nuscenes"]
    flat, classes, counts = orc.flatten(inst, 3)
    assert len(counts) == 1600 and 700_000 < len(flat) < 850_000
    offs = np.concatenate([[0], np.cumsum(counts)])
    c2c = np.stack([np.asarray(c.get_chassis2camera()) for c in rp.cm.cm_list])
    K = np.stack([c.K for c in rp.cm.cm_list])
    sample = [0, 1, 77, 160, 250, 319]
    want, _, _ = oracle_c.clip_render(flat, offs, bgr_of(classes), w2c[sample], c2c, K, BOX6, H, W)
    assert np.array_equal(whole[sample].cpu().numpy(), want)
    # tile culling (tile_bounds) only skips work: without it the frames are the same
    res = rp.resident("nuscenes")
    bounds, res.tile_bounds = res.tile_bounds, None
    try:
        unculled = rp.render_device("nuscenes", w2c=w2c[40:80])
    finally:
        res.tile_bounds = bounds
    assert bool((unculled == whole[40:80]).all())
    block = rp.render_device("nuscenes", w2c=w2c[100:140], mode="plane")
    assert bool((block == whole[100:140]).all())
    halves = torch.cat([rp.render_device("nuscenes", w2c=w2c[:160]), rp.render_device("nuscenes", w2c=w2c[160:])])
    assert bool((halves == whole).all())


# ------------------------------------------------------------------ pixel / image boundaries of the float32 fast path
@pytest.mark.parametrize("eps_set", [0, 1, 2])
def test_pixel_boundary_fuzz(rt, eps_set):
    """candidate_pixel (csrc/clip.cu) decides most pixels from a float32 estimate of q_x/q_z and takes the IEEE
    division only within 4e-3 of an integer.  Seeded fuzz of exactly that boundary through the production kernel
    (float32 vertices, BINNED, no debug outputs) against oracle.c: projected coordinates u = j + e, v = i + e with
    e from +-1e-12 to +-6e-3 (one pair per camera, eight cameras), depths from 0.5 to 128 m, including u = 0 +- e and
    u = W +- e (the image-border comparisons of cama/reproject.py:192-198) — built exactly: K = [[1024,0,cx+e],..]
    with float64 principal points, x = (j - cx) z / 1024 exactly representable in float32 for z a power of two —
    plus random depths and positions.  Discs are 5 px apart, so a centre off by one pixel changes the image."""
    from cama_b200.batched import ClipRenderer
    eps = [[1e-12, -1e-12, 1e-9, -1e-7, 1e-5, -1e-4, 3.9e-3, -4.1e-3],
           [2e-3, -3e-3, 5e-3, -6e-3, 1e-3, -1e-11, 1e-10, -1e-8],
           [4.0e-3, -4.0e-3, 2.9e-3, -5.1e-3, 0.0, 1e-13, -1e-13, 3e-6]][eps_set]
    rng = np.random.default_rng(100 + eps_set)
    cx0, cy0 = 100.0, 60.0
    K = np.stack([np.array([[1024.0, 0, cx0 + e], [0, 1024.0, cy0 + (e if c % 2 else -e)], [0, 0, 1.0]]) for c, e in enumerate(eps)])
    E = np.stack([np.eye(4)] * 8)
    zs = 2.0 ** np.arange(-1, 8)                                   # 0.5 .. 128
    jj, ii = np.meshgrid(np.arange(0, W + 1, 5), np.arange(0, H + 1, 5))   # includes u = W and v = H: never visible unless e < 0
    jj, ii = jj.ravel().astype(np.float64), ii.ravel().astype(np.float64)
    z = zs[(jj.astype(int) // 5 + ii.astype(int) // 5) % len(zs)]
    exact = np.stack([(jj - cx0) * z / 1024.0, (ii - cy0) * z / 1024.0, z], axis=1).astype(np.float32)
    assert np.array_equal(exact.astype(np.float64)[:, 0] * 1024.0 / z + cx0, jj)      # the construction is exact
    zr = rng.uniform(0.5, 200.0, size=60000)
    rand = np.stack([(rng.uniform(-3, W + 3, size=zr.size) - cx0) * zr / 1024.0, (rng.uniform(-3, H + 3, size=zr.size) - cy0) * zr / 1024.0, zr], axis=1).astype(np.float32)
    pts = np.concatenate([exact, rand])
    offs = np.array([0, len(exact) // 2, len(exact), len(pts)], np.int64)
    classes = ["lane_marking", "Road_teeth", "lane_marking"]
    box = [-1000, 1000, -1000, 1000, -1000, 1000]
    w2c = np.stack([np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)])
    w2c[1, 2, 3] = 0.0                                             # (second frame identical: two frames exercise the frame loop)
    want, cc, vc = oracle_c.clip_render(pts, offs, bgr_of(classes), w2c, E, K, box, H, W)
    r = ClipRenderer(E, K, H, W, box, device=0)
    res = r.resident([{"class": c, "points": pts[offs[i]:offs[i + 1]]} for i, c in enumerate(classes)])
    w2c_dev = to_dev(w2c.reshape(-1, 16))
    for mode in ("binned", "plane"):
        frames = r.render(res, w2c_dev, mode=mode).cpu().numpy()              # the production kernel: no debug outputs
        diff = np.argwhere((frames != want).any(-1))
        assert diff.size == 0, f"{mode}: {len(diff)} pixels differ, first at frame/cam/row/col {diff[0].tolist()} (eps {eps[diff[0][1]]})"
    _, dbg = r.render(res, w2c_dev, mode="binned", debug=True)
    assert np.array_equal(dbg["visible_counts"].cpu().numpy(), vc) and np.array_equal(dbg["crop_counts"].cpu().numpy(), cc)
    assert int(vc.sum()) > 500_000


# ------------------------------------------------------------------ culling aids and occupancy knobs never change a pixel
def test_culling_aids_and_occupancy_knobs_do_not_change_frames(rt, config2_clip):
    """camera table (cama_camera_table_build), warp bounds, tile bounds, geometry / raster occupancy: all of them
    only skip or reschedule work.  Also: a table built for ANOTHER rig is recognised by its signature and ignored."""
    import torch
    from cama_b200.batched import ClipRenderer, Reproject
    rp = Reproject(synth.CAMA_CONFIGS, config2_clip, device=0)
    r, res = rp.renderer, rp.resident("nuscenes")
    _, w2c = rp.frame_poses("nuscenes")
    w2c_dev = to_dev(w2c[:12])
    want = r.render(res, w2c_dev, mode="binned").clone()
    table = r.camera_table()
    assert int(table[16:].count_nonzero().item()) > 1000 and int((table[16:] == 0).sum().item()) > 100      # some cells see cameras, some none
    # no table at all / a table of another rig (different intrinsics => different signature)
    other = ClipRenderer(r.chassis2cam, r.intrinsics * 1.01, r.height, r.width, r.crop_box, device=0)
    for stand_in in (torch.zeros_like(table), other.camera_table()):
        r._camera_table = stand_in
        try:
            assert torch.equal(r.render(res, w2c_dev, mode="binned"), want)
        finally:
            r._camera_table = table
    # no bounds
    saved = res.warp_bounds, res.tile_bounds
    for wb, tb in ((None, saved[1]), (saved[0], None), (None, None)):
        res.warp_bounds, res.tile_bounds = wb, tb
        try:
            assert torch.equal(r.render(res, w2c_dev, mode="binned"), want)
        finally:
            res.warp_bounds, res.tile_bounds = saved
    # occupancy
    for geo, ras in ((3, 0), (2, 3), (1, 1)):
        r.geometry_ctas_per_sm, r.raster_ctas_per_sm = geo, ras
        try:
            assert torch.equal(r.render(res, w2c_dev, mode="binned"), want)
        finally:
            r.geometry_ctas_per_sm = r.raster_ctas_per_sm = 0
    assert torch.equal(r.render(res, w2c_dev, mode="plane"), want)


def test_render_maps_on_host_images_both_paths(rt):
    """CameraManager.render_maps on a host image: the overlay-record path (writeable image, width % 8 == 0) and the
    upload / download path (read-only image) paint the same bytes as the golden frames; untouched pixels keep their value."""
    from cama_b200.runtime import get_runtime
    g = load_golden("golden_clip_nuscenes_exact.npz")
    runtime = get_runtime(0)
    rng = np.random.default_rng(9)
    n_inst = len(g["inst_classes"])
    bgr = bgr_of(g["inst_classes"])
    pos = 0
    for c in range(2):                                      # frame 0, cameras 0 and 1
        counts = g["vu_counts"][0, c]
        n = int(counts.sum())
        vu = g["vu_points"][pos:pos + n]
        pos += n
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        background = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        lit = g["frames"][0, c].any(-1)
        want = np.where(lit[..., None], g["frames"][0, c], background)
        img = background.copy()
        assert runtime.render_points(img, vu, offs, bgr) is img and np.array_equal(img, want)
        frozen = background.copy()
        frozen.setflags(write=False)
        assert np.array_equal(runtime.render_points(frozen, vu, offs, bgr), want) and np.array_equal(frozen, background)
    assert n_inst == len(offs) - 1


def test_device_mosaic_layout(rt, config2_clip):
    """Scope row N4: render_device(layout="mosaic") writes straight into the 2x3 camera mosaic of
    VideoGenerator.concate_image (cama/tools.py:22-25); it must equal concate_image of the plain frames — blank,
    composited over a background that is itself a mosaic, and through the frame-group pipeline."""
    import torch
    from cama_b200.batched import Reproject
    from cama_b200.tools import concate_image
    rp = Reproject(synth.CAMA_CONFIGS, config2_clip, device=0)
    _, w2c = rp.frame_poses("nuscenes")
    w2c = w2c[:9]
    plain = rp.render_device("nuscenes", w2c=w2c).cpu().numpy()
    want = np.stack([concate_image({n: plain[f, c] for c, n in enumerate(rp.camera_names)}) for f in range(len(w2c))])
    got = rp.render_device("nuscenes", w2c=w2c, layout="mosaic")
    assert tuple(got.shape) == (9, 2 * H, 3 * W, 3) and np.array_equal(got.cpu().numpy(), want)
    # in place over a background mosaic
    rng = np.random.default_rng(3)
    bg = rng.integers(0, 256, size=(9, 2 * H, 3 * W, 3), dtype=np.uint8)
    lit = want.any(-1, keepdims=True)
    bg_dev = to_dev(bg)
    out = rp.render_device("nuscenes", w2c=w2c, layout="mosaic", background=bg_dev, out=bg_dev)
    assert np.array_equal(out.cpu().numpy(), np.where(lit, want, bg))
    # out of place (every cell copied) and the frame-group pipeline (groups of 8 frames)
    out2 = rp.render_device("nuscenes", w2c=w2c, layout="mosaic", background=to_dev(bg))
    assert np.array_equal(out2.cpu().numpy(), np.where(lit, want, bg))
    rp.renderer.pipeline_frames = 8
    try:
        assert torch.equal(rp.render_device("nuscenes", w2c=w2c, layout="mosaic"), got)
    finally:
        rp.renderer.pipeline_frames = 0


def test_split_phases_with_external_lists(rt, config2_clip):
    """The two halves of the pipeline on record lists outside the workspace (cama_clip_desc.phases / list_*), as the
    list exchange of a frame-sharded clip runs them: geometry of frame blocks [0,15), [15,29), [29,40), then ONE raster
    call over all 40 frames — from the arrays the geometry wrote and from a copy of their filled parts (what
    cama_peer_publish_lists makes on a peer): both equal the ordinary render."""
    import torch
    from cama_b200 import _native as N
    from cama_b200.batched import Reproject
    rp = Reproject(synth.CAMA_CONFIGS, config2_clip, device=0)
    r, res = rp.renderer, rp.resident("nuscenes")
    _, w2c = rp.frame_poses("nuscenes")
    w2c_dev = to_dev(w2c)
    want = r.render(res, w2c_dev, mode="binned")
    stats = r.last_stats
    F, cap, per_image = 40, int(stats["record_capacity"]), int(stats["lists_per_image"])
    n_lists = F * r.n_cams * per_image
    arrays = []
    for _ in range(2):                                     # primary, mirror
        arrays.append((torch.full((n_lists, cap), 0x7fffffff, dtype=torch.int32, device="cuda"), torch.full((n_lists,), 12345, dtype=torch.int32, device="cuda")))
    (rec, cur), (rec_m, cur_m) = arrays
    for lo, hi in ((0, 15), (15, 29), (29, 40)):
        r.enqueue_phase(res, w2c_dev[lo:hi], hi - lo, {"phases": N.PHASE_GEOMETRY, "records_ptr": rec.data_ptr(), "cursor_ptr": cur.data_ptr(),
                                                        "frame_base": lo, "frames": F}, cap)
    cur_m.copy_(cur)                                       # (what cama_peer_publish_lists does for a peer: lengths + the filled part of every list)
    filled = torch.arange(cap, device="cuda")[None, :] < cur[:, None]
    rec_m[filled] = rec[filled]
    for records, cursor in ((rec, cur), (rec_m, cur_m)):
        out = torch.empty_like(want)
        r.enqueue_phase(res, None, F, {"phases": N.PHASE_RASTER, "records_ptr": records.data_ptr(), "cursor_ptr": cursor.data_ptr(), "frame_base": 0, "frames": F},
                        cap, out=out)
        assert torch.equal(out, want)
    assert int(cur.sum().item()) == int(stats["records_total"])
    # a raster call over a sub-range of the lists' frames
    part = torch.empty_like(want[10:25])
    r.enqueue_phase(res, None, 15, {"phases": N.PHASE_RASTER, "records_ptr": rec.data_ptr(), "cursor_ptr": cur.data_ptr(), "frame_base": 10, "frames": F}, cap, out=part)
    assert torch.equal(part, want[10:25])
