"""Synthetic nuScenes-shaped clips in CAMA's on-disk clip format.

There is no dataset and no network on the build or GPU boxes, so every test and
benchmark runs on clips produced here.  The directory written by
:func:`write_clip` is the real clip layout the reference consumes
(`attribute.json`, `odometry/*.txt`, `maps/*.json`, `maps/*.npy`; layout defined
by /root/reference/dataset/nuscenes2clip.py:661-712 and read back by
/root/reference/cama/dataset_reader.py:19-43,150-294,409-411), so the unmodified
reference and this package can be pointed at the same directory.

Shapes follow SURVEY.md section 8(d):

* ``config1``  1 frame, camera_front only, one 50-vertex lane polyline.
* ``config2``  40 frames x 6 cams, 200 polylines x 50 raw vertices (0.1 m
  densify => N ~ 96 k; the CAMA-label variant densifies at 0.1 px => N ~ 1.0 M).
* ``config3``  320 frames x 6 cams, 1600 arcs around a circular trajectory.

Everything is seeded with ``numpy.random.default_rng``; this module never
touches the GPU and never imports anything from ``oracle/``.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np

CAMERA_LIST = ["camera_front_left", "camera_front", "camera_front_right",
               "camera_rear_left", "camera_rear", "camera_rear_right"]
MAP_CLASSES = ["lane_marking", "Road_teeth", "Crosswalk_Line"]

# same keys as /root/reference/config.yaml:17-24
CAMA_CONFIGS = {
    "result_dir": "maps",
    "camera_list": list(CAMERA_LIST),
    "camera_main": "camera_front",
    "height_mlp": "vision_road_mlp_ft.npy",
    "pose_prefix": "scmv",
    "cama_map_file": "map_labels.json",
    "nuscenes_map_file": "map_nuscenes.json",
}

_CAM_YAW_DEG = {"camera_front_left": 55.0, "camera_front": 0.0, "camera_front_right": -55.0,
                "camera_rear_left": 110.0, "camera_rear": 180.0, "camera_rear_right": -110.0}
_T0_MS = 1533151600000
_MAP_HALF = 300.0      # metres; CAMA BEV map is 600 m x 600 m at 0.1 m / px
_MAP_RES = 0.1


def _rot_z(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def camera_to_chassis(camera_name):
    """4x4 pose of a camera in the chassis frame (x fwd, y left, z up).

    Camera axes are the usual optical ones (x right, y down, z forward).
    """
    yaw = np.deg2rad(_CAM_YAW_DEG[camera_name])
    # optical axes expressed in a chassis-aligned frame looking along +x
    base = np.array([[0.0, 0.0, 1.0],
                     [-1.0, 0.0, 0.0],
                     [0.0, -1.0, 0.0]])
    out = np.eye(4)
    out[:3, :3] = _rot_z(yaw) @ base
    out[:3, 3] = [1.5 * np.cos(yaw), 0.5 * np.sin(yaw), 1.5]
    return out


def camera_intrinsics(camera_name):
    f = 809.2 if camera_name == "camera_rear" else 1266.4
    return np.array([[f, 0.0, 816.3], [0.0, f, 491.5], [0.0, 0.0, 1.0]])


def _quat_xyzw_from_matrix(rot):
    """Rotation matrix -> unit quaternion (x, y, z, w), w >= 0."""
    m = rot
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        w, x, y, z = 0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        w, x, y, z = (m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        w, x, y, z = (m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        w, x, y, z = (m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s
    q = np.array([x, y, z, w])
    q /= np.linalg.norm(q)
    return -q if q[3] < 0 else q


def _tum_rows(timestamps_s, poses):
    rows = np.zeros((len(poses), 8))
    for i, (t, pose) in enumerate(zip(timestamps_s, poses)):
        rows[i, 0] = t
        rows[i, 1:4] = pose[:3, 3]
        rows[i, 4:8] = _quat_xyzw_from_matrix(pose[:3, :3])
    return rows


@dataclass
class ClipSpec:
    """Everything needed to write one synthetic clip directory."""
    name: str
    n_frames: int                       # rendered frames; F+1 timestamps are written (index 0 is skipped)
    cameras: list = field(default_factory=lambda: list(CAMERA_LIST))
    polylines_m: list = field(default_factory=list)   # list of (class, (k,2) float metres, map frame)
    chassis_poses: list = field(default_factory=list)  # F+1 chassis->map 4x4
    pose_time_offset_ms: int = 0        # !=0 => pose stamps differ from frame stamps (slerp branch)
    frame_period_ms: int = 500
    height_sigma: float = 0.05
    bev_size: int = 6000                # pixels per side of the height map
    write_cama: bool = True
    write_nuscenes: bool = True
    seed: int = 0


def _straightish_trajectory(n, speed=5.0, period_s=0.5):
    """+x at 5 m/s with yaw 0.3*sin(s/8); pose n//2 is the identity."""
    poses = []
    s = (np.arange(n) - n // 2) * speed * period_s
    # integrate heading so the path is smooth; centre pose pinned at the origin
    yaw = 0.3 * np.sin(s / 8.0)
    yaw = yaw - yaw[n // 2]
    xy = np.zeros((n, 2))
    for i in range(n // 2 + 1, n):
        ds = s[i] - s[i - 1]
        xy[i] = xy[i - 1] + ds * np.array([np.cos(yaw[i - 1]), np.sin(yaw[i - 1])])
    for i in range(n // 2 - 1, -1, -1):
        ds = s[i + 1] - s[i]
        xy[i] = xy[i + 1] - ds * np.array([np.cos(yaw[i]), np.sin(yaw[i])])
    for i in range(n):
        pose = np.eye(4)
        pose[:3, :3] = _rot_z(yaw[i])
        pose[:2, 3] = xy[i]
        pose[2, 3] = 0.02 * np.sin(s[i] / 5.0)
        poses.append(pose)
    return poses


def _circular_trajectory(n, radius=180.0):
    poses = []
    for i in range(n):
        ang = 2.0 * np.pi * i / n
        pose = np.eye(4)
        pose[:3, :3] = _rot_z(ang + np.pi / 2.0)
        pose[:2, 3] = radius * np.array([np.cos(ang), np.sin(ang)])
        poses.append(pose)
    return poses


def config1_spec():
    """BASELINE.json configs[0]: one frame, CAM_FRONT, one 50-vertex lane polyline (no densify:
    vertices are 1 m apart so the 0.1 m densify is irrelevant to the anchor counts only when the
    polyline is handed over pre-densified; here the raw polyline is written and densified as usual)."""
    xs = np.arange(5.0, 55.0, 1.0)
    line = np.stack([xs, np.full_like(xs, 1.8)], axis=1)
    poses = [np.eye(4), np.eye(4)]
    return ClipSpec(name="config1", n_frames=1, cameras=["camera_front"],
                    polylines_m=[("lane_marking", line)], chassis_poses=poses,
                    bev_size=6000, seed=0)


def config2_spec(n_frames=40, n_polylines=200, seed=0, pose_time_offset_ms=0, name="config2"):
    rng = np.random.default_rng(seed)
    lines = []
    for i in range(n_polylines):
        length = rng.uniform(20.0, 80.0)
        x0 = rng.uniform(-150.0, 180.0 - length)
        y0 = rng.uniform(-38.0, 38.0)
        amp = rng.uniform(0.0, 1.5)
        phase = rng.uniform(0.0, 2 * np.pi)
        t = np.linspace(0.0, 1.0, 50)
        xs = x0 + length * t
        ys = np.clip(y0 + amp * np.sin(2 * np.pi * t + phase), -40.0, 40.0)
        lines.append((MAP_CLASSES[i % 3], np.stack([xs, ys], axis=1)))
    poses = _straightish_trajectory(n_frames + 1)
    return ClipSpec(name=name, n_frames=n_frames, polylines_m=lines, chassis_poses=poses,
                    pose_time_offset_ms=pose_time_offset_ms, seed=seed)


def config3_spec(n_frames=320, n_polylines=1600, seed=1, name="config3"):
    rng = np.random.default_rng(seed)
    lines = []
    for i in range(n_polylines):
        r = rng.uniform(120.0, 240.0)
        a0 = rng.uniform(0.0, 2 * np.pi)
        arc = rng.uniform(20.0, 80.0) / r
        t = np.linspace(0.0, 1.0, 50)
        ang = a0 + arc * t
        lines.append((MAP_CLASSES[i % 3], np.stack([r * np.cos(ang), r * np.sin(ang)], axis=1)))
    poses = _circular_trajectory(n_frames + 1)
    return ClipSpec(name=name, n_frames=n_frames, polylines_m=lines, chassis_poses=poses, seed=seed)


def tiny_spec(n_frames=3, n_polylines=12, seed=7, pose_time_offset_ms=0, name="tiny"):
    """Small clip for golden fixtures and fast parity tests."""
    spec = config2_spec(n_frames=n_frames, n_polylines=n_polylines, seed=seed,
                        pose_time_offset_ms=pose_time_offset_ms, name=name)
    spec.bev_size = 1200
    # keep the polylines near the trajectory so every camera sees something
    rng = np.random.default_rng(seed + 100)
    lines = []
    for i in range(n_polylines):
        length = rng.uniform(8.0, 30.0)
        x0 = rng.uniform(-75.0, 75.0 - length)   # some vertices fall outside the +-50 m crop box
        y0 = rng.uniform(-9.0, 9.0)
        t = np.linspace(0.0, 1.0, 7 + i % 5)
        xs = x0 + length * t
        ys = y0 + 0.8 * np.sin(2 * np.pi * t + i)
        if i % 4 == 3:                       # a few lateral elements (crosswalk-like)
            xs, ys = x0 + 0.3 * np.sin(3 * t), y0 + (length / 3.0) * (t - 0.5)
        lines.append((MAP_CLASSES[i % 3], np.stack([xs, ys], axis=1)))
    # degenerate inputs the reference handles: a one-point polyline (dropped) and a
    # polyline with a sub-resolution segment (segment dropped)
    lines.append(("lane_marking", np.array([[1.0, 1.0]])))
    lines.append(("Road_teeth", np.array([[4.0, -2.0], [5.0, -2.0], [5.05, -2.0], [5.35, -2.0]])))
    spec.polylines_m = lines
    return spec


def write_clip(spec: ClipSpec, root: str, bev_dtype=np.float32) -> str:
    """Write ``spec`` under ``root/<name>`` and return the clip path."""
    clip = os.path.join(root, spec.name)
    os.makedirs(os.path.join(clip, "odometry"), exist_ok=True)
    os.makedirs(os.path.join(clip, "maps"), exist_ok=True)
    rng = np.random.default_rng(spec.seed + 12345)

    n_stamps = spec.n_frames + 1
    assert len(spec.chassis_poses) == n_stamps
    frame_ms = [_T0_MS + i * spec.frame_period_ms for i in range(n_stamps)]

    calibration = {}
    for cam in CAMERA_LIST:
        calibration[f"{cam}_2_chassis"] = camera_to_chassis(cam).tolist()
        calibration[cam] = {"K": camera_intrinsics(cam).tolist(), "d": [0.0] * 8,
                            "image_width": 1600, "image_height": 900}
    attribute = {
        "calibration": calibration,
        "sync": {cam: list(frame_ms) for cam in CAMERA_LIST},
        "unsync": {cam: list(frame_ms) for cam in CAMERA_LIST},
    }
    with open(os.path.join(clip, "attribute.json"), "w") as fh:
        json.dump(attribute, fh)

    # pose stamps: identical to the frame stamps (exact-hit branch of seek_by_timestamp) or
    # shifted so that every query lands between two poses (slerp branch).  One extra pose is
    # appended in the shifted case so the last frame still has a right neighbour.
    off = spec.pose_time_offset_ms
    poses = list(spec.chassis_poses)
    pose_ms = [t - off for t in frame_ms]
    if off != 0:
        extra = poses[-1].copy()
        extra[:3, 3] = 2 * poses[-1][:3, 3] - poses[-2][:3, 3]
        poses.append(extra)
        pose_ms.append(pose_ms[-1] + spec.frame_period_ms)
    pose_s = [t / 1000.0 for t in pose_ms]

    # nuScenes-label branch: chassis poses under an arbitrary global offset that
    # normalize2center() removes (reference cama/dataset.py:71-76)
    offset = np.eye(4)
    offset[:3, :3] = _rot_z(0.7)
    offset[:3, 3] = [1200.0, -830.0, 3.0]
    centre = poses[len(poses) // 2]
    centre_inv = np.linalg.inv(centre)
    np.savetxt(os.path.join(clip, "odometry", "wigo_offset_clip.txt"),
               _tum_rows(pose_s, [offset @ p for p in poses]), fmt="%.12f")
    # CAMA-label branch: camera_front -> world (reference cama/dataset.py:60-69)
    cam2chassis = camera_to_chassis("camera_front")
    np.savetxt(os.path.join(clip, "odometry", "scmv_camera_front.txt"),
               _tum_rows(pose_s, [p @ cam2chassis for p in poses]), fmt="%.12f")

    if spec.write_nuscenes:
        # labels live in the centre-normalised frame
        labels = []
        for cls, xy in spec.polylines_m:
            h = np.concatenate([xy, np.zeros((len(xy), 1)), np.ones((len(xy), 1))], axis=1)
            xy_c = (centre_inv @ h.T).T[:, :2]
            labels.append({"attrs": {"type": cls}, "data": np.round(xy_c, 4).tolist()})
        with open(os.path.join(clip, "maps", "map_nuscenes.json"), "w") as fh:
            json.dump(labels, fh)

    if spec.write_cama:
        # labels in BEV pixels: world x = p[1]*0.1-300, world y = p[0]*0.1-300
        labels = []
        for cls, xy in spec.polylines_m:
            p0 = (xy[:, 1] + _MAP_HALF) / _MAP_RES
            p1 = (xy[:, 0] + _MAP_HALF) / _MAP_RES
            labels.append({"attrs": {"type": cls}, "data": np.round(np.stack([p0, p1], 1), 3).tolist()})
        with open(os.path.join(clip, "maps", "map_labels.json"), "w") as fh:
            json.dump(labels, fh)
        bev = (rng.standard_normal((spec.bev_size, spec.bev_size)) * spec.height_sigma).astype(bev_dtype)
        np.save(os.path.join(clip, "maps", "vision_road_mlp_ft.npy"), bev)
    return clip


def write_background_jpegs(clip_path: str, n_frames: int, cameras=None, seed=0):
    """Optional 1600x900 JPEG per camera-frame (needs OpenCV, only used by image-path tests)."""
    import cv2
    rng = np.random.default_rng(seed)
    with open(os.path.join(clip_path, "attribute.json")) as fh:
        attribute = json.load(fh)
    for cam in (cameras or CAMERA_LIST):
        os.makedirs(os.path.join(clip_path, cam), exist_ok=True)
        for idx in range(n_frames + 1):
            yy, xx = np.mgrid[0:900, 0:1600]
            img = np.stack([(xx * 255 // 1600), (yy * 255 // 900),
                            np.full_like(xx, int(rng.integers(0, 255)))], axis=-1).astype(np.uint8)
            img[::50, :, :] = 255
            cv2.imwrite(os.path.join(clip_path, cam, f"{attribute['sync'][cam][idx]}.jpg"), img)


def lidar_to_chassis():
    """lidar_top mounted 1.84 m above the chassis origin, 0.94 m ahead, yawed -90 deg (nuScenes-like)."""
    T = np.eye(4)
    T[:3, :3] = _rot_z(-np.pi / 2)
    T[:3, 3] = [0.94, 0.0, 1.84]
    return T


def write_lidar_sweeps(clip_path: str, n_sweeps: int = 40, n_points: int = 35000, seed: int = 0, ragged: bool = False):
    """BASELINE.json configs[4] / SURVEY 8d config 5: ``n_sweeps`` sweeps of ``n_points`` points, x,y ~ U(-50,50),
    z ~ N(0,1), rows (x y z intensity ring time) float64 — written as ``lidar_top/<ms>.bin`` the way
    dataset/nuscenes2clip.py:550-554 writes them, with the calibration and timestamp entries
    DatasetReader needs.  Sweep stamps = the clip's frame stamps (the first ``n_sweeps`` of them)."""
    with open(os.path.join(clip_path, "attribute.json")) as fh:
        attribute = json.load(fh)
    stamps = list(attribute["sync"][CAMERA_LIST[0]])[:n_sweeps]
    assert len(stamps) == n_sweeps, "the clip has fewer frame stamps than sweeps asked for"
    attribute["calibration"]["lidar_top_2_chassis"] = lidar_to_chassis().tolist()
    attribute["sync"]["lidar_top"] = stamps
    attribute["unsync"]["lidar_top"] = stamps
    with open(os.path.join(clip_path, "attribute.json"), "w") as fh:
        json.dump(attribute, fh)
    folder = os.path.join(clip_path, "lidar_top")
    os.makedirs(folder, exist_ok=True)
    rng = np.random.default_rng(seed + 777)
    for k, ms in enumerate(stamps):
        n = n_points if not ragged else int(n_points * (0.25 + 1.5 * rng.random())) * (k % 5 != 3)
        pts = np.zeros((n, 6), dtype=np.float64)
        pts[:, 0:2] = rng.uniform(-50.0, 50.0, size=(n, 2))
        pts[:, 2] = rng.standard_normal(n)
        pts[:, 3] = rng.uniform(0.0, 255.0, size=n)
        pts.tofile(os.path.join(folder, f"{ms}.bin"))
    return stamps
