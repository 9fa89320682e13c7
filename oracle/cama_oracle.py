"""CPU ORACLE for the CAMA per-frame reprojection path — TEST INFRASTRUCTURE, NOT PRODUCT.

A NumPy/SciPy/OpenCV restatement of the algorithm in the reference
(manymuch/CAMA @ 5033cb8, mounted read-only at /root/reference in the build
container).  Every function cites the reference lines it restates.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this module; the product package
``cama_b200`` never does (tests/test_native_abi.py::test_product_never_imports_the_oracle
enforces that).

Parity pinning: the reference itself has no tests or golden vectors
(SURVEY.md section 4), so this oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF, produced in the build container by tests/golden/make_golden.py
(which imports /root/reference unmodified) and committed under tests/golden/.
tests/test_oracle_golden.py checks every function here against those fixtures
bit-for-bit (float arrays with ``array_equal``; see the note on n==1 below).

Third-party arithmetic the reference leans on and that is therefore part of the
contract (versions of this image: numpy 2.3.5 + OpenBLAS 0.3.30, scipy 1.18.1,
opencv 4.13.0; the reference pins none of them, requirements.txt:1-7):

* ``T @ P.T`` goes to BLAS dgemm, which accumulates the K<=4 products of each
  output element in index order with fused multiply-adds starting from the first
  product; ``oracle.c`` and the CUDA kernels use exactly that FMA chain.  For a
  ONE-point instance NumPy dispatches to gemv instead, whose kernel sums in a
  different order without FMA — a platform detail, not part of the algorithm; the
  difference is <= 1 ulp and is covered by the 1e-5 coordinate tolerance.
* NEP-50 scalar promotion keeps the densify arithmetic in float32.
* ``cv2.circle(img, c, 2, col, -1)`` paints the 13 pixels with |dx|+|dy| <= 2,
  clipped at the image border.
"""
from __future__ import annotations

import json
import os

import numpy as np

OUTPUT_HW = (540, 960)                                 # reference cama/reproject.py:164
CROP_BOX = {"x_min": -50, "x_max": 50, "y_min": -100,   # reference cama/reproject.py:28-34
            "y_max": 100, "z_min": -200, "z_max": 200}
RESOLUTION = 0.1                                        # reference cama/reproject.py:23
MAP_EXTENT_M = 600                                      # reference cama/reproject.py:26-27
# class -> RGB, reference cama/reproject.py:13-16; render forces every class except
# lane_marking to Crosswalk_Line (cama/reproject.py:251-253)
CLASS_RGB = {"Road_teeth": (235, 73, 127), "lane_marking": (211, 211, 211),
             "Stop_Line": (211, 211, 211), "Crosswalk_Line": (255, 215, 0)}


def class_colour_arrays():
    """class -> RGB array, built afresh on every call like BaseManager.get_color_maps (cama/reproject.py:11-17)."""
    return {k: np.array(v) for k, v in CLASS_RGB.items()}


# --------------------------------------------------------------------------- pose algebra
def inv_rigid(T):
    """[R t; 0 1]^-1 = [R^T, -R^T t]; float64 result.  Reference cama/pose_transformer.py:8-21."""
    Rt = T[:3, :3].T
    out = np.eye(4)
    out[:3, :3] = Rt
    out[:3, 3] = -Rt @ T[:3, 3]
    return out


def slerp_rigid(T_left, T_right, ratio):
    """Rotation by scipy Slerp, everything else (translation, bottom row) by lerp.
    Reference cama/pose_transformer.py:24-44."""
    from scipy.spatial.transform import Rotation, Slerp
    assert 0 <= ratio <= 1
    keys = Rotation.from_matrix(np.stack([T_left[:3, :3], T_right[:3, :3]]))
    rot = Slerp([0, 1], keys)(ratio).as_matrix()
    out = T_left * (1 - ratio) + T_right * ratio
    out[:3, :3] = rot
    return out


def tum_to_poses(tum):
    """(F,8) [t x y z qx qy qz qw] -> (stamps (F,), list of F 4x4 f64).
    Reference cama/pose_transformer.py:429-438."""
    from scipy.spatial.transform import Rotation
    assert tum.shape[1] == 8
    poses = np.zeros((tum.shape[0], 4, 4))
    poses[:, 3, 3] = 1
    poses[:, :3, :3] = Rotation.from_quat(tum[:, 4:8]).as_matrix()
    poses[:, :3, 3] = tum[:, 1:4]
    return tum[:, 0].copy(), list(poses)


def seek_pose(stamps, poses, query_time, t_max_diff, interpolate):
    """Reference cama/pose_transformer.py:589-652 (raises RuntimeError exactly where it does)."""
    assert isinstance(query_time, float) and isinstance(t_max_diff, float)
    assert np.all(stamps[1:] >= stamps[:-1])
    hit = np.where(np.isclose(stamps, query_time, rtol=1e-20, atol=1e-9))[0]
    if hit.size > 0:
        return poses[hit[0]]
    right = int(np.searchsorted(stamps, query_time, side="left"))
    left = right - 1
    if interpolate:
        if right >= len(stamps):
            raise RuntimeError("query_time is out of range.")
        if right == 0 and -1e-9 < (query_time - stamps[0]) < 0:
            right, left = 1, 0
        elif query_time - stamps[0] < -1e-9:
            raise RuntimeError("query_time is out of range.")
        gap = stamps[right] - stamps[left]
        if gap > t_max_diff:
            raise RuntimeError("time gap exceeds t_max_diff")
        return slerp_rigid(poses[left], poses[right], (query_time - stamps[left]) / gap)
    d_left = query_time - stamps[left] if left >= 0 else float("inf")
    d_right = stamps[right] - query_time if right < len(stamps) else float("inf")
    if min(d_left, d_right) > t_max_diff:
        raise RuntimeError("nearest pose too far")
    return poses[left if d_left < d_right else right]


# --------------------------------------------------------------------------- clip reading
def read_attribute(clip_path):
    """Reference cama/dataset_reader.py:19-37."""
    path = os.path.join(clip_path, "attribute.json")
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    with open(path) as fh:
        return json.load(fh)


def chassis_to_camera(attribute, camera):
    """get_extrinsic("chassis", cam): only the inverse-name entry exists in clips, so the
    reference returns invT(camera_2_chassis).  Reference cama/dataset_reader.py:150-168."""
    calib = attribute["calibration"]
    direct = f"chassis_2_{camera}"
    if direct in calib:
        return np.asarray(calib[direct])
    return inv_rigid(np.asarray(calib[f"{camera}_2_chassis"]))


def scaled_intrinsics(attribute, camera, out_hw=OUTPUT_HW):
    """Reference cama/reproject.py:171-182 (row-wise rescale of K to the 540x960 output)."""
    info = attribute["calibration"][camera]
    K = np.asarray(info["K"]).copy()
    K[0, :] = K[0, :] * out_hw[1] / info["image_width"]
    K[1, :] = K[1, :] * out_hw[0] / info["image_height"]
    return K


def frame_stamps_s(attribute, camera):
    """Reference cama/dataset_reader.py:39-43."""
    t = np.asarray(attribute["sync"][camera]).astype(np.double)
    t /= 1000.0
    return t.tolist()


# --------------------------------------------------------------------------- load-time geometry
def _densify(poly_xy):
    """float32 polyline -> dense float32 points.  Reference cama/reproject.py:49-63 / 79-93.

    Per segment: num = int(len / 0.1) evaluated in float32 (NEP 50); segment skipped when
    num == 0; points start + (end-start)/num*j for j = 0..num-1, so the segment end (and the
    polyline's last vertex) is never emitted.
    """
    pts = np.asarray(poly_xy).astype(np.float32)
    seg_len = np.linalg.norm(pts[1:] - pts[:-1], axis=-1)
    out = []
    for s in range(len(seg_len)):
        num = int(seg_len[s] / RESOLUTION)
        for j in range(num):
            out.append(pts[s] + (pts[s + 1] - pts[s]) / num * j)
    return np.array(out)


def instances_from_metric_labels(labels):
    """nuScenes-style labels (metres, z=0).  Reference cama/reproject.py:42-70."""
    result = []
    for item in labels:
        if len(item["data"]) <= 1:
            continue
        dense = _densify(item["data"])
        xyz = np.concatenate((dense, np.zeros_like(dense[:, 0])[:, None]), axis=-1).reshape(-1, 3)
        result.append({"class": item["attrs"]["type"], "points": xyz})
    return result


def instances_from_pixel_labels(bev_height, labels):
    """CAMA labels (BEV pixels) + height map.  Reference cama/reproject.py:72-106, 36-40."""
    result = []
    for item in labels:
        if len(item["data"]) <= 1:
            continue
        dense = _densify(item["data"])
        cell = dense.round().astype(np.uint16)[:, ::-1].clip(0, bev_height.shape[0] - 1)
        height = bev_height[cell[:, 0], cell[:, 1]]
        world = np.zeros_like(dense)
        world[:, 0] = dense[:, 1] * RESOLUTION - MAP_EXTENT_M / 2 + 0
        world[:, 1] = dense[:, 0] * RESOLUTION - MAP_EXTENT_M / 2 + 0
        xyz = np.concatenate((world, height[:, None]), axis=-1).reshape(-1, 3)
        result.append({"class": item["attrs"]["type"], "points": xyz})
    return result


# --------------------------------------------------------------------------- per-frame geometry
def transform_instances(instances, T):
    """Homogeneous 4x4 applied to every instance; float64 out.  Reference cama/reproject.py:108-116."""
    out = []
    for inst in instances:
        p = inst["points"]
        ph = np.concatenate((p, np.ones((p.shape[0], 1))), axis=-1)
        out.append({"class": inst["class"], "points": (T @ ph.T).T[:, :3]})
    return out


def crop_instances(instances, box=None):
    """Inclusive axis-aligned box; empty instances are dropped.  Reference cama/reproject.py:118-131."""
    box = CROP_BOX if box is None else box
    out = []
    for inst in instances:
        p = inst["points"]
        keep = ((p[:, 0] >= box["x_min"]) & (p[:, 0] <= box["x_max"]) &
                (p[:, 1] >= box["y_min"]) & (p[:, 1] <= box["y_max"]) &
                (p[:, 2] >= box["z_min"]) & (p[:, 2] <= box["z_max"]))
        p = p[keep]
        if p.shape[0] > 0:
            out.append({"class": inst["class"], "points": p})
    return out


def project_instances(instances, K, width, height):
    """Pinhole projection + per-point visibility mask; returns (v,u) float64; empty instances
    dropped.  Reference cama/reproject.py:187-205."""
    out = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for inst in instances:
            q = (K @ inst["points"].T).T
            in_front = q[:, 2] > 0
            q = q[:, :] / q[:, 2:]
            keep = ((q[:, 2] > 0) & (q[:, 0] >= 0) & (q[:, 0] < width) &
                    (q[:, 1] >= 0) & (q[:, 1] < height)) & in_front
            q = q[keep]
            if q.shape[0] > 0:
                out.append({"class": inst["class"], "points": q[:, :2][:, ::-1]})
    return out


def render_instances(image, instances_vu):
    """In-place painter's loop of radius-2 filled discs, BGR.  Reference cama/reproject.py:246-257.

    The loop has the reference's exact form — rows of the int32 array handed to ``cv2.circle`` as NumPy
    scalars, the colour tuple built from the class's RGB array per instance — because this function is also
    what ``bench.py --impl reference`` times: a different spelling of the same loop costs up to 1.8x more
    per point (tools/port_vs_reference.py keeps the two within 5 % per phase)."""
    import cv2
    for inst in instances_vu:
        points = inst["points"].astype(np.int32)
        cls = inst["class"] if inst["class"] == "lane_marking" else "Crosswalk_Line"
        colour = tuple(class_colour_arrays()[cls][::-1].tolist())
        for point in points:
            cv2.circle(image, (point[1], point[0]), 2, colour, -1)
    return image


# --------------------------------------------------------------------------- the clip loop
class ClipOracle:
    """The frame x camera loop.  Reference cama/dataset.py:11-126 (ClipManager)."""

    def __init__(self, configs, clip_path):
        self.configs = configs
        self.clip_path = clip_path
        self.attribute = read_attribute(clip_path)
        self.cameras = list(configs["camera_list"])
        self.K = [scaled_intrinsics(self.attribute, c) for c in self.cameras]
        self.chassis2cam = [chassis_to_camera(self.attribute, c) for c in self.cameras]
        self.instance_maps = {}
        maps_dir = os.path.join(clip_path, configs["result_dir"])
        cama_json = os.path.join(maps_dir, configs["cama_map_file"])
        if os.path.exists(cama_json):                                  # cama/dataset.py:26-41
            with open(cama_json) as fh:
                labels = json.load(fh)
            bev = np.load(os.path.join(maps_dir, configs["height_mlp"]))
            self.instance_maps["cama"] = instances_from_pixel_labels(bev, labels)
        nus_json = os.path.join(maps_dir, configs["nuscenes_map_file"])
        if os.path.exists(nus_json):                                   # cama/dataset.py:43-51
            with open(nus_json) as fh:
                self.instance_maps["nuscenes"] = instances_from_metric_labels(json.load(fh))

    def chassis_trajectory(self, dataset):
        """-> (stamps, chassis->world poses).  Reference cama/dataset.py:60-76."""
        odo = os.path.join(self.clip_path, "odometry")
        if dataset == "nuscenes":
            stamps, poses = tum_to_poses(np.loadtxt(os.path.join(odo, "wigo_offset_clip.txt")))
            centre_inv = inv_rigid(poses[len(poses) // 2])             # pose_transformer.py:324-336
            return stamps, [centre_inv @ p for p in poses]
        main = self.configs["camera_main"]
        stamps, poses = tum_to_poses(np.loadtxt(
            os.path.join(odo, f"{self.configs['pose_prefix']}_{main}.txt")))
        chassis2cam = chassis_to_camera(self.attribute, main)
        return stamps, [p @ chassis2cam for p in poses]                # pose_transformer.py:520-537

    def world_to_chassis_per_frame(self, dataset):
        """-> list of (image_idx, float32 4x4).  Reference cama/dataset.py:86-99: index 0 skipped,
        frames whose seek raises RuntimeError skipped, float32 cast BEFORE np.linalg.inv."""
        stamps, poses = self.chassis_trajectory(dataset)
        out = []
        frame_t = frame_stamps_s(self.attribute, self.configs["camera_main"])
        for image_idx in range(1, len(frame_t)):
            try:
                c2w = seek_pose(stamps, poses, frame_t[image_idx], 0.5, True).astype(np.float32)
            except RuntimeError:
                continue
            out.append((image_idx, np.linalg.inv(c2w)))
        return out

    def frames(self, dataset):
        """Generator of (image_idx, chassis-frame cropped instances).  cama/dataset.py:78-106."""
        for image_idx, w2c in self.world_to_chassis_per_frame(dataset):
            yield image_idx, crop_instances(transform_instances(self.instance_maps[dataset], w2c))

    def project_all(self, chassis_instances):
        """-> {camera: vu instances}.  Reference cama/dataset.py:108-117."""
        h, w = OUTPUT_HW
        return {cam: project_instances(transform_instances(chassis_instances, E), K, w, h)
                for cam, E, K in zip(self.cameras, self.chassis2cam, self.K)}

    def read_resized_image(self, cam, image_idx):
        """Camera image of a frame, undistort-resized to the output size.  Reference cama/reproject.py:207-215,
        228-244: cv2.imread, then initUndistortRectifyMap (recomputed on every call, like the reference:
        d is all zeros in a clip, so the map is a pure affine resample) and a bilinear cv2.remap."""
        import cv2
        info = self.attribute["calibration"][cam]
        stamp = self.attribute["sync"][cam][image_idx]
        image = cv2.imread(os.path.join(self.clip_path, cam, f"{stamp}.jpg"))
        h, w = OUTPUT_HW
        K_origin = np.asarray(info["K"])
        mapx, mapy = cv2.initUndistortRectifyMap(K_origin, np.asarray(info["d"]), None, self.K[self.cameras.index(cam)], (w, h), cv2.CV_32FC1)
        return cv2.remap(image, mapx, mapy, interpolation=cv2.INTER_LINEAR)

    def render_vectors(self, per_cam, image_idx):
        """-> {camera: image}: the frame's camera images with the map drawn in place.  cama/dataset.py:119-126."""
        return {cam: render_instances(self.read_resized_image(cam, image_idx), per_cam[cam]) for cam in self.cameras}

    def render_clip(self, dataset, backgrounds=None):
        """All frames on blank (or given) backgrounds -> (image_idx list, uint8 [F,C,H,W,3])."""
        h, w = OUTPUT_HW
        idx, out = [], []
        for k, (image_idx, chassis_instances) in enumerate(self.frames(dataset)):
            per_cam = self.project_all(chassis_instances)
            imgs = []
            for c, cam in enumerate(self.cameras):
                img = np.zeros((h, w, 3), np.uint8) if backgrounds is None else backgrounds[k, c].copy()
                imgs.append(render_instances(img, per_cam[cam]))
            idx.append(image_idx)
            out.append(np.stack(imgs))
        return idx, (np.stack(out) if out else np.zeros((0, len(self.cameras), h, w, 3), np.uint8))


# --------------------------------------------------------------------------- flat helpers for tests
def flatten(instances, ncols):
    pts = [np.asarray(i["points"]) for i in instances]
    flat = np.concatenate(pts, 0) if pts else np.zeros((0, ncols))
    return np.ascontiguousarray(flat), [i["class"] for i in instances], [len(p) for p in pts]
