"""Clip reader: the boundary that feeds the reprojection path its calibration, stamps and poses.

Call-compatible with /root/reference/cama/dataset_reader.py (``DatasetReader``).  Only the parts
the frame loop touches matter for parity — ``get_extrinsic`` (:150-248), ``get_intrinsics``
(:278-294), ``get_sensor_timestamp`` (:39-43), ``get_odometry`` (:409-411); the sensor
generators are provided so that code written against the reference class keeps working.
Pure host-side file parsing: nothing here is worth a GPU.
"""
from __future__ import annotations

import json
import os
from collections import defaultdict, deque
from warnings import warn

import numpy as np

from .pose_transformer import invT


def _stamp_of(path):
    """'<ms>.<ext>' -> seconds."""
    return float(os.path.basename(path).split('.')[0]) / 1000.0


class DatasetReader:
    def __init__(self, pack_path=None):
        self.attribute = dict()
        self.extrinsic_graph = None
        self.pack_path = ""
        if pack_path:
            self.read_pack(pack_path)

    def read_pack(self, path):
        """Load ``<path>/attribute.json``; raises FileNotFoundError when it is missing."""
        self.pack_path = path
        attribute_path = os.path.join(path, "attribute.json")
        if not os.path.exists(attribute_path):
            raise FileNotFoundError("can not find {}".format(attribute_path))
        with open(attribute_path, 'r') as fh:
            self.attribute = json.load(fh)

    # ------------------------------------------------------------------ stamps / file lists
    def get_sensor_timestamp(self, sensor_name, sync=True):
        """Stamps of a sensor in seconds, as a python list of floats (stored as integer ms)."""
        stamps = np.asarray(self.attribute["sync" if sync else "unsync"][sensor_name]).astype(np.double)
        stamps /= 1000.0
        return stamps.tolist()

    def yield_sensor_filepath(self, sensor_name, ext, sync=True, start_idx=None, end_idx=None,
                              start_time=None, end_time=None):
        """Absolute paths ``<pack>/<sensor>/<ms>.<ext>``, selectable by index range or time range."""
        stamps = self.attribute["sync" if sync else "unsync"][sensor_name]
        seconds = np.asarray(stamps) / 1000.
        if start_time is None and end_time is None:
            chosen = stamps[start_idx:end_idx]
        else:
            if start_time is None or start_time <= seconds[0]:
                start_idx = None
            elif start_time > seconds[-1]:
                start_idx = -1
            else:
                start_idx = np.searchsorted(seconds, start_time, side="left")
            if end_time is None or end_time >= seconds[-1]:
                end_idx = None
            elif end_idx < seconds[0]:
                end_idx = -1
            else:
                end_idx = np.searchsorted(seconds, end_time, side="left") - 1
            chosen = [] if (start_idx < 0 or end_idx < 0) else stamps[start_idx:end_idx]
        folder = os.path.join(self.pack_path, sensor_name)
        for ms in chosen:
            yield os.path.join(folder, "{}.{}".format(ms, ext))

    def yield_lidar(self, start_idx=None, end_idx=None, deskewed=False):
        """(stamp, (n,6) float64: x y z intensity ring time) per sweep."""
        for name in self.yield_sensor_filepath("lidar_top", "bin", start_idx=start_idx, end_idx=end_idx):
            if deskewed:
                name = name.replace("lidar_top", "deskewed_lidar_top")
            yield _stamp_of(name), np.fromfile(name, dtype=np.double).reshape(-1, 6)

    def _yield_json_frames(self, folder, group, key):
        with open(os.path.join(self.pack_path, folder, "data.json"), 'r') as fh:
            frames = json.load(fh)
        for ms in self.attribute[group][key]:
            yield float(ms) / 1000.0, frames[str(ms)]

    def yield_IMU(self, start_idx=None, end_idx=None, start_time=None, end_time=None):
        yield from self._yield_json_frames("IMU", "unsync", "IMU")

    def yield_GNSS(self, start_idx=None, end_idx=None):
        yield from self._yield_json_frames("UB482", "unsync", "UB482")

    def yield_wheel(self, sync=True, start_idx=None, end_idx=None):
        yield from self._yield_json_frames("wheel", "sync" if sync else "unsync", "wheel")

    def yield_camera(self, camera="camera_front", start_idx=None, end_idx=None):
        import cv2
        for name in self.yield_sensor_filepath(camera, "jpg", start_idx=start_idx, end_idx=end_idx):
            yield _stamp_of(name), cv2.imread(name)

    def yield_semantic(self, camera="camera_front", start_idx=None, end_idx=None):
        import cv2
        for name in self.yield_sensor_filepath(camera, "png", start_idx=start_idx, end_idx=end_idx):
            name = name.replace(camera, "seg_" + camera)
            yield _stamp_of(name), cv2.imread(name, cv2.IMREAD_UNCHANGED)

    # ------------------------------------------------------------------ calibration
    def _direct_extrinsic(self, from_sensor, to_sensor):
        """Stored edge, its rigid inverse, the identity — or None when the two are not adjacent."""
        if from_sensor == to_sensor:
            return np.eye(4, dtype=np.float32)
        calib = self.attribute["calibration"]
        forward = "{}_2_{}".format(from_sensor, to_sensor)
        if forward in calib:
            return np.asarray(calib[forward])
        backward = "{}_2_{}".format(to_sensor, from_sensor)
        if backward in calib:
            return invT(np.asarray(calib[backward]))
        return None

    def _build_graph(self):
        graph = defaultdict(list)
        for key in self.attribute["calibration"]:
            if "_2_" in key:
                a, b = key.split('_2_')
                graph[a].append(b)
                graph[b].append(a)
        self.extrinsic_graph = graph

    def get_extrinsic_path(self, from_sensor, to_sensor):
        """Breadth-first sensor chain ``[from, ..., to]`` over the calibration edges, or None."""
        if self.extrinsic_graph is None:
            self._build_graph()
        if from_sensor == to_sensor:
            return None
        done = []
        frontier = deque([[from_sensor]])
        while frontier:
            chain = frontier.popleft()
            tail = chain[-1]
            if tail in done:
                continue
            for nxt in self.extrinsic_graph[tail]:
                grown = chain + [nxt]
                frontier.append(grown)
                if nxt == to_sensor:
                    return grown
            done.append(tail)
        return None

    def get_extrinsic(self, from_sensor, to_sensor):
        """4x4 taking coordinates in ``from_sensor`` to ``to_sensor``, chaining calibration edges
        along the shortest path when the pair is not stored directly; None (and a message) when
        the two sensors are not connected."""
        direct = self._direct_extrinsic(from_sensor, to_sensor)
        if direct is not None:
            return direct
        chain = self.get_extrinsic_path(from_sensor, to_sensor)
        if chain is None:
            print("extrinsic path not found!")
            return None
        total = np.eye(4, dtype=np.float32)
        for a, b in zip(chain[:-1], chain[1:]):
            total = self._direct_extrinsic(a, b) @ total
        return total

    def get_all_sensors(self):
        names = []
        for key in self.attribute["calibration"]:
            names += key.split('_2_')
        return list(set(names))

    def get_intrinsic(self, sensor):
        warn("get_intrinsic() is deprecated, use get_intrinsics() instead")
        entry = self.attribute["calibration"][sensor]
        return np.asarray(entry["K"]), np.asarray(entry["d"])

    def get_intrinsics(self, sensor):
        """{"K","d","width","height","hfov"} of a camera; missing entries are None."""
        entry = self.attribute["calibration"][sensor]
        return {"K": np.asarray(entry.get("K", None)), "d": np.asarray(entry.get("d", None)),
                "width": entry.get("image_width", None), "height": entry.get("image_height", None),
                "hfov": entry.get("fov", None)}

    # ------------------------------------------------------------------ odometry sources
    def get_odometry(self, name_txt):
        return np.loadtxt(os.path.join(self.pack_path, "odometry", name_txt))

    def get_GNSS_tum(self):
        """(N,8) TUM rows from the GNSS json (both the list- and the dict-valued layout)."""
        rows = []
        for t, frame in self.yield_GNSS():
            pos, ori = frame["position"], frame["orientation"]
            if "x" in pos:
                rows.append([t, pos["x"], pos["y"], pos["z"], ori["x"], ori["y"], ori["z"], ori["w"]])
            else:
                warn("Warning(Deprecation): clip/pack results extracted by packstreamer will not be supported in the future")
                rows.append([t, pos[0], pos[1], pos[2], ori[0], ori[1], ori[2], ori[3]])
        return np.asarray(rows)

    def get_wheel_tum(self, sync=False):
        """(N,8) TUM rows from wheel odometry (roll/pitch/yaw layout, or planar x/y/yaw layout)."""
        from scipy.spatial.transform import Rotation
        rows = []
        for t, frame in self.yield_wheel(sync=sync):
            if "roll" in frame:
                warn("Warning(Deprecation): clip/pack results extracted by packstreamer will not be supported in the future")
                q = Rotation.from_euler("XYZ", [frame["roll"], frame["pitch"], frame["yaw"]], degrees=False).as_quat()
                rows.append([t, frame["x"], frame["y"], frame["z"], q[0], q[1], q[2], q[3]])
            else:
                q = Rotation.from_euler("XYZ", [0, 0, frame["yaw"]], degrees=False).as_quat()
                rows.append([t, frame["x"], frame["y"], 0, q[0], q[1], q[2], q[3]])
        return np.asarray(rows)
