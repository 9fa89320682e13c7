"""Pose algebra of the reprojection path (host side, NumPy/SciPy).

Call-compatible with /root/reference/cama/pose_transformer.py (``invT``, ``SlerpTransform``,
``PoseTransformer`` with its whole public surface).  A clip has at most a few hundred 4x4 poses,
and the float32 ``np.linalg.inv`` that follows them in the frame loop has to be bit-identical to
the reference's, so this stays on the host by design (DESIGN.md "What runs where"); the batched
device path consumes the per-frame world->chassis matrices produced from here.

Semantics that matter for parity, all taken from the reference:
* every result is float64; ``invT`` never calls a general inverse (pose_transformer.py:8-21);
* ``SlerpTransform`` slerps the rotation and *lerps the rest of the matrix* (:24-44);
* ``seek_by_timestamp`` demands python floats, treats stamps within 1e-9 s as exact hits and
  raises ``RuntimeError`` for out-of-range / too-sparse queries (:589-652) — the frame loop
  turns that exception into "skip this frame" (cama/dataset.py:90-96).
"""
from __future__ import annotations

from datetime import datetime
from warnings import warn

import numpy as np
from scipy.spatial.transform import Rotation, Slerp

__all__ = ["invT", "SlerpTransform", "PoseTransformer"]


def invT(transform):
    """Inverse of a rigid 4x4 as [R^T | -R^T t] (reference pose_transformer.py:8-21)."""
    rot_t = transform[:3, :3].T
    out = np.eye(4)
    out[:3, :3] = rot_t
    out[:3, 3] = -rot_t @ transform[:3, 3]
    return out


def SlerpTransform(transform_left, transform_right, ratio):
    """Pose between two poses: rotation by spherical interpolation, every other entry by linear
    interpolation (reference pose_transformer.py:24-44)."""
    assert 0 <= ratio <= 1, "ratio must between 0 to 1"
    assert transform_left.shape == transform_right.shape == (4, 4), "transform must be ndarray with 4x4"
    ends = Rotation.from_matrix(np.stack((transform_left[:3, :3], transform_right[:3, :3])))
    blended = transform_left * (1 - ratio) + transform_right * ratio
    blended[:3, :3] = Slerp([0, 1], ends)(ratio).as_matrix()
    return blended


def _need_batch_of(array, width, what):
    msg = f"{what} must be np.array in shape [B, {width}]"
    assert len(array.shape) == 2, msg
    assert array.shape[1] == width, msg


class PoseTransformer:
    """Trajectory container: absolute poses (sensor_i -> world), relative poses
    (inv(T_{i+1}) @ T_i) and their stamps, convertible between rotation representations."""

    def __init__(self, euler_order="ZXY", degree=False):
        self.euler_order = euler_order
        self.degree = degree
        self.reset()

    def reset(self):
        """Forget every pose and stamp."""
        self.relative_rotation = []      # N-1 x (3,3)
        self.relative_translation = []   # N-1 x (3,) or (3,1)
        self.relative_transform = []     # N-1 x (4,4)
        self.absolute_transform = []     # N   x (4,4)
        self.timestamps = []             # (N,1) once loaded

    # ------------------------------------------------------------------ internal state changes
    def _relative_from_parts(self):
        # reference :174-181 — appends, one inverted [R|t] per stored (rotation, translation) pair
        assert len(self.relative_rotation) == len(self.relative_translation)
        for rot, trans in zip(self.relative_rotation, self.relative_translation):
            step = np.eye(4, dtype=np.float64)
            step[:3, :3] = rot
            step[:3, 3] = trans
            self.relative_transform.append(invT(step))

    def _relative_from_absolute(self):
        # reference :183-196
        count = len(self.absolute_transform)
        if count == 0:
            raise RuntimeError("please load absolute first,                by using loadtxt()")
        self.relative_transform, self.relative_rotation, self.relative_translation = [], [], []
        for k in range(count - 1):
            step = invT(self.absolute_transform[k + 1]) @ self.absolute_transform[k]
            self.relative_transform.append(step)
            self.relative_rotation.append(step[:3, :3])
            self.relative_translation.append(step[:3, 3:])

    def _absolute_from_relative(self):
        # reference :198-207 — chain from the identity
        if len(self.relative_transform) == 0:
            self._relative_from_parts()
        assert len(self.relative_transform) > 0
        chain = [np.eye(4, dtype=np.float64)]
        for step in self.relative_transform:
            chain.append(chain[-1] @ step)
        self.absolute_transform = chain

    def _ensure_absolute(self):
        if len(self.absolute_transform) == 0:
            self._absolute_from_relative()

    def _require_loaded(self):
        if len(self.relative_transform) == 0 and len(self.absolute_transform) == 0:
            raise RuntimeError("please load data first!")

    def _apply_to_absolute(self, fn):
        self._ensure_absolute()
        self.absolute_transform = [fn(pose) for pose in self.absolute_transform]

    # ------------------------------------------------------------------ loaders
    def from_relative_transform(self, transform_array):
        assert transform_array.shape[1] == 4
        assert transform_array.shape[2] == 4
        self.relative_transform = transform_array
        self.absolute_transform = []

    def from_absolute_transform(self, transform_array):
        assert transform_array.shape[1] == 4
        assert transform_array.shape[2] == 4
        self.absolute_transform = transform_array
        self._relative_from_parts()

    def from_axis_angle(self, axis_angles, absolute):
        """(B,3) rotation vectors, global (``absolute``) or frame-to-frame."""
        if absolute:
            self.from_absolute_axis_angle(axis_angles)
        else:
            self.from_relative_axis_angle(axis_angles)

    def from_relative_axis_angle(self, axis_angles):
        _need_batch_of(axis_angles, 3, "axis_angles")
        self.absolute_transform = []
        self.relative_rotation = [Rotation.from_rotvec(v).as_matrix() for v in axis_angles]

    def _absolute_block(self, count, dtype, what):
        """Existing absolute poses as one (B,4,4) array, or B identities when nothing is loaded."""
        if len(self.absolute_transform) == 0:
            return np.tile(np.eye(4, dtype=dtype)[np.newaxis, :, :], (count, 1, 1))
        assert len(self.absolute_transform) == count, \
            f"previous stored absolute transform number not matched with input {what}"
        return np.asarray(self.absolute_transform)

    def from_absolute_axis_angle(self, axis_angles):
        _need_batch_of(axis_angles, 3, "axis_angles")
        rotations = Rotation.from_rotvec(axis_angles).as_matrix()
        block = self._absolute_block(rotations.shape[0], rotations.dtype, "axis angles")
        block[:, :3, :3] = rotations
        self.absolute_transform = list(block)

    def from_absolute_translation(self, translations):
        _need_batch_of(translations, 3, "translations")
        block = self._absolute_block(translations.shape[0], translations.dtype, "translations")
        block[:, :3, 3] = translations
        self.absolute_transform = list(block)

    def from_relative_quaternion(self, quaternions):
        _need_batch_of(quaternions, 4, "quaternions")
        self.absolute_transform = []
        self.relative_rotation = [Rotation.from_quat(q).as_matrix() for q in quaternions]

    def from_relative_eulers(self, eulers):
        self.absolute_transform = []
        self.relative_rotation = [
            Rotation.from_euler(seq=self.euler_order, angles=e, degrees=self.degree).as_matrix() for e in eulers]

    def from_translation(self, translations, absolute):
        """(B,3) translations, global (``absolute``) or frame-to-frame."""
        if absolute:
            self.from_absolute_translation(translations)
        else:
            self.from_relative_translation(translations)

    def from_relative_translation(self, translations):
        self.absolute_transform = []
        self.relative_translation = [t for t in translations]

    def load_timestamp(self, timestamps, style="unix", relative=True):
        if style == "unix":
            self._set_unix_stamps(timestamps)
        elif style == "kitti":
            # 'YYYY-mm-dd HH:MM:SS.fffffffff' — nanosecond digits beyond the microseconds are cut
            self._set_unix_stamps([datetime.strptime(t[:-4], '%Y-%m-%d %H:%M:%S.%f').timestamp() for t in timestamps])
        else:
            raise NotImplementedError(
                "style {} not supported yet.\nCurrently support [unix(tum), kitti]".format(style))

    def _set_unix_stamps(self, timestamps):
        if isinstance(timestamps, list):
            self.timestamps = np.expand_dims(np.asarray(timestamps), axis=-1)
            return
        assert timestamps.shape[0] > 0
        if timestamps.ndim == 1:
            self.timestamps = np.expand_dims(timestamps, axis=-1)
        elif timestamps.ndim == 2:
            self.timestamps = timestamps
        else:
            raise RuntimeError("input timestamp shape {} incorrect!".format(timestamps.shape))

    def loadarray(self, array, style="tum"):
        """Poses (+ stamps) from an evo-style array: ``tum`` (N,8: t x y z qx qy qz qw),
        ``kitti`` (N,12: row-major 3x4) or ``asl`` (N,17, EuRoC)."""
        self.reset()
        if style == "tum":
            assert array.shape[1] == 8
            self.timestamps = array[:, 0:1]
            self._load_rt(Rotation.from_quat(array[:, 4:8]).as_matrix(), array[:, 1:4])
        elif style == "kitti":
            assert array.shape[1] == 12
            count = array.shape[0]
            last_row = np.zeros((count, 1, 4))
            last_row[:, :, -1] = 1
            self.absolute_transform = np.concatenate((array.reshape(-1, 3, 4), last_row), axis=1)
            self._relative_from_absolute()
        elif style == "asl":
            assert array.shape[1] == 17
            stamps = array[:, 0] * 1e-9                    # ns -> s
            xyzw = array[:, [5, 6, 7, 4]]                  # stored w x y z
            self._load_rt(Rotation.from_quat(xyzw).as_matrix(), array[:, 1:4])
            self.timestamps = np.expand_dims(np.array(stamps), axis=1)
        else:
            raise NotImplementedError(
                "style {} not supported yet.\nCurrently support [tum, kitit, asl]".format(style))

    def _load_rt(self, rotations, translations):
        block = np.zeros((rotations.shape[0], 4, 4))
        block[:, 3, 3] = 1
        block[:, :3, :3] = rotations
        block[:, :3, 3] = translations
        self.absolute_transform = list(block)
        self._relative_from_absolute()

    # ------------------------------------------------------------------ exporters
    def as_quaternions(self, absolute=True):
        self._ensure_absolute()
        if not absolute:
            raise NotImplementedError("sorry, not yet supported :-(")
        return [Rotation.from_matrix(pose[:3, :3]).as_quat() for pose in self.absolute_transform]

    def _rotation_export(self, absolute, convert):
        self._require_loaded()
        if absolute:
            self._ensure_absolute()
            return convert(Rotation.from_matrix(np.asarray(self.absolute_transform)[:, :3, :3]))
        if len(self.relative_transform) == 0:
            self._relative_from_absolute()
        rows = [[convert(Rotation.from_matrix(step[:3, :3]))] for step in self.relative_transform]
        return np.concatenate(rows, axis=0)

    def as_euler(self, absolute):
        return self._rotation_export(absolute, lambda r: r.as_euler(seq=self.euler_order, degrees=self.degree))

    def as_axis_angle(self, absolute):
        return self._rotation_export(absolute, lambda r: r.as_rotvec())

    def as_axisangle(self, absolute):
        warn("Warning(Deprecation): as_axisangle is renamed to as_axis_angle, please consider update")
        return self.as_axis_angle(absolute=absolute)

    def as_translations(self, absolute):
        self._require_loaded()
        if absolute:
            self._ensure_absolute()
            return np.asarray([pose[:3, 3] for pose in self.absolute_transform])
        if len(self.relative_transform) == 0:
            self._relative_from_absolute()
        return np.concatenate([[step[:3, 3]] for step in self.relative_transform], axis=0)

    def as_trans_quat(self, absolute=True):
        quats = np.asarray(self.as_quaternions(absolute=absolute))
        trans = np.asarray(self.as_translations(absolute=absolute))
        return np.concatenate((trans, quats), axis=1)

    def as_transform(self, absolute=True):
        """(B,4,4) poses: world-referenced when ``absolute`` else frame-to-frame."""
        if not absolute:
            return np.asarray(self.relative_transform)
        self._ensure_absolute()
        return np.asarray(self.absolute_transform)

    def dumparray(self, style="tum"):
        if style != "tum":
            raise NotImplementedError(
                "style {} not supported yet.\nCurrently support [tum]".format(style))
        if len(self.relative_transform) == 0 and len(self.absolute_transform) == 0 and len(self.relative_translation) == 0:
            raise RuntimeError("No poses found, pleas load poses first")
        if self.timestamps.shape[0] == 0:
            raise RuntimeError("No timestamps found, pleas load timestamps first")
        self._ensure_absolute()
        n_stamps, n_poses = self.timestamps.shape[0], len(self.absolute_transform)
        if n_stamps + 1 == n_poses:
            self.absolute_transform = self.absolute_transform[1:]      # stamps describe poses 1..N
        elif n_stamps != n_poses:
            raise RuntimeError(
                "num of timestamps = {} while num of absolute transform = {}\n".format(n_stamps, n_poses) +
                "they should be equal or num of timestamps +1 = num of absolute transform")
        return np.concatenate((self.timestamps, self.as_trans_quat(absolute=True)), axis=1)

    def get_timestamps(self):
        if len(self.timestamps) == 0:
            raise RuntimeError("please load timestamps first, from loadtxt()")
        return self.timestamps

    # ------------------------------------------------------------------ whole-trajectory edits
    def normalize2origin(self):
        """Re-express the trajectory relative to its first pose."""
        self._ensure_absolute()
        anchor = invT(self.absolute_transform[0])
        self.absolute_transform = [anchor @ pose for pose in self.absolute_transform]

    def normalize2center(self):
        """Re-express the trajectory relative to its middle pose (index len//2)."""
        self._ensure_absolute()
        anchor = invT(self.absolute_transform[len(self.absolute_transform) // 2])
        self.absolute_transform = [anchor @ pose for pose in self.absolute_transform]

    def rotate(self, extrinsic):
        """Deprecated spelling of :meth:`right_rotate`."""
        warn("Warning(Deprecation): rotate function may lead misunderstanding\nPlease consider using transform()")
        assert extrinsic.shape == (4, 4)
        self._apply_to_absolute(lambda pose: pose @ extrinsic)

    def left_rotate(self, extrinsic):
        """T_i <- extrinsic @ T_i."""
        assert extrinsic.shape == (4, 4)
        self._apply_to_absolute(lambda pose: extrinsic @ pose)

    def right_rotate(self, extrinsic):
        """T_i <- T_i @ extrinsic (camera->world times chassis->camera gives chassis->world)."""
        assert extrinsic.shape == (4, 4)
        self._apply_to_absolute(lambda pose: pose @ extrinsic)

    def transform(self, extrinsic):
        """Move the trajectory from sensor A to sensor B: T_i <- X @ T_i @ inv(X), X = A->B."""
        assert extrinsic.shape == (4, 4)
        self._apply_to_absolute(lambda pose: extrinsic @ pose @ invT(extrinsic))

    def sort_by_timestamps(self):
        n_stamps = self.timestamps.shape[0]
        if n_stamps < 2:
            raise RuntimeError("there are only {} timestamps".format(n_stamps))
        order = np.argsort(self.timestamps[:, 0])
        if len(self.absolute_transform) == n_stamps:
            self.absolute_transform = list(np.asarray(self.absolute_transform)[order])
            self.timestamps = self.timestamps[order]
            return
        if n_stamps == len(self.relative_rotation) and n_stamps == len(self.relative_translation):
            self._relative_from_parts()
        elif n_stamps != len(self.relative_transform):
            raise NotImplementedError("whooops! not supported yet")
        if n_stamps != len(self.relative_transform):
            raise RuntimeError("# of timestamps = {} but # relative transform = {}".format(n_stamps, len(self.relative_transform)))
        self.relative_transform = list(np.asarray(self.relative_transform)[order])
        self.timestamps = self.timestamps[order]

    # ------------------------------------------------------------------ lookup
    def seek_by_timestamp(self, query_time: float, t_max_diff: float, interpolate=False):
        """Pose at ``query_time`` (python float seconds).

        ``interpolate=True``: the query must lie inside the stamped range and the two stamps
        around it may be at most ``t_max_diff`` apart; the result is ``SlerpTransform`` of the
        two neighbours.  ``interpolate=False``: nearest stamped pose, at most ``t_max_diff``
        away.  A stamp within 1e-9 s of the query short-circuits both modes.  Anything else
        raises ``RuntimeError``.
        """
        assert isinstance(query_time, float), f"query_time must be float, not {type(query_time)}"
        assert isinstance(t_max_diff, float), f"t_max_diff must be float, not {type(t_max_diff)}"
        if len(self.relative_transform) == 0 and len(self.absolute_transform) == 0 and len(self.relative_translation) == 0:
            raise RuntimeError("No poses found, pleas load poses first")
        if self.timestamps.shape[0] == 0:
            raise RuntimeError("No timestamps found, pleas load timestamps first")
        self._ensure_absolute()
        stamps = self.timestamps[:, 0]
        assert np.all(stamps[1:] >= stamps[:-1]), "timestamps must be sorted"

        exact = np.where(np.isclose(stamps, query_time, rtol=1e-20, atol=1e-9))[0]
        if exact.size > 0:
            return self.absolute_transform[exact[0]]

        hi = np.searchsorted(stamps, query_time, side="left")
        lo = hi - 1
        if not interpolate:
            gap_lo = query_time - self.timestamps[lo] if lo >= 0 else float("inf")
            gap_hi = self.timestamps[hi] - query_time if hi < stamps.shape[0] else float("inf")
            nearest = min(gap_lo, gap_hi)[0]
            if nearest > t_max_diff:
                raise RuntimeError(f"time_diff = {nearest} is greater than t_max_diff {t_max_diff}")
            return self.absolute_transform[lo if gap_lo < gap_hi else hi]

        if hi >= stamps.shape[0]:
            raise RuntimeError("query_time is out of range.")
        before_first = query_time - self.timestamps[0]
        if hi == 0 and -1e-9 < before_first < 0:
            lo, hi = 0, 1
        elif before_first < -1e-9:
            raise RuntimeError("query_time is out of range.")
        span = self.timestamps[hi] - self.timestamps[lo]
        if span > t_max_diff:
            raise RuntimeError(f"time_diff = {span} is greater than t_max_diff {t_max_diff}")
        return SlerpTransform(self.absolute_transform[lo], self.absolute_transform[hi], (query_time - self.timestamps[lo]) / span)
