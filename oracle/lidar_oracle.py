"""TEST INFRASTRUCTURE — NumPy restatement of the LiDAR voxel aggregation (include/cama_b200.h,
cama_lidar_accumulate).  Only tests/, __graft_entry__.smoke() and bench/tools CPU baselines may import it.

PARITY UNPINNED: the reference snapshot has no voxel accumulation (it lives on the camav2 branch,
/root/reference/README.md:17-20).  What IS taken from the reference: the point transform
`(T @ P.T).T[:, :3]` in float64 (/root/reference/cama/reproject.py:108-116) on rows read as
/root/reference/cama/dataset_reader.py:45-51 reads them.  The voxel rule is this repo's definition:
index = floor((world - origin) / voxel) per axis, counted when inside the grid.
"""
import numpy as np


def transform_points(points_xyz, T):
    """cama/reproject.py:112-115: ones appended (float64), (T @ P.T).T[:, :3]."""
    pts = np.asarray(points_xyz, dtype=np.float64)
    h = np.concatenate([pts, np.ones((pts.shape[0], 1))], axis=1)
    return (np.asarray(T, dtype=np.float64) @ h.T).T[:, :3]


def accumulate(sweeps, transforms, origin, voxel, dims, counts=None):
    """-> (counts uint32 [nz,ny,nx], points counted)."""
    nx, ny, nz = (int(d) for d in dims)
    if counts is None:
        counts = np.zeros((nz, ny, nx), dtype=np.uint32)
    origin = np.asarray(origin, dtype=np.float64)
    voxel = np.asarray(voxel, dtype=np.float64)
    inside_total = 0
    for pts, T in zip(sweeps, transforms):
        pts = np.asarray(pts, dtype=np.float64)
        if pts.size == 0:
            continue
        pts = pts.reshape(len(pts), -1)
        world = transform_points(pts[:, :3], T)
        with np.errstate(invalid="ignore", over="ignore"):
            q = np.floor((world - origin) / voxel)
            ok = (q[:, 0] >= 0) & (q[:, 0] < nx) & (q[:, 1] >= 0) & (q[:, 1] < ny) & (q[:, 2] >= 0) & (q[:, 2] < nz)
        idx = q[ok].astype(np.int64)
        np.add.at(counts, (idx[:, 2], idx[:, 1], idx[:, 0]), np.uint32(1))
        inside_total += int(ok.sum())
    return counts, inside_total
