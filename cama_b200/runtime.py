"""Device runtime: one libcama_b200 context per GPU, torch tensors as the device-memory container.

Everything that touches the GPU goes through here.  There is no CPU fallback anywhere in the
package: without a CUDA device (or without the built library) these calls raise.
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _native as N

_RUNTIMES = {}
_LOCK = threading.Lock()

CROP_KEYS = ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")


def _torch():
    import torch
    return torch


class Runtime:
    """Context + scratch memory for one CUDA device."""

    def __init__(self, index):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("cama_b200 needs a CUDA device (B200, sm_100a); it has no CPU path")
        self.index = index
        self.device = torch.device("cuda", index)
        handle = ctypes.c_void_p()
        N.check(N.lib().cama_ctx_create(index, ctypes.byref(handle)))
        self.ctx = handle
        self._scratch = {}

    # -------------------------------------------------------------- plumbing
    def stream(self):
        """The torch current stream of this device as the void* the C ABI takes."""
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        _torch().cuda.current_stream(self.device).synchronize()

    def launches(self):
        n = ctypes.c_uint64()
        N.check(N.lib().cama_ctx_launch_count(self.ctx, ctypes.byref(n)))
        return int(n.value)

    def sm_count(self):
        n = ctypes.c_int()
        N.check(N.lib().cama_ctx_sm_count(self.ctx, ctypes.byref(n)))
        return int(n.value)

    def profile_enable(self, max_calls):
        """Record per-phase CUDA events for the next ``max_calls`` clip renders (0 = off)."""
        N.check(N.lib().cama_ctx_profile_enable(self.ctx, int(max_calls)))

    def profile_read(self):
        """-> float array [calls, CLIP_PHASES] of milliseconds (waits for the recorded calls)."""
        calls = ctypes.c_int()
        N.check(N.lib().cama_ctx_profile_calls(self.ctx, ctypes.byref(calls)))
        out = np.zeros((calls.value, N.CLIP_PHASES), dtype=np.float32)
        for i in range(calls.value):
            N.check(N.lib().cama_ctx_profile_read(self.ctx, i, out[i].ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
        return out

    def scratch(self, name, nbytes):
        """A cached uint8 device buffer of at least nbytes (grown geometrically, 512-B aligned by torch)."""
        torch = _torch()
        buf = self._scratch.get(name)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._scratch[name] = buf
        return buf

    def scratch_tensor(self, name, shape, dtype):
        """A cached device tensor of at least ``shape`` rows (first dimension may be larger)."""
        torch = _torch()
        buf = self._scratch.get(name)
        if buf is None or buf.dtype != dtype or tuple(buf.shape[1:]) != tuple(shape[1:]) or buf.shape[0] < shape[0]:
            buf = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._scratch[name] = buf
        return buf

    def to_device(self, array, dtype=None):
        torch = _torch()
        arr = np.ascontiguousarray(array, dtype=dtype)
        if arr.size == 0:
            return torch.empty(arr.shape, dtype=getattr(torch, str(arr.dtype)), device=self.device)
        return torch.from_numpy(arr).to(self.device, non_blocking=False)

    @staticmethod
    def ptr(tensor):
        return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None and tensor.numel() > 0 else ctypes.c_void_p(0)

    # -------------------------------------------------------------- load-time densify (scope row N2)
    def densify(self, polylines, resolution, bev_height=None, solution=0.1, half_width=300.0, half_height=300.0,
                center_x=0.0, center_y=0.0):
        """cama_densify_plan + cama_densify_fill.

        polylines   list of (k_i, 2) arrays, k_i >= 2 (label vertices; float32 after the reference's astype)
        bev_height  None (metric labels: z = 0) or float32 [rows, cols] height map (CAMA pixel labels)
        -> (vertices: torch float32 [total, 4] on the device in CAMA_VERTEX_F32X4 layout, per-polyline point counts)
        """
        torch = _torch()
        raw = np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1, 2) for p in polylines], axis=0) if polylines else np.zeros((0, 2), np.float32)
        sizes = np.array([len(p) for p in polylines], dtype=np.int64)
        raw_poly = np.repeat(np.arange(len(polylines), dtype=np.int32), sizes)
        n_raw = int(raw.shape[0])
        d_raw, d_poly = self.to_device(raw), self.to_device(raw_poly)
        d_start = torch.empty(n_raw + 1, dtype=torch.int64, device=self.device)
        N.check(N.lib().cama_densify_plan(self.ctx, self.ptr(d_raw), self.ptr(d_poly), n_raw, ctypes.c_float(float(np.float32(resolution))),
                                          ctypes.c_void_p(d_start.data_ptr()), self.stream()))
        seg_start = d_start.cpu().numpy()
        total = int(seg_start[-1])
        first_raw = np.concatenate([[0], np.cumsum(sizes)])
        counts = seg_start[first_raw[1:]] - seg_start[first_raw[:-1]] if len(polylines) else np.zeros(0, np.int64)
        verts = torch.empty((total, 4), dtype=torch.float32, device=self.device)
        d_bev = None
        rows = cols = 0
        if bev_height is not None:
            assert bev_height.dtype == np.float32 and bev_height.ndim == 2
            d_bev = self.to_device(bev_height)
            rows, cols = bev_height.shape
        f = lambda v: ctypes.c_float(float(np.float32(v)))
        N.check(N.lib().cama_densify_fill(self.ctx, self.ptr(d_raw), self.ptr(d_poly), n_raw, ctypes.c_void_p(d_start.data_ptr()), total,
                                          self.ptr(d_bev) if d_bev is not None else None, rows, cols, f(solution), f(half_width), f(half_height),
                                          f(center_x), f(center_y), self.ptr(verts), self.stream()))
        return verts, counts

    # -------------------------------------------------------------- per-call operators (numpy in / numpy out)
    def transform_points(self, points, T):
        """cama_transform_points: (n,3) float32/float64 -> (n,3) float64 = (T @ [p;1])[:3]."""
        torch = _torch()
        pts = np.ascontiguousarray(points)
        if pts.dtype != np.float32:
            pts = pts.astype(np.float64, copy=False)
        n = pts.shape[0]
        T64 = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        d_in = self.to_device(pts)
        d_out = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        N.check(N.lib().cama_transform_points(self.ctx, self.ptr(d_in), int(pts.dtype == np.float32), n, N.dptr(T64),
                                              self.ptr(d_out), self.stream()))
        return d_out.cpu().numpy()

    def _compact_scratch(self, n):
        need = ctypes.c_size_t()
        N.check(N.lib().cama_compact_workspace_bytes(n, ctypes.byref(need)))
        return self.scratch("compact", need.value), need.value

    def crop_points(self, flat, offsets, box6, T=None):
        """cama_crop_points on a flat instance list -> (survivors (m,3) f64, new offsets)."""
        torch = _torch()
        pts = np.ascontiguousarray(flat)
        is_f32 = pts.dtype == np.float32
        if not is_f32:
            pts = pts.astype(np.float64, copy=False)
        if is_f32 and T is None:
            pts, is_f32 = pts.astype(np.float64), False
        n, n_inst = pts.shape[0], len(offsets) - 1
        d_in = self.to_device(pts)
        d_off = self.to_device(offsets, np.int64)
        d_out = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        d_out_off = torch.empty(n_inst + 1, dtype=torch.int64, device=self.device)
        ws, ws_bytes = self._compact_scratch(n)
        T64 = None if T is None else np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        box = np.ascontiguousarray(box6, dtype=np.float64)
        N.check(N.lib().cama_crop_points(self.ctx, self.ptr(d_in), int(is_f32), n, N.dptr(T64), N.dptr(box), self.ptr(d_off), n_inst,
                                         self.ptr(d_out), self.ptr(d_out_off), self.ptr(ws), ws_bytes, self.stream()))
        out_off = d_out_off.cpu().numpy()
        return d_out[:int(out_off[-1])].cpu().numpy(), out_off

    def crop_points_resident(self, d_in, is_f32, d_off, n_inst, box6, T):
        """cama_crop_points on points that are already on the device (a dataset's dense vertices, uploaded once).
        -> (survivors device [m,3] f64, device offsets [n_inst+1], host offsets)"""
        torch = _torch()
        n = int(d_in.shape[0])
        d_out = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        d_out_off = torch.empty(n_inst + 1, dtype=torch.int64, device=self.device)
        ws, ws_bytes = self._compact_scratch(n)
        T64 = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        box = np.ascontiguousarray(box6, dtype=np.float64)
        N.check(N.lib().cama_crop_points(self.ctx, self.ptr(d_in), int(is_f32), n, N.dptr(T64), N.dptr(box), self.ptr(d_off), n_inst,
                                         self.ptr(d_out), self.ptr(d_out_off), self.ptr(ws), ws_bytes, self.stream()))
        out_off = d_out_off.cpu().numpy()
        return d_out[:int(out_off[-1])], d_out_off, out_off

    def project_points_cameras(self, d_pts, d_off, n_inst, cameras, width, height):
        """cama_project_points for several cameras on the same device-resident points: every camera's launches are
        enqueued back to back, then ONE synchronisation brings all (v,u) arrays and offsets back.
        cameras: list of (K 3x3, T 4x4).  -> list of ((k,2) f64 host array, host offsets) per camera."""
        torch = _torch()
        n, n_cams = int(d_pts.shape[0]), len(cameras)
        if n == 0:
            return [(np.zeros((0, 2)), np.zeros(n_inst + 1, np.int64), None, None) for _ in cameras]
        d_vu = torch.empty((n_cams, n, 2), dtype=torch.float64, device=self.device)
        d_offs = torch.empty((n_cams, n_inst + 1), dtype=torch.int64, device=self.device)
        need = ctypes.c_size_t()
        N.check(N.lib().cama_compact_workspace_bytes(n, ctypes.byref(need)))
        ws = self.scratch("compact_cams", need.value * n_cams)          # one slice per camera: the launches of different cameras are in flight together
        for c, (K, T) in enumerate(cameras):
            T64 = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
            K64 = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
            N.check(N.lib().cama_project_points(self.ctx, self.ptr(d_pts), n, N.dptr(T64), N.dptr(K64), int(width), int(height),
                                                self.ptr(d_off), n_inst, ctypes.c_void_p(d_vu[c].data_ptr()), ctypes.c_void_p(d_offs[c].data_ptr()),
                                                ctypes.c_void_p(ws.data_ptr() + c * need.value), need.value, self.stream()))
        offs = d_offs.cpu().numpy()                                  # (synchronises)
        most = int(offs[:, -1].max())
        vu = d_vu[:, :most].cpu().numpy() if most else np.zeros((n_cams, 0, 2))
        # (the device copies ride along: render_maps on the list this becomes does not upload the points again)
        return [(vu[c, :int(offs[c, -1])], offs[c], d_vu[c, :int(offs[c, -1])], d_offs[c]) for c in range(n_cams)]

    def project_points(self, flat, offsets, K, width, height, T=None):
        """cama_project_points -> ((k,2) f64 (v,u), new offsets)."""
        torch = _torch()
        pts = np.ascontiguousarray(flat, dtype=np.float64)
        n, n_inst = pts.shape[0], len(offsets) - 1
        d_in = self.to_device(pts)
        d_off = self.to_device(offsets, np.int64)
        d_out = torch.empty((n, 2), dtype=torch.float64, device=self.device)
        d_out_off = torch.empty(n_inst + 1, dtype=torch.int64, device=self.device)
        ws, ws_bytes = self._compact_scratch(n)
        T64 = None if T is None else np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        K64 = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        N.check(N.lib().cama_project_points(self.ctx, self.ptr(d_in), n, N.dptr(T64), N.dptr(K64), int(width), int(height),
                                            self.ptr(d_off), n_inst, self.ptr(d_out), self.ptr(d_out_off), self.ptr(ws), ws_bytes,
                                            self.stream()))
        out_off = d_out_off.cpu().numpy()
        return d_out[:int(out_off[-1])].cpu().numpy(), out_off

    def render_points(self, image, vu_flat, offsets, inst_bgr, device_points=None):
        """render_maps on a host image (uint8 HxWx3 numpy): stamps in place and returns it.
        ``device_points``: (device (k,2) float64 (v,u), device int64 offsets) when the points are on the device already.

        The image itself never crosses PCIe when its width is a multiple of 8 (and it is writeable and contiguous):
        cama_render_points_overlay returns the lit 8-pixel chunks (~160 KB for a 540x960 frame of config 2) and
        cama_overlay_apply_host draws them into ``image``.  Otherwise: upload, cama_render_points, download."""
        torch = _torch()
        assert image.dtype == np.uint8 and image.ndim == 3 and image.shape[2] == 3, "image must be uint8 [H,W,3]"
        n_inst = len(offsets) - 1
        if device_points is not None:
            d_vu, d_off = device_points
            n = int(d_vu.shape[0])
        else:
            vu = np.ascontiguousarray(vu_flat, dtype=np.float64)
            n = vu.shape[0]
        if n == 0:
            return image
        height, width = image.shape[:2]
        if device_points is None:
            d_vu = self.to_device(vu)
            d_off = self.to_device(offsets, np.int64)
        d_bgr = self._cached_bgr(inst_bgr)
        need = ctypes.c_size_t()
        N.check(N.lib().cama_render_workspace_bytes(height, width, ctypes.byref(need)))
        ws = self.scratch("render", need.value)
        if width % 8 == 0 and image.flags.writeable and image.flags.c_contiguous:
            capacity = height * (width // 8)
            records = self.scratch("render_records", capacity * 32)
            count = self.scratch_tensor("render_count", (4,), torch.int32)
            N.check(N.lib().cama_render_points_overlay(self.ctx, self.ptr(d_vu), n, self.ptr(d_off), n_inst, self.ptr(d_bgr), height, width,
                                                       self.ptr(records), self.ptr(count), capacity, self.ptr(ws), need.value, self.stream()))
            lit = int(count[:1].cpu().item())                            # (synchronises)
            if lit:
                host = self._pinned_records(lit)
                host[:lit * 32].copy_(records[:lit * 32])
                target = N.OverlayTarget(image.ctypes.data, 1, 1, height, width, 0, None)
                N.check(N.lib().cama_overlay_apply_host(host.data_ptr(), lit, N.OVERLAY_BGR, None, ctypes.byref(target), N.OVERLAY_DRAW, 1))
            return image
        d_img = self.to_device(image)
        N.check(N.lib().cama_render_points(self.ctx, self.ptr(d_vu), n, self.ptr(d_off), n_inst, self.ptr(d_bgr), self.ptr(d_img),
                                           height, width, self.ptr(ws), need.value, self.stream()))
        result = d_img.cpu().numpy()
        if image.flags.writeable:
            image[...] = result            # the reference draws in place (cv2.circle mutates its argument)
            return image
        return result

    def _cached_bgr(self, inst_bgr):
        """instance colours on the device; the same array (by content) is uploaded once"""
        arr = np.ascontiguousarray(inst_bgr, dtype=np.uint8)
        key = arr.tobytes()
        hit = self._scratch.get("bgr_cache")
        if hit is None or hit[0] != key:
            hit = self._scratch["bgr_cache"] = (key, self.to_device(arr, np.uint8))
        return hit[1]

    def _pinned_records(self, n_records):
        torch = _torch()
        buf = self._scratch.get("render_records_host")
        if buf is None or buf.numel() < n_records * 32:
            buf = torch.empty(max(n_records * 32, 1 << 20), dtype=torch.uint8, pin_memory=True)
            self._scratch["render_records_host"] = buf
        return buf


def get_runtime(device=None) -> Runtime:
    """Runtime of ``device`` (int, torch.device or None = current CUDA device)."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("cama_b200 needs a CUDA device (B200, sm_100a); it has no CPU path")
    if device is None:
        index = torch.cuda.current_device()
    elif isinstance(device, int):
        index = device
    else:
        dev = torch.device(device)
        index = dev.index if dev.index is not None else torch.cuda.current_device()
    with _LOCK:
        if index not in _RUNTIMES:
            _RUNTIMES[index] = Runtime(index)
        return _RUNTIMES[index]
