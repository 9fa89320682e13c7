"""ctypes binding of libcama_b200.so (include/cama_b200.h).

There is deliberately no CPU fallback: if the library is missing, cannot be loaded, or no sm_100
device is present, the calls raise.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

from . import build as _build

CAMA_OK = 0
CAMA_E_INVALID, CAMA_E_CUDA, CAMA_E_WORKSPACE, CAMA_E_CAPACITY, CAMA_E_NODEVICE, CAMA_E_UNSUPPORTED = -1, -2, -3, -4, -5, -6
VERTEX_F32X4, VERTEX_F64X3 = 0, 1
CLIP_AUTO, CLIP_PLANE, CLIP_BINNED = 0, 1, 2
MAX_CAMERAS = 8
TILE_VERTICES = 256
WARP_VERTICES = 32
MAX_PEERS = 8
PEER_HEADER_BYTES = 256
PEER_HANDLE_BYTES = 64
PHASE_ALL, PHASE_GEOMETRY, PHASE_RASTER = 0, 1, 2
CAMERA_TABLE_BYTES = 16 + 256 * 256
OVERLAY_BGR, OVERLAY_PALETTE = 0, 1
OVERLAY_RECORD_BYTES = {OVERLAY_BGR: 32, OVERLAY_PALETTE: 12}
OVERLAY_DRAW, OVERLAY_BLANK, OVERLAY_DRAW_CHUNKS, OVERLAY_BLANK_CHUNKS = 0, 1, 2, 3
CLIP_PHASES = 4
PHASE_NAMES = ("prep", "geometry", "lists", "raster")
ABI_VERSION = 3


class CamaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libcama_b200 error {code}: {message}")
        self.code = code


class CapacityError(CamaError):
    """A record list of a clip render overflowed; rerun with stats.record_capacity_needed."""


class ClipDesc(Structure):
    _fields_ = [
        ("struct_bytes", c_uint32), ("mode", c_int32),
        ("n_frames", c_int32), ("n_cams", c_int32), ("n_instances", c_int32),
        ("height", c_int32), ("width", c_int32), ("vertex_layout", c_int32),
        ("n_vertices", c_int64),
        ("vertices", c_void_p), ("vertex_instance", c_void_p), ("world2chassis", c_void_p),
        ("chassis2cam", POINTER(c_double)), ("intrinsics", POINTER(c_double)),
        ("crop_box", c_double * 6),
        ("instance_bgr", c_void_p), ("background", c_void_p), ("frames", c_void_p),
        ("crop_counts", c_void_p), ("visible_counts", c_void_p), ("vu_dense", c_void_p),
        ("record_capacity", c_int64),
        ("tile_bounds", c_void_p), ("warp_bounds", c_void_p),
        ("overlay_records", c_void_p), ("overlay_count", c_void_p), ("overlay_capacity", c_int64),
        ("overlay_format", c_int32), ("pipeline_frames", c_int32), ("instance_palette", c_void_p),
        ("overlay_mirrors", c_void_p * 8), ("overlay_n_mirrors", c_int32), ("reserved0", c_int32), ("overlay_image_base", c_int64),
        ("camera_table", c_void_p), ("geometry_ctas_per_sm", c_int32), ("raster_ctas_per_sm", c_int32),
        ("mosaic_cols", c_int32), ("mosaic_tile_of_cam", c_int32 * 8),
        ("phases", c_int32), ("list_frame_base", c_int32), ("list_frames", c_int32), ("reserved2", c_int32),
        ("list_records", c_void_p), ("list_cursor", c_void_p),
    ]


class OverlayTarget(Structure):
    _fields_ = [("pixels", c_void_p), ("n_frames", c_int64), ("n_cams", c_int32), ("height", c_int32), ("width", c_int32),
                ("grid_cols", c_int32), ("tile_of_cam", c_void_p)]


class VoxelGrid(Structure):
    _fields_ = [("origin", c_double * 3), ("voxel", c_double * 3), ("dims", c_int32 * 3), ("reserved", c_int32)]


class ClipStats(Structure):
    _fields_ = [
        ("records_total", c_int64), ("record_capacity_needed", c_int64), ("record_capacity", c_int64),
        ("overflow", c_int32), ("mode", c_int32), ("band_rows", c_int32), ("n_bands", c_int32),
        ("overlay_records", c_int64), ("lists_per_image", c_int32), ("reserved", c_int32),
    ]


# every symbol include/cama_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cama_abi_version": (c_int, []),
    "cama_last_error": (c_char_p, []),
    "cama_device_count": (c_int, [POINTER(c_int)]),
    "cama_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "cama_ctx_destroy": (c_int, [c_void_p]),
    "cama_ctx_launch_count": (c_int, [c_void_p, POINTER(c_uint64)]),
    "cama_ctx_sm_count": (c_int, [c_void_p, POINTER(c_int)]),
    "cama_ctx_profile_enable": (c_int, [c_void_p, c_int]),
    "cama_ctx_profile_calls": (c_int, [c_void_p, POINTER(c_int)]),
    "cama_ctx_profile_read": (c_int, [c_void_p, c_int, POINTER(ctypes.c_float)]),
    "cama_transform_points": (c_int, [c_void_p, c_void_p, c_int, c_int64, POINTER(c_double), c_void_p, c_void_p]),
    "cama_compact_workspace_bytes": (c_int, [c_int64, POINTER(c_size_t)]),
    "cama_crop_points": (c_int, [c_void_p, c_void_p, c_int, c_int64, POINTER(c_double), POINTER(c_double), c_void_p, c_int64,
                                 c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cama_project_points": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_double), POINTER(c_double), c_int, c_int, c_void_p,
                                    c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cama_render_workspace_bytes": (c_int, [c_int, c_int, POINTER(c_size_t)]),
    "cama_render_points": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                   c_size_t, c_void_p]),
    "cama_render_points_overlay": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64,
                                           c_void_p, c_size_t, c_void_p]),
    "cama_densify_plan": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_float, c_void_p, c_void_p]),
    "cama_densify_fill": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int] + [ctypes.c_float] * 5
                          + [c_void_p, c_void_p]),
    "cama_remap_bilinear": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "cama_camera_table_build": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double), c_int, POINTER(c_double), c_int, c_int, c_void_p, c_void_p]),
    "cama_clip_workspace_bytes": (c_int, [POINTER(ClipDesc), POINTER(c_size_t)]),
    "cama_clip_render": (c_int, [c_void_p, POINTER(ClipDesc), c_void_p, c_size_t, c_void_p]),
    "cama_clip_stats_read": (c_int, [c_void_p, POINTER(ClipDesc), c_void_p, c_void_p, POINTER(ClipStats)]),
    "cama_overlay_apply_host": (c_int, [c_void_p, c_int64, c_int, c_void_p, POINTER(OverlayTarget), c_int, c_int]),
    "cama_lidar_accumulate": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cama_overlay_fetch_apply": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, POINTER(OverlayTarget), c_int, c_int, c_void_p]),
    "cama_peer_slot_bytes": (c_int, [c_int64, c_int, POINTER(c_size_t)]),
    "cama_peer_alloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p), c_void_p]),
    "cama_peer_free": (c_int, [c_void_p, c_void_p]),
    "cama_peer_open": (c_int, [c_void_p, c_int, c_void_p, POINTER(c_void_p)]),
    "cama_peer_close": (c_int, [c_void_p, c_void_p]),
    "cama_peer_publish": (c_int, [c_void_p, c_void_p, c_uint32, POINTER(c_void_p), c_int, c_void_p]),
    "cama_peer_publish_cursors": (c_int, [c_void_p, c_void_p, c_int64, c_int64, POINTER(c_void_p), c_int, c_uint32, POINTER(c_void_p), c_int, c_void_p]),
    "cama_peer_publish_lists": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, POINTER(c_void_p), POINTER(c_void_p), c_int, c_uint32,
                                        POINTER(c_void_p), c_int, c_void_p]),
    "cama_peer_wait": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_uint32, c_int, c_void_p, c_void_p]),
    "cama_frames_clear": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "cama_peer_expand": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int, c_uint32, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                                 c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cama_host_bandwidth_probe": (c_int, [c_int64, c_int, POINTER(c_double), POINTER(c_double)]),
    "cama_overlay_expand": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
}

_LIB = None


def library_path():
    return _build.LIB_PATH


def lib():
    """Load (building first if the sources changed) and type the shared library."""
    global _LIB
    if _LIB is None:
        path = _build.ensure_built()
        handle = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)           # AttributeError here = header/library mismatch: fail loudly
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.cama_abi_version() != ABI_VERSION:
            raise CamaError(CAMA_E_INVALID, f"ABI version {handle.cama_abi_version()} != binding {ABI_VERSION}")
        _LIB = handle
    return _LIB


def check(code):
    if code == CAMA_OK:
        return
    msg = lib().cama_last_error().decode("utf-8", "replace")
    raise (CapacityError if code == CAMA_E_CAPACITY else CamaError)(code, msg)


def dptr(array_or_none):
    """ctypes double* of a contiguous float64 numpy array (or NULL)."""
    if array_or_none is None:
        return None
    return array_or_none.ctypes.data_as(POINTER(c_double))
