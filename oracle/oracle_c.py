"""ctypes binding of oracle/liboracle.so (plain-C oracle) — TEST INFRASTRUCTURE, NOT PRODUCT.

See oracle.c for what each entry point restates (reference file:line) and how parity is pinned.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.orc_transform_f64.argtypes = [vp, vp, i64, vp]
        L.orc_transform_f32.argtypes = [vp, vp, i64, vp]
        L.orc_crop_mask.argtypes = [vp, i64, vp, vp]
        L.orc_project.argtypes = [vp, vp, i64, i32, i32, vp, vp]
        L.orc_stamp.argtypes = [vp, i32, i32, vp, i64, vp]
        L.orc_clip_render.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, i64, vp, i32, i32, vp, vp, vp, i64, i64]
        L.orc_densify_f32.argtypes = [vp, i64, vp]
        L.orc_densify_f32.restype = i64
        L.orc_pixel_to_world_f32.argtypes = [vp, i64, vp, i64, i64, vp]
        for fn in (L.orc_transform_f64, L.orc_transform_f32, L.orc_crop_mask, L.orc_project, L.orc_stamp,
                   L.orc_clip_render, L.orc_pixel_to_world_f32):
            fn.restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def transform(T, pts):
    T = np.ascontiguousarray(T, np.float64)
    out = np.empty((len(pts), 3), np.float64)
    if pts.dtype == np.float32:
        pts = np.ascontiguousarray(pts)
        lib().orc_transform_f32(_p(T), _p(pts), len(pts), _p(out))
    else:
        pts = np.ascontiguousarray(pts, np.float64)
        lib().orc_transform_f64(_p(T), _p(pts), len(pts), _p(out))
    return out


def crop_mask(pts, box6):
    pts = np.ascontiguousarray(pts, np.float64)
    box = np.ascontiguousarray(box6, np.float64)
    keep = np.empty(len(pts), np.uint8)
    lib().orc_crop_mask(_p(pts), len(pts), _p(box), _p(keep))
    return keep.astype(bool)


def project(K, pts, width, height):
    K = np.ascontiguousarray(K, np.float64)
    pts = np.ascontiguousarray(pts, np.float64)
    vu = np.empty((len(pts), 2), np.float64)
    keep = np.empty(len(pts), np.uint8)
    lib().orc_project(_p(K), _p(pts), len(pts), width, height, _p(vu), _p(keep))
    return vu, keep.astype(bool)


def stamp(image, vu, bgr):
    assert image.dtype == np.uint8 and image.flags.c_contiguous and image.shape[2] == 3
    vu = np.ascontiguousarray(vu, np.float64)
    col = np.ascontiguousarray(bgr, np.uint8)
    lib().orc_stamp(_p(image), image.shape[0], image.shape[1], _p(vu), len(vu), _p(col))
    return image


def clip_render(verts, inst_offsets, inst_bgr, w2c, c2cam, K, box6, height, width,
                frames=None, want_counts=True, frame_range=None):
    verts = np.ascontiguousarray(verts, np.float32)
    offs = np.ascontiguousarray(inst_offsets, np.int64)
    bgr = np.ascontiguousarray(inst_bgr, np.uint8)
    w2c = np.ascontiguousarray(w2c, np.float32).reshape(-1, 16)
    c2cam = np.ascontiguousarray(c2cam, np.float64).reshape(-1, 16)
    K = np.ascontiguousarray(K, np.float64).reshape(-1, 9)
    box = np.ascontiguousarray(box6, np.float64)
    F, C, I = len(w2c), len(c2cam), len(offs) - 1
    if frames is None:
        frames = np.zeros((F, C, height, width, 3), np.uint8)
    crop_counts = np.zeros((F, I), np.int32) if want_counts else None
    vis_counts = np.zeros((F, C, I), np.int32) if want_counts else None
    lo, hi = frame_range if frame_range is not None else (0, F)
    lib().orc_clip_render(_p(verts), _p(offs), I, _p(bgr), _p(w2c), F, _p(c2cam), _p(K), C, _p(box),
                          height, width, _p(frames), _p(crop_counts), _p(vis_counts), lo, hi)
    return frames, crop_counts, vis_counts


def densify(poly_xy):
    poly = np.ascontiguousarray(poly_xy, np.float32)
    n = lib().orc_densify_f32(_p(poly), len(poly), None)
    out = np.empty((n, 2), np.float32)
    lib().orc_densify_f32(_p(poly), len(poly), _p(out))
    return out


def pixel_to_world(dense_xy, bev):
    dense = np.ascontiguousarray(dense_xy, np.float32)
    bev = np.ascontiguousarray(bev, np.float32)
    out = np.empty((len(dense), 3), np.float32)
    lib().orc_pixel_to_world_f32(_p(dense), len(dense), _p(bev), bev.shape[0], bev.shape[1], _p(out))
    return out
