"""The C-ABI library: builds for sm_100a, loads, exports every symbol the header declares, and
refuses to run without a GPU (no silent fallback).  No compute happens here."""
import ctypes
import os
import re

import pytest

from conftest import REPO
from cama_b200 import _native as N
from cama_b200 import build as B


def header_symbols():
    text = open(os.path.join(REPO, "include", "cama_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cama_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    path = B.ensure_built()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    names = header_symbols()
    assert len(names) >= 16
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/cama_b200.h but not exported"
    assert sorted(N.SIGNATURES) == names, "ctypes binding and header disagree"
    assert N.lib().cama_abi_version() == N.ABI_VERSION


def test_sass_is_sm100a_with_bulk_copy_and_fp64():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", B.ensure_built()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("DFMA", "UBLKCP", "VIMNMX3.U16x2", "ATOMS.CAS", "MUFU.RCP", "MATCH.ANY", "REDUX"):
        assert mnemonic in sass, mnemonic


def test_descriptor_layout_and_workspace_query():
    d = N.ClipDesc()
    d.struct_bytes = ctypes.sizeof(N.ClipDesc)
    d.mode = N.CLIP_AUTO
    d.n_frames, d.n_cams, d.n_instances, d.height, d.width = 40, 6, 200, 540, 960
    d.vertex_layout, d.n_vertices = N.VERTEX_F32X4, 100000
    need = ctypes.c_size_t()
    N.check(N.lib().cama_clip_workspace_bytes(ctypes.byref(d), ctypes.byref(need)))
    binned = need.value
    lists = 40 * 6 * 34                                  # 34 bands of 16 rows, one record list each
    assert lists * 8192 * 4 <= binned < lists * 8192 * 4 + (1 << 22)       # default capacity: 8192 4-byte records per list
    d.record_capacity = 5000
    N.check(N.lib().cama_clip_workspace_bytes(ctypes.byref(d), ctypes.byref(need)))
    assert lists * 5000 * 4 <= need.value < lists * 5000 * 4 + (1 << 22)
    d.record_capacity = 0
    d.mode = N.CLIP_PLANE
    N.check(N.lib().cama_clip_workspace_bytes(ctypes.byref(d), ctypes.byref(need)))
    assert need.value >= 40 * 6 * 540 * 960 * 4
    d.struct_bytes = 8
    assert N.lib().cama_clip_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == N.CAMA_E_INVALID
    assert b"size mismatch" in N.lib().cama_last_error()
    d.struct_bytes = ctypes.sizeof(N.ClipDesc)
    d.mode, d.width = N.CLIP_BINNED, 962          # rows not a multiple of 16 bytes: binned mode refuses
    assert N.lib().cama_clip_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == N.CAMA_E_UNSUPPORTED
    d.mode = N.CLIP_AUTO
    N.check(N.lib().cama_clip_workspace_bytes(ctypes.byref(d), ctypes.byref(need)))   # falls back to PLANE


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    assert N.lib().cama_ctx_create(0, ctypes.byref(ctx)) == N.CAMA_E_NODEVICE
    from cama_b200.runtime import get_runtime
    with pytest.raises(RuntimeError, match="no CPU path"):
        get_runtime()
    from cama_b200.reproject import MapManager
    import numpy as np
    with pytest.raises(RuntimeError):
        MapManager().transform_3d_instance_maps([{"class": "lane_marking", "points": np.zeros((2, 3), np.float32)}], np.eye(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "cama_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "liboracle" not in text, f
