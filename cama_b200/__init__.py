"""cama_b200 — B200-native (sm_100a) implementation of CAMA's per-frame map-reprojection path.

The module layout mirrors the reference package ``cama`` (manymuch/CAMA) so that code written
against it keeps working:

    cama.dataset.ClipManager                      -> cama_b200.dataset.ClipManager
    cama.reproject.{MapManager, CameraManager}    -> cama_b200.reproject.{...}
    cama.pose_transformer.{PoseTransformer, ...}  -> cama_b200.pose_transformer.{...}
    cama.dataset_reader.DatasetReader             -> cama_b200.dataset_reader.DatasetReader
    cama.tools.{VideoGenerator, load_json}        -> cama_b200.tools.{...}

``install_as_cama()`` registers these modules under the ``cama`` names, which makes the
reference's unmodified ``main.py`` (``from cama.dataset import ClipManager`` ...) run on this
package.  ``cama_b200.batched.Reproject`` is the batched whole-clip entry point.

Importing the package touches neither CUDA nor the native library; the first device call loads
``_lib/libcama_b200.so`` (building it with nvcc if the sources changed) and raises if there is no
sm_100 GPU — there is no CPU fallback.
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"

_SUBMODULES = ("pose_transformer", "dataset_reader", "tools", "reproject", "dataset")


def install_as_cama(force=False):
    """Alias this package's modules as ``cama`` / ``cama.<module>`` in ``sys.modules``."""
    if "cama" in sys.modules and not force and getattr(sys.modules["cama"], "__name__", "") != __name__:
        raise RuntimeError("a different 'cama' package is already imported; pass force=True to shadow it")
    package = sys.modules[__name__]
    sys.modules["cama"] = package
    for name in _SUBMODULES:
        sys.modules[f"cama.{name}"] = importlib.import_module(f"{__name__}.{name}")
    return package


def __getattr__(name):
    lazy = {"ClipManager": "dataset", "MapManager": "reproject", "CameraManager": "reproject",
            "PoseTransformer": "pose_transformer", "DatasetReader": "dataset_reader", "Reproject": "batched",
            "ClipRenderer": "batched"}
    if name in lazy:
        return getattr(importlib.import_module(f"{__name__}.{lazy[name]}"), name)
    raise AttributeError(name)
