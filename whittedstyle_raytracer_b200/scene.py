"""Host-side scene handle: parses a config the way the reference's driver does
(PPMGenerator + main()'s bunny.obj load), builds the reference-identical BVH and
exposes the flattened WrtSceneDesc / WrtCamera for upload."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import cabi


class SceneError(RuntimeError):
    """Config/asset error; the message is what the reference prints after 'ERROR:'."""


class Scene:
    def __init__(self, config_path=None, *, text=None, obj_path=None, asset_dir=None, glass=False):
        self._lib = cabi.load_host()
        self._h = C.c_void_p()
        enc = lambda s: None if s is None else os.fspath(s).encode()
        variant = 1 if glass else 0
        if text is not None:
            rc = self._lib.wrt_scene_load_text(text.encode(), enc(obj_path), enc(asset_dir), variant, C.byref(self._h))
        else:
            rc = self._lib.wrt_scene_load(enc(config_path), enc(obj_path), enc(asset_dir), variant, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise SceneError(self._lib.wrt_host_last_error().decode(errors="replace"))

    @classmethod
    def from_workdir(cls, workdir, name, *, bunny=True, glass=False):
        """Loads `<workdir>/<name>.txt` with `<workdir>` as the cwd-equivalent (textures, bunny.obj)."""
        wd = Path(workdir)
        obj = wd / "bunny.obj" if bunny else None
        return cls(wd / f"{name}.txt", obj_path=obj, asset_dir=wd, glass=glass)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.wrt_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def desc(self) -> cabi.WrtSceneDesc:
        return self._lib.wrt_scene_desc(self._h).contents

    @property
    def desc_ptr(self):
        return self._lib.wrt_scene_desc(self._h)

    @property
    def camera(self) -> cabi.WrtCamera:
        return self._lib.wrt_scene_camera(self._h).contents

    @property
    def camera_ptr(self):
        return self._lib.wrt_scene_camera(self._h)

    @property
    def width(self) -> int:
        return self.camera.width

    @property
    def height(self) -> int:
        return self.camera.height

    @property
    def n_prims(self) -> int:
        return self.desc.n_prims

    @property
    def bvh_depth(self) -> int:
        return self._lib.wrt_scene_bvh_depth(self._h)

    @property
    def upload_bytes(self) -> int:
        return self._lib.wrt_scene_upload_bytes(self._h)

    @property
    def output_name(self) -> str:
        return self._lib.wrt_scene_output_name(self._h).decode()

    def set_imsize(self, w: int, h: int) -> None:
        if self._lib.wrt_scene_set_imsize(self._h, w, h) != 0:
            raise SceneError(self._lib.wrt_host_last_error().decode())

    def set_soft_shadows(self, soft: bool) -> None:
        self._lib.wrt_scene_set_shadow_type(self._h, 1 if soft else 0)

    def prim_object(self) -> np.ndarray:
        d = self.desc
        return np.ctypeslib.as_array(d.prim_object, shape=(d.n_prims,)).copy() if d.n_prims else np.zeros(0, np.int32)

    def nodes(self) -> np.ndarray:
        d = self.desc
        if d.n_nodes == 0:
            return np.zeros((0, 8), np.float32)
        return np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_float)), shape=(d.n_nodes, 8)).copy()


def write_ppm_p3(path, rgb: np.ndarray) -> None:
    """ASCII P3 writer with the reference's layout (PPMGenerator.hpp:631-646)."""
    lib = cabi.load_host()
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    if lib.wrt_write_ppm_p3(os.fspath(path).encode(), w, h, rgb.ctypes.data) != 0:
        raise OSError(lib.wrt_host_last_error().decode())


def read_ppm_p3(path) -> np.ndarray:
    tok = Path(path).read_text().split()
    assert tok[0] == "P3"
    w, h = int(tok[1]), int(tok[2])
    return np.array(tok[4:4 + w * h * 3], dtype=np.int64).reshape(h, w, 3)
