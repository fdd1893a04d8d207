"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see oracle/blackstar_oracle.h).  Nothing under
blackstar_b200/ imports it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


class OrcCamera(ctypes.Structure):
    _fields_ = [("pos", ctypes.c_double * 3), ("look_at", ctypes.c_double * 3),
                ("up", ctypes.c_double * 3), ("fov", ctypes.c_double)]


class OrcScene(ctypes.Structure):
    _fields_ = [("step_size", ctypes.c_double), ("bloom_strength", ctypes.c_double),
                ("star_intensity", ctypes.c_double), ("star_saturation", ctypes.c_double),
                ("disk_hsi", ctypes.c_double * 3), ("disk_opacity", ctypes.c_double),
                ("disk_inner", ctypes.c_double), ("disk_outer", ctypes.c_double),
                ("bloom_divider", ctypes.c_int32), ("width", ctypes.c_int32),
                ("height", ctypes.c_int32), ("supersampling", ctypes.c_int32)]


ORC_STAR_DTYPE = np.dtype([("pos", "<f8", (3,)), ("hue", "<f8"), ("sat", "<f8"),
                           ("mag", "<i4"), ("pad_", "<i4")])


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc -O2 -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("blackstar_oracle.c", "oracle_thirdparty.c", "blackstar_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        L.orc_hsi_to_rgb.argtypes = [ctypes.c_double] * 3 + [dp]
        L.orc_normalize.argtypes = [dp, dp]
        L.orc_look_at_rows.argtypes = [dp] * 6
        L.orc_to_word8.argtypes = [ctypes.c_double]
        L.orc_to_word8.restype = ctypes.c_uint8
        L.orc_tree_build.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.orc_tree_build.restype = ctypes.c_void_p
        L.orc_tree_free.argtypes = [ctypes.c_void_p]
        L.orc_tree_size.argtypes = [ctypes.c_void_p]
        L.orc_tree_size.restype = ctypes.c_size_t
        L.orc_in_radius.argtypes = [ctypes.c_void_p, ctypes.c_double, dp, ctypes.c_void_p, ctypes.c_size_t]
        L.orc_in_radius.restype = ctypes.c_size_t
        L.orc_star_color.argtypes = [ctypes.c_int, dp, dp]
        L.orc_read_ppm.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.orc_read_ppm.restype = ctypes.c_size_t
        L.orc_star_lookup.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, dp, dp]
        L.orc_srgb.argtypes = [ctypes.c_double]
        L.orc_srgb.restype = ctypes.c_double
        L.orc_blend.argtypes = [dp, dp, dp]
        L.orc_generate_ray.argtypes = [ctypes.POINTER(OrcCamera)] + [ctypes.c_int] * 4 + [dp, dp]
        L.orc_rk4.argtypes = [ctypes.c_double, ctypes.c_double, dp, dp, dp, dp]
        L.orc_trace_ray.argtypes = [ctypes.POINTER(OrcCamera), ctypes.POINTER(OrcScene), ctypes.c_void_p,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
        L.orc_trace_ray.restype = ctypes.c_uint32
        L.orc_render.argtypes = [ctypes.POINTER(OrcCamera), ctypes.POINTER(OrcScene), ctypes.c_void_p,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.POINTER(ctypes.c_uint64)]
        L.orc_render.restype = ctypes.c_int
        L.orc_supersample.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.orc_box_blur.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.orc_bloom.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                ctypes.c_void_p]
        L.orc_to_srgb8.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        _lib = L
    return _lib


def _d3(v):
    return (ctypes.c_double * 3)(*[float(x) for x in v])


def to_orc(cfg) -> Tuple[OrcCamera, OrcScene]:
    """blackstar_b200.config.Config -> oracle structs."""
    cam, scn = cfg.camera, cfg.scene
    c = OrcCamera(_d3(cam.position), _d3(cam.lookAt), _d3(cam.upVec), cam.fov)
    s = OrcScene(scn.stepSize, scn.bloomStrength, scn.starIntensity, scn.starSaturation,
                 _d3(scn.diskColor), scn.diskOpacity, scn.diskInner, scn.diskOuter,
                 scn.bloomDivider, scn.resolution[0], scn.resolution[1], 1 if scn.supersampling else 0)
    return c, s


class Tree:
    """KdMap built by the restated kdt `build` (oracle_thirdparty.c)."""

    def __init__(self, stars: Optional[np.ndarray]):
        self._h = None
        n = 0 if stars is None else len(stars)
        if n:
            arr = np.ascontiguousarray(stars)
            assert arr.dtype.itemsize == 48
            self._h = lib().orc_tree_build(arr.ctypes.data, n)
        self.n = n

    @property
    def handle(self):
        return self._h

    def in_radius(self, radius: float, q) -> np.ndarray:
        idx = np.zeros(4096, dtype=np.uint32)
        n = lib().orc_in_radius(self._h, radius, _d3(q), idx.ctypes.data, idx.size)
        return idx[:n].copy()

    def lookup(self, intensity: float, saturation: float, vel) -> np.ndarray:
        out = (ctypes.c_double * 3)()
        lib().orc_star_lookup(self._h, intensity, saturation, _d3(vel), out)
        return np.array(out[:])

    def __del__(self):
        if self._h and _lib is not None:
            _lib.orc_tree_free(self._h)
            self._h = None


def hsi_to_rgb(h, s, i):
    out = (ctypes.c_double * 3)()
    lib().orc_hsi_to_rgb(h, s, i, out)
    return np.array(out[:])


def rk4(h, h2, vel, pos):
    nv, npos = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
    lib().orc_rk4(h, h2, _d3(vel), _d3(pos), nv, npos)
    return np.array(nv[:]), np.array(npos[:])


def generate_ray(cfg, w, h, x, y):
    c, _ = to_orc(cfg)
    v, p = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
    lib().orc_generate_ray(ctypes.byref(c), w, h, x, y, v, p)
    return np.array(v[:]), np.array(p[:])


def trace_ray(cfg, tree: Optional[Tree], w, h, x, y):
    c, s = to_orc(cfg)
    out = (ctypes.c_double * 3)()
    steps = lib().orc_trace_ray(ctypes.byref(c), ctypes.byref(s), tree.handle if tree else None, w, h, x, y, out)
    return np.array(out[:]), int(steps)


def render(cfg, tree: Optional[Tree] = None, row0: int = 0, row1: Optional[int] = None, nthreads: int = 0):
    """Raytracer.render incl. supersample; returns (rows x W x 3 float64, total rk4 steps)."""
    c, s = to_orc(cfg)
    W, H = cfg.scene.resolution
    if row1 is None:
        row1 = H
    out = np.zeros((row1 - row0, W, 3), dtype=np.float64)
    steps = ctypes.c_uint64(0)
    rc = lib().orc_render(ctypes.byref(c), ctypes.byref(s), tree.handle if tree else None, row0, row1,
                          nthreads, out.ctypes.data, ctypes.byref(steps))
    if rc != 0:
        raise ValueError("orc_render: bad rows")
    return out, int(steps.value)


def supersample(img: np.ndarray) -> np.ndarray:
    h, w, _ = img.shape
    img = np.ascontiguousarray(img, dtype=np.float64)
    out = np.zeros((h // 2, w // 2, 3), dtype=np.float64)
    lib().orc_supersample(img.ctypes.data, h, w, out.ctypes.data)
    return out


def box_blur(r: int, passes: int, img: np.ndarray) -> np.ndarray:
    out = np.array(img, dtype=np.float64, order="C", copy=True)
    h, w, _ = out.shape
    lib().orc_box_blur(r, passes, out.ctypes.data, h, w)
    return out


def bloom(strength: float, divider: int, img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.float64)
    h, w, _ = img.shape
    out = np.zeros_like(img)
    lib().orc_bloom(strength, divider, img.ctypes.data, h, w, out.ctypes.data)
    return out


def to_srgb8(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.float64)
    out = np.zeros(img.shape, dtype=np.uint8)
    lib().orc_to_srgb8(img.ctypes.data, img.size // 3, out.ctypes.data)
    return out


def read_ppm(data: bytes) -> np.ndarray:
    n = max(0, (len(data) - 28) // 28)
    out = np.zeros(n, dtype=ORC_STAR_DTYPE)
    got = lib().orc_read_ppm(data, len(data), out.ctypes.data, n)
    return out[:got]
