"""ctypes binding of libblackstar_b200.so (the C ABI in include/blackstar_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C blackstar_b200/csrc``.
If it is missing this module raises: there is no CPU fallback and no other backend.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
#: the in-tree library; BLACKSTAR_B200_LIB overrides it (used by tools/ to time experimental builds)
LIB_PATH = os.environ.get("BLACKSTAR_B200_LIB") or os.path.join(_HERE, "libblackstar_b200.so")

#: every symbol include/blackstar_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "bsb_create", "bsb_create_on", "bsb_destroy", "bsb_last_error", "bsb_version", "bsb_set_stream",
    "bsb_set_option", "bsb_set_stars", "bsb_set_stars_ppm", "bsb_star_count", "bsb_render",
    "bsb_render_device", "bsb_bloom", "bsb_bloom_device", "bsb_render_full", "bsb_to_srgb8",
    "bsb_to_srgb8_device", "bsb_render_full_srgb8", "bsb_measure_fp64_peak", "bsb_measure_hbm_copy",
    "bsb_selftest_rinv5", "bsb_render_full_both", "bsb_render_full_device", "bsb_synchronize",
    "bsb_bloom_h_device", "bsb_bloom_v_device", "bsb_bloom_to_device", "bsb_download_2d", "bsb_set_stars_file",
]


class BlackstarError(RuntimeError):
    """A non-zero status from the C ABI (message from bsb_last_error)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[bsb status {code}] {message}")
        self.code = code
        self.message = message


class CStats(ctypes.Structure):
    """bsb_stats"""
    _fields_ = [("rays", ctypes.c_uint64), ("steps", ctypes.c_uint64), ("capped", ctypes.c_uint64),
                ("star_hits", ctypes.c_uint64), ("trace_ms", ctypes.c_double), ("bloom_ms", ctypes.c_double),
                ("gather_ms", ctypes.c_double), ("d2h_ms", ctypes.c_double), ("total_ms", ctypes.c_double),
                ("n_gpus", ctypes.c_int32), ("launches", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load() -> ctypes.CDLL:
    """dlopen the library and declare the prototypes.  Raises if the .so is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  blackstar_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, cp, d, i, sz = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double, ctypes.c_int, ctypes.c_size_t
    dp = ctypes.POINTER(ctypes.c_double)
    L.bsb_create.argtypes = [i]; L.bsb_create.restype = vp
    L.bsb_create_on.argtypes = [ctypes.POINTER(ctypes.c_int), i]; L.bsb_create_on.restype = vp
    L.bsb_destroy.argtypes = [vp]; L.bsb_destroy.restype = None
    L.bsb_last_error.argtypes = [vp]; L.bsb_last_error.restype = cp
    L.bsb_version.argtypes = []; L.bsb_version.restype = cp
    L.bsb_set_stream.argtypes = [vp, vp]; L.bsb_set_stream.restype = i
    L.bsb_set_option.argtypes = [vp, cp, d]; L.bsb_set_option.restype = i
    L.bsb_set_stars.argtypes = [vp, vp, sz]; L.bsb_set_stars.restype = i
    L.bsb_set_stars_ppm.argtypes = [vp, cp, sz]; L.bsb_set_stars_ppm.restype = i
    L.bsb_set_stars_file.argtypes = [vp, cp, sz]; L.bsb_set_stars_file.restype = i
    L.bsb_star_count.argtypes = [vp]; L.bsb_star_count.restype = sz
    L.bsb_render.argtypes = [vp, vp, vp, i, i, vp, vp]; L.bsb_render.restype = i
    L.bsb_render_device.argtypes = [vp, vp, vp, i, i, vp, vp]; L.bsb_render_device.restype = i
    L.bsb_bloom.argtypes = [vp, d, i, i, i, vp, vp]; L.bsb_bloom.restype = i
    L.bsb_bloom_device.argtypes = [vp, d, i, i, i, vp, vp]; L.bsb_bloom_device.restype = i
    L.bsb_render_full.argtypes = [vp, vp, vp, vp, vp]; L.bsb_render_full.restype = i
    L.bsb_to_srgb8.argtypes = [vp, i, i, vp, vp]; L.bsb_to_srgb8.restype = i
    L.bsb_to_srgb8_device.argtypes = [vp, i, i, vp, vp]; L.bsb_to_srgb8_device.restype = i
    L.bsb_render_full_srgb8.argtypes = [vp, vp, vp, vp, vp]; L.bsb_render_full_srgb8.restype = i
    L.bsb_render_full_both.argtypes = [vp, vp, vp, vp, vp, vp]; L.bsb_render_full_both.restype = i
    L.bsb_render_full_device.argtypes = [vp, vp, vp, i, i]; L.bsb_render_full_device.restype = i
    L.bsb_synchronize.argtypes = [vp]; L.bsb_synchronize.restype = i
    L.bsb_bloom_h_device.argtypes = [vp, i, i, i, vp, vp, vp]; L.bsb_bloom_h_device.restype = i
    L.bsb_bloom_v_device.argtypes = [vp, d, i, i, i, i, vp, vp, vp, vp, vp]; L.bsb_bloom_v_device.restype = i
    L.bsb_bloom_to_device.argtypes = [vp, d, i, i, i, vp, vp, vp]; L.bsb_bloom_to_device.restype = i
    L.bsb_download_2d.argtypes = [vp, vp, sz, vp, sz, sz, i]; L.bsb_download_2d.restype = i
    L.bsb_measure_fp64_peak.argtypes = [vp, dp]; L.bsb_measure_fp64_peak.restype = i
    L.bsb_measure_hbm_copy.argtypes = [vp, sz, i, dp]; L.bsb_measure_hbm_copy.restype = i
    L.bsb_selftest_rinv5.argtypes = [vp, d, d, i, dp, dp]; L.bsb_selftest_rinv5.restype = i
    _lib = L
    return L
