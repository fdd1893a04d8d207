"""Host-side mirror of the reference's render pipeline, forwarding to the C ABI.

Same names and argument meaning as the Haskell exports this path replaces:

* ``render``      -- Raytracer.render :: Config -> StarTree -> Image   (src/Raytracer.hs:53-67)
* ``bloom``       -- ImageFilters.bloom :: Double -> Int -> Image -> IO Image (src/ImageFilters.hs:80-86)
* ``write_img``   -- Raytracer.writeImg (src/Raytracer.hs:29-32)
* ``do_render``   -- Main.doRender's body (app/Main.hs:105-118): render, then bloom iff
  ``bloomStrength /= 0``

A ``Renderer`` owns one ``bsb_ctx`` (one or more GPUs of this process).  Every image is a
``numpy`` array of shape (H, W, 4) float32: linear RGB + alpha = 1.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .config import Config, to_c
from .starmap import STAR_DTYPE


class Renderer:
    """One bsb_ctx.  ``devices``: explicit CUDA device indices (default: device 0)."""

    def __init__(self, devices: Optional[Sequence[int]] = None, n_gpus: Optional[int] = None):
        self._L = _lib.load()
        if devices is None and n_gpus is None:
            devices = [0]
        if devices is not None:
            arr = (ctypes.c_int * len(devices))(*devices)
            self._ctx = self._L.bsb_create_on(arr, len(devices))
        else:
            self._ctx = self._L.bsb_create(int(n_gpus))
        if not self._ctx:
            raise _lib.BlackstarError(-1, self._L.bsb_last_error(None).decode())
        self.last_stats: Optional[dict] = None

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_ctx", None):
            self._L.bsb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise _lib.BlackstarError(rc, self._L.bsb_last_error(self._ctx).decode())

    def set_option(self, key: str, value: float):
        self._check(self._L.bsb_set_option(self._ctx, key.encode(), float(value)))

    def set_stream(self, cuda_stream: Optional[int]):
        """Run device 0's work on an existing cudaStream_t (e.g. torch's current stream).
        ``None`` restores the ctx's own stream; handle 0 (torch's default stream) is passed as
        cudaStreamLegacy, because NULL means "restore" in the C ABI."""
        if cuda_stream is None:
            handle = 0
        else:
            handle = int(cuda_stream) or 1  # (cudaStream_t)0x1 == cudaStreamLegacy
        self._check(self._L.bsb_set_stream(self._ctx, ctypes.c_void_p(handle)))

    # ------------------------------------------------------------------ star map
    def set_stars(self, stars: Optional[np.ndarray]):
        """Replaces the StarTree argument of render (readTreeFromFile, src/StarMap.hs:82-85)."""
        if stars is None or len(stars) == 0:
            self._check(self._L.bsb_set_stars(self._ctx, None, 0))
            return
        arr = np.ascontiguousarray(stars)
        if arr.dtype.itemsize != STAR_DTYPE.itemsize:
            raise ValueError("stars must have dtype blackstar_b200.starmap.STAR_DTYPE")
        self._check(self._L.bsb_set_stars(self._ctx, arr.ctypes.data, len(arr)))

    def set_stars_ppm(self, data: bytes):
        self._check(self._L.bsb_set_stars_ppm(self._ctx, data, len(data)))

    def set_stars_file(self, data: bytes):
        """The file ``--starmap`` names: the reference's stars.kdt tree file or a PPM catalogue."""
        self._check(self._L.bsb_set_stars_file(self._ctx, data, len(data)))

    @property
    def star_count(self) -> int:
        return int(self._L.bsb_star_count(self._ctx))

    # ------------------------------------------------------------------ render
    def render(self, cfg: Config, row0: int = 0, row1: Optional[int] = None,
               out: Optional[np.ndarray] = None) -> np.ndarray:
        """Raytracer.render (incl. supersample) for rows [row0,row1) of the final image."""
        W, H = cfg.scene.resolution
        if row1 is None:
            row1 = H
        rows = max(0, row1 - row0)
        if out is None:
            out = np.empty((rows, W, 4), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size >= rows * W * 4
        cam, scn = to_c(cfg)
        st = _lib.CStats()
        rc = self._L.bsb_render(self._ctx, ctypes.byref(cam), ctypes.byref(scn), row0, row1,
                                out.ctypes.data, ctypes.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return out

    def render_device(self, cfg: Config, dev_ptr: int, row0: int = 0, row1: Optional[int] = None,
                      want_stats: bool = False) -> Optional[dict]:
        """Same, into device memory (a raw pointer, e.g. ``tensor.data_ptr()``); asynchronous
        unless ``want_stats``."""
        W, H = cfg.scene.resolution
        if row1 is None:
            row1 = H
        cam, scn = to_c(cfg)
        st = _lib.CStats()
        rc = self._L.bsb_render_device(self._ctx, ctypes.byref(cam), ctypes.byref(scn), row0, row1,
                                       ctypes.c_void_p(dev_ptr), ctypes.byref(st) if want_stats else None)
        if want_stats:
            self.last_stats = st.as_dict()
        self._check(rc)
        return self.last_stats if want_stats else None

    # ------------------------------------------------------------------ bloom
    def bloom(self, strength: float, divider: int, img: np.ndarray) -> np.ndarray:
        """ImageFilters.bloom strength divider img."""
        img = np.ascontiguousarray(img, dtype=np.float32)
        H, W, C = img.shape
        assert C == 4
        out = np.empty_like(img)
        self._check(self._L.bsb_bloom(self._ctx, float(strength), int(divider), W, H, img.ctypes.data, out.ctypes.data))
        return out

    def bloom_device(self, strength: float, divider: int, W: int, H: int, src_ptr: int, dst_ptr: int, rgb8_ptr: int = 0):
        """ImageFilters.bloom on device memory; with ``rgb8_ptr`` writeImg's sRGB8 map of the result is
        written too (fused into the second launch), and ``dst_ptr`` may then be 0."""
        self._check(self._L.bsb_bloom_to_device(self._ctx, float(strength), int(divider), W, H, ctypes.c_void_p(src_ptr),
                                                ctypes.c_void_p(dst_ptr or None), ctypes.c_void_p(rgb8_ptr or None)))

    def download_2d(self, host_ptr: int, host_pitch: int, dev_ptr: int, dev_pitch: int, width_bytes: int, rows: int):
        self._check(self._L.bsb_download_2d(self._ctx, ctypes.c_void_p(host_ptr), host_pitch, ctypes.c_void_p(dev_ptr),
                                            dev_pitch, width_bytes, rows))

    # ------------------------------------------------------------------ doRender
    def do_render(self, cfg: Config, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Main.doRender between scene load and writeImg: render on every GPU of the ctx,
        gather, bloom iff bloomStrength /= 0 (app/Main.hs:109-118)."""
        W, H = cfg.scene.resolution
        if out is None:
            out = np.empty((H, W, 4), dtype=np.float32)
        cam, scn = to_c(cfg)
        st = _lib.CStats()
        rc = self._L.bsb_render_full(self._ctx, ctypes.byref(cam), ctypes.byref(scn), out.ctypes.data, ctypes.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return out

    def do_render_srgb8(self, cfg: Config, out: Optional[np.ndarray] = None) -> np.ndarray:
        """do_render followed by writeImg's sRGB + toWord8 map on the device -> (H, W, 3) uint8."""
        W, H = cfg.scene.resolution
        if out is None:
            out = np.empty((H, W, 3), dtype=np.uint8)
        cam, scn = to_c(cfg)
        st = _lib.CStats()
        rc = self._L.bsb_render_full_srgb8(self._ctx, ctypes.byref(cam), ctypes.byref(scn), out.ctypes.data, ctypes.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return out

    def render_full_device(self, cfg: Config, want_float: bool = True, want_rgb8: bool = False):
        """doRender with the frame left on the GPU(s) of the ctx; asynchronous (``synchronize``)."""
        cam, scn = to_c(cfg)
        self._check(self._L.bsb_render_full_device(self._ctx, ctypes.byref(cam), ctypes.byref(scn), int(want_float), int(want_rgb8)))

    def synchronize(self):
        self._check(self._L.bsb_synchronize(self._ctx))

    # ------------------------------------------------------------------ distributed bloom (one process per GPU)
    def bloom_h_device(self, radius: int, W: int, rows: int, src_ptr: int, midT_ptr: int, imgT_ptr: int = 0):
        """Horizontal half of the bloom on a tile of rows: [rows][W] -> transposed [W][rows]."""
        self._check(self._L.bsb_bloom_h_device(self._ctx, int(radius), int(W), int(rows), ctypes.c_void_p(src_ptr),
                                               ctypes.c_void_p(midT_ptr), ctypes.c_void_p(imgT_ptr or None)))

    def bloom_v_device(self, strength: float, radius: int, H: int, cols: int, seg_mid: Sequence[int], seg_img: Sequence[int],
                       seg_rows: Sequence[int], out_ptr: int = 0, rgb8_ptr: int = 0):
        """Vertical half + combine (+ sRGB8) on a band of ``cols`` columns assembled from row-tile pieces."""
        n = len(seg_rows)
        mids = (ctypes.c_void_p * n)(*[ctypes.c_void_p(p) for p in seg_mid])
        imgs = (ctypes.c_void_p * n)(*[ctypes.c_void_p(p) for p in seg_img])
        rows = (ctypes.c_int * n)(*[int(x) for x in seg_rows])
        self._check(self._L.bsb_bloom_v_device(self._ctx, float(strength), int(radius), int(H), int(cols), n, mids, imgs, rows,
                                               ctypes.c_void_p(out_ptr or None), ctypes.c_void_p(rgb8_ptr or None)))

    def to_srgb8(self, img: np.ndarray) -> np.ndarray:
        """The per-pixel map of writeImg: toWord8 . fmap sRGB  (src/Raytracer.hs:23-32)."""
        img = np.ascontiguousarray(img, dtype=np.float32)
        H, W, C = img.shape
        assert C == 4
        out = np.empty((H, W, 3), dtype=np.uint8)
        self._check(self._L.bsb_to_srgb8(self._ctx, W, H, img.ctypes.data, out.ctypes.data))
        return out

    def to_srgb8_device(self, W: int, H: int, src_ptr: int, dst_ptr: int):
        """writeImg's pixel map on device memory: float4 frame -> packed RGB8 (asynchronous)."""
        self._check(self._L.bsb_to_srgb8_device(self._ctx, W, H, ctypes.c_void_p(src_ptr), ctypes.c_void_p(dst_ptr)))

    # ------------------------------------------------------------------ roofline helpers
    def measure_fp64_peak(self) -> float:
        v = ctypes.c_double()
        self._check(self._L.bsb_measure_fp64_peak(self._ctx, ctypes.byref(v)))
        return v.value

    def measure_hbm_copy(self, nbytes: int = 1 << 30, reps: int = 5) -> float:
        v = ctypes.c_double()
        self._check(self._L.bsb_measure_hbm_copy(self._ctx, nbytes, reps, ctypes.byref(v)))
        return v.value

    def selftest_rinv5(self, q_lo: float = 0.25, q_hi: float = 1e5, n: int = 1 << 20) -> Tuple[float, float]:
        a, b = ctypes.c_double(), ctypes.c_double()
        self._check(self._L.bsb_selftest_rinv5(self._ctx, q_lo, q_hi, n, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value


def write_img(img: np.ndarray, path: str, renderer: Optional[Renderer] = None):
    """Raytracer.writeImg (src/Raytracer.hs:29-32): sRGB, 8 bit, PNG RGB8."""
    from PIL import Image
    if img.dtype != np.uint8:
        if renderer is None:
            raise ValueError("write_img needs a Renderer for the sRGB map of a float image")
        img = renderer.to_srgb8(img)
    Image.fromarray(np.ascontiguousarray(img)).save(path, format="PNG")
