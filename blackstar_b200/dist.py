"""One-process-per-GPU execution of the path (torchrun): row tiles, ONE all-to-all, column bands.

The reference's only parallelism is data-parallel over pixels inside one process (massiv ``Par``,
src/Raytracer.hs:66).  Rays are independent, so rank k traces a contiguous tile of rows with
``bsb_render_device``.  The bloom (src/ImageFilters.hs:28-86) is separable:

* its horizontal sweeps need whole rows    -> every rank filters its own row tile, no communication
  (``bsb_bloom_h_device``; the result and the tile itself are written TRANSPOSED, [W][rows]);
* its vertical sweeps need whole columns   -> one all-to-all over NVLink re-cuts the frame from row
  tiles into column bands: what rank j needs of rank i's transposed tile is rows c0_j..c1_j of it, a
  contiguous piece (``exchange_transposed``: one grouped NCCL send/recv, the path's only collective);
* rank j filters its band vertically, adds the original and (optionally) maps it to sRGB8
  (``bsb_bloom_v_device``), and copies the band into the host frame over ITS OWN PCIe link: the host
  frame lives in shared memory that every rank has page-locked, so N DMA engines fill it in parallel.

Without bloom the row tiles go straight to the host frame.  ``torch`` is plumbing here (device memory,
streams, ``torch.distributed``); the arithmetic is in libblackstar_b200.so.
"""
from __future__ import annotations

import ctypes
import mmap
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .config import Config


def row_tiles(height: int, world: int) -> List[Tuple[int, int]]:
    """Rows [H k/N, H (k+1)/N) for k = 0..N-1."""
    return [(height * k // world, height * (k + 1) // world) for k in range(world)]


def even_row_tiles(height: int, world: int) -> List[Tuple[int, int]]:
    """``row_tiles`` with even interior boundaries (the bloom kernel writes pairs of adjacent rows)."""
    edges = [0] + [(height * k // world) & ~1 for k in range(1, world)] + [height]
    return [(edges[k], max(edges[k], edges[k + 1])) for k in range(world)]


def col_bands(width: int, world: int) -> List[Tuple[int, int]]:
    """Column bands with even boundaries (the bloom kernel pairs adjacent columns); the same cut
    bsb_render_full makes inside the library."""
    edges = [0]
    for k in range(world):
        e = width if k == world - 1 else min(width, (width * (k + 1) // world) & ~1)
        edges.append(max(edges[-1], e))
    return [(edges[k], edges[k + 1]) for k in range(world)]


def balanced_tiles(height: int, rows_per_ms: Sequence[float], extra_ms: Optional[Sequence[float]] = None,
                   even: bool = False) -> List[Tuple[int, int]]:
    """Contiguous row tiles such that every rank finishes at the same time.

    Rank k traces ``rows_per_ms[k]`` rows per millisecond (measured on its previous tile) and has
    ``extra_ms[k]`` of work that only it does.  Solve r_k / rate_k + extra_k = tau with sum r_k = H.
    Every rank evaluates this on the same all-gathered numbers, so they agree without further
    communication.
    """
    n = len(rows_per_ms)
    extra_ms = list(extra_ms) if extra_ms is not None else [0.0] * n
    rate = [max(float(x), 1e-9) for x in rows_per_ms]
    tau = (height + sum(r * e for r, e in zip(rate, extra_ms))) / sum(rate)
    want = [max(0.0, r * (tau - e)) for r, e in zip(rate, extra_ms)]
    scale = height / max(sum(want), 1e-9)
    edges, acc = [0], 0.0
    for k in range(n):
        acc += want[k] * scale
        e = int(round(acc))
        if even:
            e &= ~1
        edges.append(height if k == n - 1 else min(height, max(edges[-1], e)))
    return [(edges[k], edges[k + 1]) for k in range(n)]


def tiles_from_measurements(height: int, vals: Sequence[Sequence[float]], even: bool = False) -> List[Tuple[int, int]]:
    """vals[k] = (rows traced, trace ms[, rank-only ms]) as all-gathered in ``calibrate``.
    Ranks with too small a tile to time use the mean rate of the others; rank-only work is capped at
    half of an even share so that a hiccup cannot starve a rank of rows."""
    n = len(vals)
    ok = [v[0] / v[1] for v in vals if v[0] >= 8 and v[1] > 1e-3]
    mean_rate = sum(ok) / len(ok) if ok else 1.0
    rates = [(v[0] / v[1]) if (v[0] >= 8 and v[1] > 1e-3) else mean_rate for v in vals]
    even_ms = height / max(sum(rates), 1e-9)
    extra = [min(max(v[2] if len(v) > 2 else 0.0, 0.0), 0.5 * even_ms) for v in vals]
    return balanced_tiles(height, rates, extra, even=even)


def gather_tiles(full: Optional[torch.Tensor], tile: Optional[torch.Tensor], tiles: List[Tuple[int, int]],
                 rank: int, world: int, group=None) -> None:
    """Gather row tiles into ``full`` (H x W x C) on rank 0 (rank 0's own tile must already be in
    place).  Not on the default path any more (the frame is assembled in host memory); kept for callers
    that want the whole pre-bloom frame on one GPU."""
    if world == 1:
        return
    ops = []
    if rank == 0:
        for k in range(1, world):
            r0, r1 = tiles[k]
            if r1 > r0:
                ops.append(dist.P2POp(dist.irecv, full[r0:r1], k, group))
    else:
        r0, r1 = tiles[rank]
        if r1 > r0:
            ops.append(dist.P2POp(dist.isend, tile, 0, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def exchange_transposed(sendT: Sequence[torch.Tensor], recv: Sequence[torch.Tensor], tiles: List[Tuple[int, int]],
                        bands: List[Tuple[int, int]], rank: int, world: int, group=None) -> None:
    """The all-to-all that re-cuts the frame from row tiles into column bands.

    ``sendT``: this rank's transposed tiles, each of shape (W, h_rank, C).  ``recv``: this rank's
    bands, each a flat tensor holding, for every source rank i in order, a (w_rank, h_i, C) piece
    (offset w_rank * r0_i * C) -- column l of the band is then piece i's row l, for rows r0_i..r1_i.
    One grouped send/recv; the own piece is a local copy."""
    h_me = tiles[rank][1] - tiles[rank][0]
    c0_me, c1_me = bands[rank]
    w_me = c1_me - c0_me
    ops = []
    for s, r in zip(sendT, recv):
        C = s.shape[-1] if s.dim() == 3 else 1
        flat_r = r.view(-1)
        for j in range(world):              # what I send to j: rows c0_j..c1_j of my transposed tile
            c0, c1 = bands[j]
            if c1 <= c0 or h_me <= 0:
                continue
            piece = s[c0:c1]
            if j == rank:
                off = w_me * tiles[rank][0] * C
                flat_r[off:off + piece.numel()].copy_(piece.reshape(-1))
            else:
                ops.append(dist.P2POp(dist.isend, piece, j, group))
        for i in range(world):              # what I receive from i: my band's rows r0_i..r1_i
            h_i = tiles[i][1] - tiles[i][0]
            if i == rank or h_i <= 0 or w_me <= 0:
                continue
            off = w_me * tiles[i][0] * C
            ops.append(dist.P2POp(dist.irecv, flat_r[off:off + w_me * h_i * C].view(w_me, h_i, C), i, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class SharedHostFrame:
    """A host frame in POSIX shared memory that every rank of the node maps and page-locks, so that
    each rank's GPU can DMA its part of the finished image straight into the buffer rank 0 reads."""

    def __init__(self, shape: Tuple[int, ...], dtype, rank: int, world: int, register: bool = True):
        self.shape, self.rank = tuple(shape), rank
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        name = [None]
        if rank == 0:
            name[0] = f"/dev/shm/bsb_frame_{os.getpid()}_{id(self):x}_{np.dtype(dtype).name}"
            with open(name[0], "wb") as f:
                f.truncate(max(nbytes, 1))
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        self.path = name[0]
        self._f = open(self.path, "r+b")
        self._mm = mmap.mmap(self._f.fileno(), max(nbytes, 1))
        self.array = np.frombuffer(self._mm, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self.ptr = self.array.ctypes.data
        self.nbytes = nbytes
        self._registered = False
        if register and torch.cuda.is_available() and nbytes > 0:
            rc = torch.cuda.cudart().cudaHostRegister(self.ptr, nbytes, 0)
            self._registered = int(rc) == 0
        if world > 1:
            dist.barrier()
        if rank == 0:
            os.unlink(self.path)     # the mappings keep it alive; nothing is left behind in /dev/shm

    def close(self):
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.ptr)
            self._registered = False


class DistributedFrame:
    """Per-rank state for rendering one scene across the ranks of a process group.

    ``step()`` = Main.doRender's device work for one frame (app/Main.hs:105-118): trace my row tile,
    horizontal bloom on it, all-to-all, vertical bloom + combine on my column band (iff
    bloomStrength /= 0).  The result stays in HBM: ``self.band`` (H x w_rank x 4 float) and / or
    ``self.band8`` (H x w_rank x 3 uint8) -- or, without bloom, the row tile.  ``step_to_host``
    additionally copies my part into the shared host frame.
    """

    def __init__(self, renderer, cfg: Config, rank: int, world: int, device: torch.device):
        self.r, self.cfg, self.rank, self.world, self.device = renderer, cfg, rank, world, device
        W, H = cfg.scene.resolution
        self.W, self.H = W, H
        scn = cfg.scene
        self.bloom = scn.bloomStrength != 0
        self.radius = W // scn.bloomDivider if self.bloom else 0
        self.launches = 0
        self.bands = col_bands(W, world)
        self._bufs = {}
        self._set_tiles(even_row_tiles(H, world))
        # run the library's kernels on torch's current stream so they order with the NCCL ops
        self.r.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.host32: Optional[SharedHostFrame] = None
        self.host8: Optional[SharedHostFrame] = None

    # ------------------------------------------------------------------ buffers
    def _buf(self, key, shape, dtype=torch.float32):
        t = self._bufs.get(key)
        n = int(np.prod(shape))
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t[:n].view(*shape) if n else t[:0]

    def _set_tiles(self, tiles: List[Tuple[int, int]]):
        self.tiles = tiles
        r0, r1 = tiles[self.rank]
        self.tile = self._buf("tile", (max(r1 - r0, 0), self.W, 4))

    # ------------------------------------------------------------------ calibration
    def calibrate(self, iterations: int = 2):
        """Adaptive row tiling: measure every rank's trace rate on its current tile, then re-cut the
        rows so all ranks finish together.  A renderer in production does this from frame to frame."""
        if self.world == 1:
            return
        self.step()                      # untimed: NCCL sets its peer connections up on first use
        torch.cuda.synchronize()
        for _ in range(iterations):
            r0, r1 = self.tiles[self.rank]
            st = self.r.render_device(self.cfg, self.tile.data_ptr(), r0, r1, want_stats=True)
            mine = torch.tensor([float(r1 - r0), float(st["trace_ms"])], dtype=torch.float64, device=self.device)
            allv = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(allv, mine)
            self._set_tiles(tiles_from_measurements(self.H, [v.tolist() for v in allv], even=True))

    # ------------------------------------------------------------------ one frame
    def step(self, want_stats: bool = False, want_float: bool = True, want_rgb8: bool = False):
        scn = self.cfg.scene
        r0, r1 = self.tiles[self.rank]
        h = r1 - r0
        st = self.r.render_device(self.cfg, self.tile.data_ptr(), r0, r1, want_stats=want_stats)
        self.launches += 2 if h > 0 else 0   # ray tables + trace
        if not self.bloom:
            if want_rgb8 and h > 0:
                self.tile8 = self._buf("tile8", (h, self.W, 3), torch.uint8)
                self.r.to_srgb8_device(self.W, h, self.tile.data_ptr(), self.tile8.data_ptr())
                self.launches += 1
            return st
        c0, c1 = self.bands[self.rank]
        w = c1 - c0
        if self.world == 1:
            # nothing to exchange: the two-launch bloom of the library on the whole frame
            self.band = self.tile if want_float else None
            self.band8 = self._buf("band8", (self.H, self.W, 3), torch.uint8) if want_rgb8 else None
            self.r.bloom_device(scn.bloomStrength, scn.bloomDivider, self.W, self.H, self.tile.data_ptr(),
                                self.tile.data_ptr() if want_float else 0, self.band8.data_ptr() if want_rgb8 else 0)
            self.launches += 2
            return st
        midT = self._buf("midT", (self.W, max(h, 0), 4))
        imgT = self._buf("imgT", (self.W, max(h, 0), 4))
        if h > 0:
            self.r.bloom_h_device(self.radius, self.W, h, self.tile.data_ptr(), midT.data_ptr(), imgT.data_ptr())
            self.launches += 2
        rmid = self._buf("rmid", (w * self.H * 4,))
        rimg = self._buf("rimg", (w * self.H * 4,))
        exchange_transposed([midT, imgT], [rmid, rimg], self.tiles, self.bands, self.rank, self.world)
        if w > 0:
            self.band = self._buf("band", (self.H, w, 4)) if want_float else None
            self.band8 = self._buf("band8", (self.H, w, 3), torch.uint8) if want_rgb8 else None
            offs = [w * t0 * 16 for t0, _ in self.tiles]   # bytes
            self.r.bloom_v_device(scn.bloomStrength, self.radius, self.H, w,
                                  [rmid.data_ptr() + o for o in offs], [rimg.data_ptr() + o for o in offs],
                                  [t1 - t0 for t0, t1 in self.tiles],
                                  self.band.data_ptr() if want_float else 0, self.band8.data_ptr() if want_rgb8 else 0)
            self.launches += 1
        return st

    # ------------------------------------------------------------------ host output
    def _host(self, rgb8: bool) -> SharedHostFrame:
        if rgb8:
            if self.host8 is None:
                self.host8 = SharedHostFrame((self.H, self.W, 3), np.uint8, self.rank, self.world)
            return self.host8
        if self.host32 is None:
            self.host32 = SharedHostFrame((self.H, self.W, 4), np.float32, self.rank, self.world)
        return self.host32

    def step_to_host(self, rgb8: bool = False) -> np.ndarray:
        """``step()`` + device->host copy of MY part of the finished frame into the shared host frame
        (every rank over its own PCIe link), then a barrier: when it returns, the whole frame is in
        host memory (returned as a numpy view; rank 0 is the one that would hand it to the PNG writer)."""
        host = self._host(rgb8)
        px = 3 if rgb8 else 16
        self.step(want_float=not rgb8, want_rgb8=rgb8)
        if self.bloom and self.world > 1:
            c0, c1 = self.bands[self.rank]
            src = self.band8 if rgb8 else self.band
            if c1 > c0:
                self.r.download_2d(host.ptr + c0 * px, self.W * px, src.data_ptr(), (c1 - c0) * px, (c1 - c0) * px, self.H)
        elif self.bloom:
            src = self.band8 if rgb8 else self.band
            self.r.download_2d(host.ptr, self.W * px, src.data_ptr(), self.W * px, self.W * px, self.H)
        else:
            r0, r1 = self.tiles[self.rank]
            src = self.tile8 if rgb8 else self.tile
            if r1 > r0:
                self.r.download_2d(host.ptr + r0 * self.W * px, self.W * px, src.data_ptr(), self.W * px, self.W * px, r1 - r0)
        self.r.synchronize()          # also surfaces BSB_ERR_STEPCAP of this frame's (asynchronous) trace launch
        if self.world > 1:
            dist.barrier()
        return host.array

    def close(self):
        for h in (self.host32, self.host8):
            if h is not None:
                h.close()
