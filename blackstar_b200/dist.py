"""One-process-per-GPU execution of the path (torchrun): row tiles + ONE gather.

The reference's only parallelism is data-parallel over pixels inside one process
(massiv ``Par``, src/Raytracer.hs:66).  Rays are independent, so the final image is cut
into contiguous row tiles, rank k renders rows [H k/N, H (k+1)/N) with ``bsb_render_device``
and the tiles are gathered on rank 0 with a single grouped NCCL send/recv over NVLink; bloom
runs on rank 0 afterwards because its vertical reach (3 * (W div 25) rows) is about a whole
tile at N = 8.  No other collective exists on this path.

``torch`` is plumbing here (device memory, streams, ``torch.distributed``); the arithmetic
is in libblackstar_b200.so.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .config import Config


def row_tiles(height: int, world: int) -> List[Tuple[int, int]]:
    """Rows [H k/N, H (k+1)/N) for k = 0..N-1 (the same split bsb_render_full uses)."""
    return [(height * k // world, height * (k + 1) // world) for k in range(world)]


def balanced_tiles(height: int, rows_per_ms: Sequence[float], extra_ms: Sequence[float]) -> List[Tuple[int, int]]:
    """Contiguous row tiles such that every rank finishes at the same time.

    Rank k traces ``rows_per_ms[k]`` rows per millisecond (measured on its previous tile) and
    has ``extra_ms[k]`` of work that only it does (rank 0: receiving the gather + bloom).
    Solve r_k / rate_k + extra_k = tau with sum r_k = H.  Every rank evaluates this on the same
    all-gathered numbers, so they agree without further communication.
    """
    n = len(rows_per_ms)
    rate = [max(float(x), 1e-9) for x in rows_per_ms]
    tau = (height + sum(r * e for r, e in zip(rate, extra_ms))) / sum(rate)
    want = [max(0.0, r * (tau - e)) for r, e in zip(rate, extra_ms)]
    scale = height / max(sum(want), 1e-9)
    edges, acc = [0], 0.0
    for k in range(n):
        acc += want[k] * scale
        edges.append(height if k == n - 1 else min(height, max(edges[-1], int(round(acc)))))
    return [(edges[k], edges[k + 1]) for k in range(n)]


def gather_tiles(full: Optional[torch.Tensor], tile: Optional[torch.Tensor], tiles: List[Tuple[int, int]],
                 rank: int, world: int, group=None) -> None:
    """Gather row tiles into ``full`` (H x W x C) on rank 0.  Rank 0's own tile must already
    be in place (it renders straight into ``full``).  One grouped send/recv: with the NCCL
    backend this is the path's single collective; with gloo (CPU tests) the same calls work."""
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


def tiles_from_measurements(height: int, vals: Sequence[Sequence[float]]) -> List[Tuple[int, int]]:
    """vals[k] = (rows traced, trace ms, rank-only ms) as all-gathered in ``calibrate``.
    Ranks with too small a tile to time use the mean rate of the others; rank 0's rank-only
    work is capped at half of an even share so that a hiccup cannot starve it of rows."""
    n = len(vals)
    ok = [v[0] / v[1] for v in vals if v[0] >= 8 and v[1] > 1e-3]
    mean_rate = sum(ok) / len(ok) if ok else 1.0
    rates = [(v[0] / v[1]) if (v[0] >= 8 and v[1] > 1e-3) else mean_rate for v in vals]
    even_ms = height / max(sum(rates), 1e-9)
    extra = [min(max(vals[0][2], 0.0), 0.5 * even_ms)] + [0.0] * (n - 1)
    return balanced_tiles(height, rates, extra)


class TiledFrame:
    """Per-rank state for rendering one scene across the ranks of a process group.

    ``step()`` = Main.doRender's device work for one frame: trace my tile, gather on rank 0,
    bloom on rank 0 (iff bloomStrength /= 0).  The result stays in HBM (``self.full`` on
    rank 0).  ``step_to_host(out)`` additionally copies it into a pinned host tensor.
    """

    def __init__(self, renderer, cfg: Config, rank: int, world: int, device: torch.device):
        self.r, self.cfg, self.rank, self.world, self.device = renderer, cfg, rank, world, device
        W, H = cfg.scene.resolution
        self.W, self.H = W, H
        self.full = torch.empty((H, W, 4), dtype=torch.float32, device=device) if rank == 0 else None
        self.launches = 0
        self._set_tiles(row_tiles(H, world))
        # run the library's kernels on torch's current stream so they order with the NCCL ops
        self.r.set_stream(torch.cuda.current_stream(device).cuda_stream)

    def _set_tiles(self, tiles: List[Tuple[int, int]]):
        self.tiles = tiles
        r0, r1 = tiles[self.rank]
        if self.rank == 0:
            self.tile = self.full[r0:r1]
        else:
            self.tile = torch.empty((max(r1 - r0, 0), self.W, 4), dtype=torch.float32, device=self.device)

    def calibrate(self, iterations: int = 2):
        """Adaptive row tiling: measure every rank's trace rate on its current tile and rank 0's
        rank-only work (gather + bloom), then re-cut the rows so all ranks finish together
        (``balanced_tiles``).  A renderer in production would do this from frame to frame."""
        if self.world == 1:
            return
        scn = self.cfg.scene
        self.step()                      # untimed: NCCL sets its peer connections up on first use
        torch.cuda.synchronize()
        for _ in range(iterations):
            r0, r1 = self.tiles[self.rank]
            st = self.r.render_device(self.cfg, self.tile.data_ptr(), r0, r1, want_stats=True)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gather_tiles(self.full, self.tile, self.tiles, self.rank, self.world)
            if self.rank == 0 and scn.bloomStrength != 0:
                self.r.bloom_device(scn.bloomStrength, scn.bloomDivider, self.W, self.H, self.full.data_ptr(),
                                    self.full.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            mine = torch.tensor([float(r1 - r0), float(st["trace_ms"]), e0.elapsed_time(e1) if self.rank == 0 else 0.0],
                                dtype=torch.float64, device=self.device)
            allv = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(allv, mine)
            vals = [v.tolist() for v in allv]
            self._set_tiles(tiles_from_measurements(self.H, vals))

    def step(self, want_stats: bool = False):
        r0, r1 = self.tiles[self.rank]
        st = self.r.render_device(self.cfg, self.tile.data_ptr(), r0, r1, want_stats=want_stats)
        self.launches += 2 if r1 > r0 else 0   # ray tables + trace
        gather_tiles(self.full, self.tile, self.tiles, self.rank, self.world)
        scn = self.cfg.scene
        if self.rank == 0 and scn.bloomStrength != 0:  # app/Main.hs:113
            self.r.bloom_device(scn.bloomStrength, scn.bloomDivider, self.W, self.H, self.full.data_ptr(),
                                self.full.data_ptr())
            self.launches += 2
        return st

    def step_to_host(self, host_out: Optional[torch.Tensor]):
        """``step()`` + device->host copy of the finished frame into pinned ``host_out`` (rank 0).
        Synchronous with respect to the stream: the next frame starts after the copy."""
        self.step()
        if self.rank == 0:
            host_out.copy_(self.full, non_blocking=True)

    def step_to_host_srgb8(self, host_u8: Optional[torch.Tensor]):
        """``step()`` + writeImg's sRGB / 8-bit map on rank 0's GPU + device->host copy of the RGB8
        image (what the PNG writer consumes): 3 bytes per pixel cross PCIe instead of 16."""
        self.step()
        if self.rank == 0:
            if getattr(self, "_u8", None) is None:
                self._u8 = torch.empty((self.H, self.W, 3), dtype=torch.uint8, device=self.device)
            self.r.to_srgb8_device(self.W, self.H, self.full.data_ptr(), self._u8.data_ptr())
            self.launches += 1
            host_u8.copy_(self._u8, non_blocking=True)

    # ---- pipelined variant: the copy of frame i overlaps the trace of frame i+1 ------------
    def enable_double_buffering(self):
        """Second device frame + a copy stream, so ``step_to_host_pipelined`` can overlap the
        D2H of one frame with the tracing of the next (what a batch / animation driver does)."""
        if self.rank != 0 or getattr(self, "_frames", None) is not None:
            return
        self._frames = [self.full, torch.empty_like(self.full)]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._copied = [None, None]   # event: the copy out of frame buffer b has finished
        self._flip = 0

    def step_to_host_pipelined(self, host_outs: Optional[Sequence[torch.Tensor]]):
        """Like ``step_to_host`` but frame i is rendered into device buffer i % 2 and copied to
        ``host_outs[i % 2]`` on a side stream.  Call ``drain()`` before reading the last frame."""
        if self.rank == 0:
            b = self._flip
            self._flip ^= 1
            if self._copied[b] is not None:
                torch.cuda.current_stream(self.device).wait_event(self._copied[b])  # buffer b is free again
            self.full = self._frames[b]
            r0, r1 = self.tiles[0]
            self.tile = self.full[r0:r1]
        self.step()
        if self.rank == 0:
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(done)
                host_outs[b].copy_(self.full, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._copied[b] = ev

    def drain(self):
        if self.rank == 0 and getattr(self, "_copy_stream", None) is not None:
            self._copy_stream.synchronize()
