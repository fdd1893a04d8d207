"""World-size-2..4 gloo tests of the multi-rank host logic: row tiles, the all-to-all that re-cuts
them into column bands, the shared host frame.  Tiles are produced by the oracle and the two bloom
kernels are replaced by numpy here (no GPU in this container); on the GPU box the same
exchange_transposed() moves the output of bsb_bloom_h_device over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blackstar_b200 import config
from blackstar_b200.dist import (SharedHostFrame, balanced_tiles, col_bands, even_row_tiles, exchange_transposed,
                                 gather_tiles, row_tiles, tiles_from_measurements)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_tiles_partition():
    for H in (1, 7, 113, 1080, 4096):
        for n in (1, 2, 3, 4, 8):
            t = row_tiles(H, n)
            assert t[0][0] == 0 and t[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(t, t[1:]))
            assert max(r1 - r0 for r0, r1 in t) - min(r1 - r0 for r0, r1 in t) <= 1


def test_even_tiles_and_bands_partition_everything():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 9000), st.integers(1, 8))
    def check(size, world):
        for cut in (even_row_tiles(size, world), col_bands(size, world)):
            assert len(cut) == world and cut[0][0] == 0 and cut[-1][1] == size
            assert all(a[1] == b[0] for a, b in zip(cut, cut[1:])) and all(b >= a for a, b in cut)
            assert all(a % 2 == 0 for a, _ in cut)           # interior boundaries are even (pairs of lines)
            if size >= 4 * world:                            # big enough: nobody is empty, nobody has twice the share
                assert min(b - a for a, b in cut) >= 1 and max(b - a for a, b in cut) <= size // world + 3
    check()


def test_balanced_tiles():
    # equal rates, no extra work -> the plain split
    assert balanced_tiles(4096, [10.0] * 8, [0.0] * 8) == row_tiles(4096, 8)
    # rank 0 has 1 ms of rank-only work at 64 rows/ms -> it gets ~56 rows fewer than the others
    t = balanced_tiles(4096, [64.0] * 8, [1.0] + [0.0] * 7)
    sizes = [b - a for a, b in t]
    assert t[0][0] == 0 and t[-1][1] == 4096 and all(a[1] == b[0] for a, b in zip(t, t[1:]))
    assert sizes[0] == 456 and set(sizes[1:]) == {520}
    fin = [sz / 64.0 + (1.0 if k == 0 else 0.0) for k, sz in enumerate(sizes)]
    assert max(fin) - min(fin) < 0.02
    # a slow rank gets fewer rows; degenerate inputs stay valid partitions
    t = balanced_tiles(1000, [1.0, 3.0], [0.0, 0.0])
    assert t == [(0, 250), (250, 1000)]
    t = balanced_tiles(7, [1.0, 1.0, 1.0], [100.0, 0.0, 0.0])
    assert t[0] == (0, 0) and t[-1][1] == 7


def test_tiles_from_measurements_is_robust():
    even = [[512, 8.1, 0.0]] * 8
    assert tiles_from_measurements(4096, even) == row_tiles(4096, 8)
    t = tiles_from_measurements(4096, [[512, 8.1, 1.2]] + [[512, 8.1, 0.0]] * 7)
    assert t[0] == (0, 446) and t[-1][1] == 4096
    # a rank whose tile was too small to time uses the others' rate; a wild rank-only time is capped
    t = tiles_from_measurements(4096, [[0, 0.0, 30.0]] + [[585, 9.3, 0.0]] * 7)
    assert 250 < t[0][1] < 330
    t = tiles_from_measurements(4096, [[512, 8.1, 1e9]] + [[512, 8.3, 0.0]] * 7)
    assert t[0][1] > 250 and all(b > a for a, b in t)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _box3(a, r, axis):
    """Three sweeps of boxBlur's 1-D filter (window [x-r+1, x+r] / (2r+1), zeros outside) along `axis`."""
    a = np.moveaxis(a, axis, 0)
    n = a.shape[0]
    x = np.arange(n)
    hi, lo = np.minimum(x + r, n - 1) + 1, np.maximum(x - r + 1, 0)
    for _ in range(3):
        S = np.concatenate([np.zeros((1,) + a.shape[1:]), np.cumsum(a, axis=0)])
        a = (S[hi] - S[lo]) / (2 * r + 1)
    return np.moveaxis(a, 0, axis)


def _worker(rank, world, port, H, W, divider, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as po
        cfg = config.with_resolution(config.load_config(os.path.join(ROOT, "scenes", "default-aa.yaml")), W, H)
        tiles, bands = even_row_tiles(H, world), col_bands(W, world)
        r0, r1 = tiles[rank]
        c0, c1 = bands[rank]
        r = W // divider
        # my row tile, traced by the oracle (no GPU here); then the pipeline of DistributedFrame.step with
        # numpy standing in for the two kernels
        img, _ = po.render(cfg, None, r0, r1, nthreads=1)                       # (h, W, 3) float64
        midT = torch.from_numpy(np.ascontiguousarray(_box3(img, r, 1).transpose(1, 0, 2)))   # H^3 of my rows, transposed
        imgT = torch.from_numpy(np.ascontiguousarray(img.transpose(1, 0, 2)))
        rmid = torch.zeros((c1 - c0) * H * 3, dtype=torch.float64)
        rimg = torch.zeros((c1 - c0) * H * 3, dtype=torch.float64)
        exchange_transposed([midT, imgT], [rmid, rimg], tiles, bands, rank, world)
        # my band: column l = the pieces of every source rank, in tile order
        w = c1 - c0
        col_mid = np.concatenate([rmid.numpy()[w * t0 * 3:w * t1 * 3].reshape(w, t1 - t0, 3) for t0, t1 in tiles], axis=1)
        col_img = np.concatenate([rimg.numpy()[w * t0 * 3:w * t1 * 3].reshape(w, t1 - t0, 3) for t0, t1 in tiles], axis=1)
        band = (col_img + cfg.scene.bloomStrength * _box3(col_mid, r, 1)).transpose(1, 0, 2)   # (H, w, 3)
        # "N DMA engines fill the host frame": every rank writes its band into the shared frame
        host = SharedHostFrame((H, W, 3), np.float64, rank, world, register=False)
        host.array[:, c0:c1] = band
        dist.barrier()
        if rank == 0:
            np.save(out_path, np.array(host.array))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W,divider", [(2, 22, 24, 5), (3, 21, 26, 4), (4, 10, 9, 3)])
def test_distributed_bloom_pipeline_equals_whole_frame(tmp_path, world, H, W, divider):
    """Row tiles -> horizontal bloom -> all-to-all into column bands -> vertical bloom + combine ->
    shared host frame, on world_size 2..4 over gloo, against the oracle's bloom of the whole frame.
    Uneven tiles, bands narrower than the radius, a rank with an empty band (W=9 over 4 ranks)."""
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), H, W, divider, out), nprocs=world, join=True)
    from oracle import pyoracle as po
    cfg = config.with_resolution(config.load_config(os.path.join(ROOT, "scenes", "default-aa.yaml")), W, H)
    whole, _ = po.render(cfg, None)
    ref = po.bloom(cfg.scene.bloomStrength, divider, whole)
    got = np.load(out)
    assert np.abs(got - ref).max() < 1e-12


def test_gather_tiles_still_works(tmp_path):
    out = str(tmp_path / "full.npy")
    mp.spawn(_gather_worker, args=(2, _free_port(), 21, 24, out), nprocs=2, join=True)
    from oracle import pyoracle as po
    cfg = config.with_resolution(config.load_config(os.path.join(ROOT, "scenes", "default-aa.yaml")), 24, 21)
    whole, _ = po.render(cfg, None)
    np.testing.assert_array_equal(np.load(out), whole.astype(np.float32))


def _gather_worker(rank, world, port, H, W, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as po
        cfg = config.with_resolution(config.load_config(os.path.join(ROOT, "scenes", "default-aa.yaml")), W, H)
        tiles = row_tiles(H, world)
        r0, r1 = tiles[rank]
        img, _ = po.render(cfg, None, r0, r1, nthreads=1)
        mine = torch.from_numpy(img.astype(np.float32))
        full = torch.zeros((H, W, 3), dtype=torch.float32) if rank == 0 else None
        if rank == 0:
            full[r0:r1] = mine
        gather_tiles(full, mine, tiles, rank, world)
        if rank == 0:
            np.save(out_path, full.numpy())
    finally:
        dist.destroy_process_group()
