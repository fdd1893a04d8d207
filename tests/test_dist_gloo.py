"""World-size-2 (and 3) gloo tests of the multi-rank host logic: row tiling + the single
gather.  Tiles are produced by the oracle here (no GPU in this container); on the GPU box the
same gather_tiles() moves tiles rendered by bsb_render_device over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blackstar_b200 import config
from blackstar_b200.dist import balanced_tiles, gather_tiles, row_tiles, tiles_from_measurements

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_tiles_partition():
    for H in (1, 7, 113, 1080, 4096):
        for n in (1, 2, 3, 4, 8):
            t = row_tiles(H, n)
            assert t[0][0] == 0 and t[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(t, t[1:]))
            assert max(r1 - r0 for r0, r1 in t) - min(r1 - r0 for r0, r1 in t) <= 1


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


def _worker(rank, world, port, H, W, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as po
        cfg = config.with_resolution(config.load_config(os.path.join(ROOT, "scenes", "default-aa.yaml")), W, H)
        tiles = row_tiles(H, world)
        r0, r1 = tiles[rank]
        img, _ = po.render(cfg, None, r0, r1, nthreads=1)
        mine = torch.from_numpy(np.concatenate([img, np.ones(img.shape[:2] + (1,))], axis=2).astype(np.float32))
        full = torch.zeros((H, W, 4), dtype=torch.float32) if rank == 0 else None
        if rank == 0:
            full[r0:r1] = mine
        gather_tiles(full, mine, tiles, rank, world)
        if rank == 0:
            np.save(out_path, full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_of_row_tiles_equals_whole_frame(tmp_path, world):
    H, W = 21, 24   # 21 rows over 2 or 3 ranks: uneven tiles
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), H, W, out), nprocs=world, join=True)
    from oracle import pyoracle as po
    cfg = config.with_resolution(config.load_config(os.path.join(ROOT, "scenes", "default-aa.yaml")), W, H)
    whole, _ = po.render(cfg, None)
    got = np.load(out)
    np.testing.assert_array_equal(got[..., :3], whole.astype(np.float32))
    assert (got[..., 3] == 1).all()
