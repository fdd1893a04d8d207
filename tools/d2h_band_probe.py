#!/usr/bin/env python
"""How fast is the 2-D device->host copy of a column band (what every rank of the multi-GPU pipeline does with
its part of the frame) compared with a contiguous copy of the same bytes?  Page-locked host frame."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from blackstar_b200.render import Renderer

H, W = 4096, 4096
r = Renderer(devices=[0])
r.set_stream(torch.cuda.current_stream().cuda_stream)
for px, name in ((3, "RGB8"), (16, "float4")):
    host = torch.empty((H, W * px), dtype=torch.uint8, pin_memory=True)
    for nb in (8, 4, 2, 1):
        wb = W // nb
        dev = torch.zeros((H, wb * px), dtype=torch.uint8, device="cuda")
        def f2d():
            r.download_2d(host.data_ptr(), W * px, dev.data_ptr(), wb * px, wb * px, H)
        def f1d():
            host.view(-1)[: H * wb * px].copy_(dev.view(-1), non_blocking=True)
        res = []
        for f in (f2d, f1d):
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(20):
                f()
            torch.cuda.synchronize()
            res.append((time.perf_counter() - t) / 20 * 1e3)
        mb = H * wb * px / 1e6
        print(f"{name} band 1/{nb} of the frame ({wb * px} B rows, {mb:.1f} MB): 2-D copy {res[0]:.3f} ms = {mb / res[0]:.1f} GB/s; "
              f"contiguous {res[1]:.3f} ms = {mb / res[1]:.1f} GB/s")
r.close()
