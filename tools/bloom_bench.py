#!/usr/bin/env python
"""Times the bloom launches (and the sRGB8 map) alone on cuda:0, frame resident in HBM.
usage: python tools/bloom_bench.py [W H]...   (default 4096 4096 and 8192 8192)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blackstar_b200.render import Renderer  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    sizes = [(4096, 4096), (8192, 8192)]
    if len(sys.argv) > 2:
        a = [int(x) for x in sys.argv[1:]]
        sizes = list(zip(a[0::2], a[1::2]))
    r = Renderer(devices=[0])
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    for W, H in sizes:
        img = torch.rand((H, W, 4), device="cuda", dtype=torch.float32)
        img[..., 3] = 1
        out = torch.empty_like(img)
        u8 = torch.empty((H, W, 3), device="cuda", dtype=torch.uint8)
        ms = timed(lambda: r.bloom_device(0.4, 25, W, H, img.data_ptr(), out.data_ptr()))
        ms8 = timed(lambda: r.to_srgb8_device(W, H, out.data_ptr(), u8.data_ptr()))
        copy = timed(lambda: out.copy_(img))
        alg = 2.0 * W * H * 16
        print(f"{W}x{H}: bloom {ms:.4f} ms = {alg / ms / 1e6:.0f} GB/s algorithmic; srgb8 map {ms8:.4f} ms; "
              f"torch copy of the frame {copy:.4f} ms = {alg / copy / 1e6:.0f} GB/s")
    r.close()


if __name__ == "__main__":
    main()
