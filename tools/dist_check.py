#!/usr/bin/env python
"""Run under torchrun: renders one scene with blackstar_b200.dist.DistributedFrame (one process per GPU,
row tiles -> all-to-all -> column bands -> shared host frame) and compares the host frame on rank 0 with
the 1-GPU render of the same scene.  Writes a small JSON report (tests/test_gpu_multi.py reads it)."""
import argparse
import dataclasses
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blackstar_b200 import config, starmap  # noqa: E402
from blackstar_b200.dist import DistributedFrame  # noqa: E402
from blackstar_b200.render import Renderer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, nargs=2, default=[640, 362])
    ap.add_argument("--scene", default="default-aa.yaml")
    ap.add_argument("--no-bloom", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    cfg = config.with_resolution(config.load_config(os.path.join(ROOT, "scenes", args.scene)), *args.res)
    if args.no_bloom:
        cfg = config.Config(scene=dataclasses.replace(cfg.scene, bloomStrength=0.0), camera=cfg.camera)
    stars = starmap.synthetic_stars(50000, seed=8)
    r = Renderer(devices=[local])
    r.set_stars(stars)
    frame = DistributedFrame(r, cfg, rank, world, device)
    frame.calibrate()
    got = np.array(frame.step_to_host(rgb8=False))
    got8 = np.array(frame.step_to_host(rgb8=True))
    got_again = np.array(frame.step_to_host(rgb8=False))
    rep = None
    if rank == 0:
        with Renderer(devices=[local]) as r1:
            r1.set_stars(stars)
            ref = r1.do_render(cfg)
            ref8 = r1.do_render_srgb8(cfg)
        d8 = np.abs(got8.astype(int) - ref8.astype(int))
        rep = {"world": world, "tiles": frame.tiles, "bands": frame.bands,
               "max_abs_err_f32": float(max(np.abs(got - ref).max(), np.abs(got_again - ref).max())),
               "srgb8_max_diff": int(d8.max()), "srgb8_frac_diff": float((d8 != 0).mean()), "launches": frame.launches}
        print(json.dumps(rep))
        if args.out:
            json.dump(rep, open(args.out, "w"))
    frame.close()
    r.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
