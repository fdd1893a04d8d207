#!/usr/bin/env python
"""Hardware numbers for BASELINE.json configs[3] and [4] (not the headline: these go to profiles/).

  C4  scenes/lensing-disk.yaml at 8192x8192 (x4 supersampling, 268 M rays) row-tiled over N GPUs.
      Both arms of the multi-GPU pipeline: one process per GPU (bench.py under torchrun) and one process
      driving all GPUs (bench.py --inlib), plus the library's own per-stage clock (bsb_stats).
  C5  animations/default-ani.yaml with nFrames = 600 at 1920x1080 (x4 supersampling), frames sharded over
      N GPUs the way the reference's batch mode would consume them (app/Main.hs:64-77: a directory of
      scene files): `animate` writes the 600 YAMLs, they are dealt round-robin into N directories, and N
      copies of host/blackstar (one per GPU, CUDA_VISIBLE_DEVICES) render + bloom + sRGB8 on the device,
      encode PNGs on the host cores and write them.  Reported: frames/s with and without the PNG encoder,
      so the bottleneck (GPU, PCIe or zlib) has a name.

usage: python tools/bench_c4_c5.py [--gpus N] [--out profiles/r02_c4_c5.json] [--frames 600] [--skip-c4] [--skip-c5]
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def last_json_line(text):
    for ln in reversed(text.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    raise RuntimeError("no JSON line in: " + text[-2000:])


def run_c4(n, steps, warmup):
    out = {}
    env = dict(os.environ)
    common = ["--scene", "lensing-disk.yaml", "--res", "8192", "8192", "--steps", str(steps), "--warmup", str(warmup), "--no-cpu-baseline"]
    if n > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
               "--master-port", "29533", os.path.join(ROOT, "bench.py"), "--gpus", str(n)] + common
    else:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1"] + common
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1800)
    if r.returncode != 0:
        out["one_process_per_gpu"] = {"error": r.stderr[-1500:]}
    else:
        out["one_process_per_gpu"] = last_json_line(r.stdout)
    if n > 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--inlib"] + common
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1800)
        out["one_process_all_gpus"] = last_json_line(r.stdout) if r.returncode == 0 else {"error": r.stderr[-1500:]}
    # the library's own per-stage clock, in-library arm
    from blackstar_b200 import config, starmap
    from blackstar_b200.render import Renderer
    import numpy as np
    cfg = config.with_resolution(config.load_config(os.path.join(ROOT, "scenes", "lensing-disk.yaml")), 8192, 8192)
    with Renderer(n_gpus=n) if n > 1 else Renderer(devices=[0]) as rr:
        rr.set_stars(starmap.synthetic_stars())
        buf = np.zeros((8192, 8192, 3), dtype=np.uint8)
        for _ in range(3):
            rr.do_render_srgb8(cfg, out=buf)
        best = None
        for _ in range(3):
            rr.do_render_srgb8(cfg, out=buf)
            if best is None or rr.last_stats["total_ms"] < best["total_ms"]:
                best = dict(rr.last_stats)
    out["bsb_stats_srgb8"] = best
    out["bsb_stats_note"] = ("trace_ms = slowest GPU's tile; gather_ms = the all-to-all as seen by the slowest GPU (includes waiting "
                            "for the slowest tile); bloom_ms = horizontal + vertical bloom launches (T = 1024 threads per line at "
                            "8192); d2h_ms = the parallel copies into the caller's (pageable) buffer")
    return out


def run_c5(n, n_frames, keep_png, png_level=6):
    from blackstar_b200 import animation, config, starmap
    from blackstar_b200.render import Renderer
    work = tempfile.mkdtemp(prefix="bsb_c5_")
    rep = {"frames": n_frames, "gpus": n, "resolution": [1920, 1080], "png_zlib_level": png_level}
    try:
        anim = animation.load_animation(os.path.join(ROOT, "animations", "default-ani.yaml"))
        anim.nFrames = n_frames
        anim.scene.resolution = (1920, 1080)
        anim.scene.supersampling = True
        t0 = time.perf_counter()
        paths = animation.write_frames(anim, "default-ani", os.path.join(work, "all"))
        rep["write_yaml_s"] = time.perf_counter() - t0
        for k in range(n):
            os.makedirs(os.path.join(work, f"in{k}"))
        for i, p in enumerate(paths):
            shutil.move(p, os.path.join(work, f"in{i % n}", os.path.basename(p)))
        ppm = os.path.join(work, "stars.ppm")
        with open(ppm, "wb") as f:
            f.write(starmap.synthetic_catalogue())
        exe = os.path.join(ROOT, "host", "blackstar")
        # ---- (a) the whole thing: N copies of the blackstar executable, batch mode, PNGs written
        procs = []
        t0 = time.perf_counter()
        for k in range(n):
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(k), BSB_PNG_LEVEL=str(png_level),
                       BSB_PNG_THREADS=str(max(1, (os.cpu_count() or 8) // n)))
            procs.append(subprocess.Popen([exe, "-f", "-s", ppm, "-o", os.path.join(work, f"out{k}"), os.path.join(work, f"in{k}")],
                                          env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True))
        errs = [p.communicate()[1] for p in procs]
        dt = time.perf_counter() - t0
        n_png = sum(len([f for f in os.listdir(os.path.join(work, f"out{k}")) if f.endswith(".png")]) for k in range(n)
                    if os.path.isdir(os.path.join(work, f"out{k}")))
        rep["with_png"] = {"seconds": dt, "frames_per_s": n_png / dt, "pngs_written": n_png, "rc": [p.returncode for p in procs],
                           "includes": "process start, CUDA context + star tree per process (~1 s), YAML parse, render, bloom, sRGB8, "
                                       "D2H, PNG deflate on the host cores, file write",
                           "stderr_tail": [e[-200:] for e in errs if e]}
        if keep_png:
            os.makedirs(keep_png, exist_ok=True)
            for name in ("default-ani_000.png", f"default-ani_{n_frames // 2:03d}.png"):
                for k in range(n):
                    p = os.path.join(work, f"out{k}", name)
                    if os.path.exists(p):
                        shutil.copy(p, os.path.join(keep_png, "c5_" + name))
        # ---- (a') ONE copy of the executable over all GPUs: every frame is row-tiled over the N GPUs of one ctx
        # (bsb_create(0)), one background writer deflates with all host cores while the next frame is traced
        for k in range(n):
            for f in os.listdir(os.path.join(work, f"in{k}")):
                shutil.copy(os.path.join(work, f"in{k}", f), os.path.join(work, "all", f))
        env = dict(os.environ, BSB_PNG_LEVEL=str(png_level))
        env.pop("CUDA_VISIBLE_DEVICES", None)
        t0 = time.perf_counter()
        pr = subprocess.run([exe, "-f", "-s", ppm, "-o", os.path.join(work, "out_all"), os.path.join(work, "all")], env=env,
                            stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        dt = time.perf_counter() - t0
        n_png = len([f for f in os.listdir(os.path.join(work, "out_all")) if f.endswith(".png")]) if os.path.isdir(os.path.join(work, "out_all")) else 0
        rep["with_png_one_process"] = {"seconds": dt, "frames_per_s": n_png / dt, "pngs_written": n_png, "rc": pr.returncode,
                                       "includes": "ONE process, one ctx over all GPUs (each frame row-tiled over them), one "
                                                   "background PNG writer deflating with every host core", "stderr_tail": pr.stderr[-200:]}
        # ---- (b) the device side alone: same frames, same sharding, RGB8 into host memory, no encoder
        cfgs = animation.generate_frames(anim)
        rs = [Renderer(devices=[k]) for k in range(n)]
        stars = starmap.synthetic_stars()
        for r in rs:
            r.set_stars(stars)
            r.do_render_srgb8(cfgs[0])

        def work_fn(k):
            for i in range(k, n_frames, n):
                rs[k].do_render_srgb8(cfgs[i])
        th = [threading.Thread(target=work_fn, args=(k,)) for k in range(n)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        rep["render_only"] = {"seconds": dt, "frames_per_s": n_frames / dt, "mrays_per_s": n_frames * 1920 * 1080 * 4 / dt / 1e6,
                              "includes": "bsb_render_full_srgb8 per frame into pageable host memory (render, bloom, sRGB8, D2H), "
                                          "one host thread per GPU, no PNG"}
        for r in rs:
            r.close()
        a, b = rep["with_png"]["frames_per_s"], rep["render_only"]["frames_per_s"]
        rep["bottleneck"] = ("PNG deflate (zlib) on the host cores" if a < 0.7 * b else "the GPUs") + \
            f": {a:.1f} frames/s end to end vs {b:.1f} frames/s for the device side alone; os.cpu_count() = {os.cpu_count()}"
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_c4_c5.json"))
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--skip-c4", action="store_true")
    ap.add_argument("--skip-c5", action="store_true")
    ap.add_argument("--keep-png", default="")
    a = ap.parse_args()
    import torch
    n = a.gpus or torch.cuda.device_count()
    rep = {"gpus": n, "host_cpus": os.cpu_count()}
    if not a.skip_c4:
        rep["C4_lensing_disk_8192"] = run_c4(n, a.steps, a.warmup)
    if not a.skip_c5:
        rep["C5_animation"] = run_c5(n, a.frames, a.keep_png, 6)
        rep["C5_animation_png_level1"] = run_c5(n, a.frames, "", 1)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rep, open(a.out, "w"), indent=1)
    brief = {"gpus": n}
    if "C4_lensing_disk_8192" in rep:
        c4 = rep["C4_lensing_disk_8192"]
        for k in ("one_process_per_gpu", "one_process_all_gpus"):
            if k in c4 and "value" in c4[k]:
                brief["C4 " + k] = {"value_Mrays_s": c4[k]["value"], "ms": c4[k]["ms_per_step"], "e2e_ms": c4[k]["e2e"]["ms_per_step"]}
        brief["C4 stats"] = c4.get("bsb_stats_srgb8")
    if "C5_animation" in rep:
        brief["C5"] = {k: rep["C5_animation"].get(k) for k in ("with_png", "with_png_one_process", "render_only", "bottleneck")}
        brief["C5 zlib level 1"] = rep["C5_animation_png_level1"].get("with_png")
    print(json.dumps(brief, indent=1))


if __name__ == "__main__":
    main()
