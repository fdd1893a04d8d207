#!/usr/bin/env python
"""bench.py -- Mrays/s of the hot path on BASELINE.json's headline workload.

Workload (config.workload): scenes/default-aa.yaml at 4096x4096 final resolution: 4x
supersampling (67.1 M traced rays/frame), synthetic 468 861-star catalogue looked up in the
k-d tree, bloom (BASELINE.json configs[2], north_star's target scene).  A "step" is one frame:
trace (row-tiled over the ranks) -> one NCCL gather on rank 0 -> bloom on rank 0.

  value : whole-job Mrays/s with everything resident in HBM (star tree uploaded once; frame
          left in HBM), CUDA events, max over ranks.
  e2e   : the same frame through the public API with HOST buffers: scene/camera structs
          in, pinned host framebuffer out (D2H inside the timed region).
  --impl reference : the reference's CPU path.  GHC is not in this image, so this is the C
          port of the reference (oracle/), all host threads, a bounded band of the same frame.

One process per GPU (torchrun) for --gpus > 1; scaling is STRONG (same frame, more GPUs).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENE = "default-aa.yaml"
RES = (4096, 4096)
FLOPS_PER_STEP = 156  # 141 (rk4 as written) + 15 (findColor), sqrt/div = 1 flop
DP_INSTR_PER_STEP = 61.4  # FP64-pipe instructions the kernel issues per RK4 step (ncu, incl. ray setup)
OTHER_INSTR_PER_STEP = 19.1  # all other instructions per RK4 step (ncu)
METRIC = "Mrays/sec on default.yaml at 4096x4096"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--res", type=int, nargs=2, default=list(RES), help="override the frame size (debug only)")
    ap.add_argument("--variant", type=int, default=None, help="trace schedule 0..3 (default: library default)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_workload(res):
    from blackstar_b200 import config
    cfg = config.load_config(os.path.join(ROOT, "scenes", SCENE))
    return config.with_resolution(cfg, res[0], res[1])


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU leg
class CpuArm:
    """The C port of the reference (oracle) on a band of rows of the SAME frame, all host
    threads (pthreads over rows, dynamic scheduling = massiv's Par on -N capabilities)."""

    def __init__(self, cfg, stars, threads=0):
        from oracle import pyoracle as po
        self.po, self.cfg, self.threads = po, cfg, threads
        self.tree = po.Tree(stars)
        self.W, self.H = cfg.scene.resolution
        self.ss = 4 if cfg.scene.supersampling else 1
        self.cores = threads if threads > 0 else (os.cpu_count() or 1)
        self.per_row = None

    @staticmethod
    def _thread_candidates():
        """Thread counts worth trying: what the OS reports, the scheduler affinity, and the cgroup
        CPU quota if the container has one (nproc can overstate what the job may use)."""
        n = os.cpu_count() or 1
        cands = {n}
        try:
            cands.add(len(os.sched_getaffinity(0)))
        except Exception:
            pass
        try:
            q, per = open("/sys/fs/cgroup/cpu.max").read().split()
            if q != "max":
                cands.add(max(1, int(round(int(q) / int(per)))))
        except Exception:
            pass
        if n >= 16:
            cands.add(n // 2)
        return sorted(c for c in cands if c >= 1)

    def calibrate(self):
        """Pick the thread count that gives the CPU port its best throughput (the baseline must
        not be sandbagged by oversubscription), then size the band."""
        best = None
        cands = [self.threads] if self.threads > 0 else self._thread_candidates()
        for nt in cands:
            cal = min(self.H, max(2, 2 * nt))                     # 2 rows per thread
            c0 = max(0, self.H // 2 - cal // 2)
            t = time.perf_counter()
            self.po.render(self.cfg, self.tree, c0, c0 + cal, nthreads=nt)
            per_row = max(time.perf_counter() - t, 1e-4) / cal
            if best is None or per_row < best[0]:
                best = (per_row, nt, cal)
        self.per_row, self.threads, self.cal = best
        self.cores = self.threads
        self.tried = cands

    def sample(self, seconds):
        """Returns (Mrays/s, description, rays, secs) for a band sized to ~`seconds`."""
        if self.per_row is None:
            self.calibrate()
        rows = int(max(self.cal, min(self.H, seconds / self.per_row)))
        r0 = max(0, self.H // 2 - rows // 2)
        t = time.perf_counter()
        _, steps = self.po.render(self.cfg, self.tree, r0, r0 + rows, nthreads=self.threads)
        dt = time.perf_counter() - t
        rays = rows * self.W * self.ss
        desc = (f"rows [{r0},{r0 + rows}) of the {self.W}x{self.H} frame ({rays} rays, {steps} RK4 steps, "
                f"{dt:.1f} s, {self.cores} threads = best of {self.tried}; os.cpu_count() = {os.cpu_count()})")
        return rays / dt / 1e6, desc, rays, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from blackstar_b200 import starmap
    cfg = load_workload(args.res)
    arm = CpuArm(cfg, starmap.synthetic_stars())
    total = args.steps + args.warmup
    per_step = max(1.0, min(20.0, 150.0 / max(1, total)))
    vals, last = [], None
    for i in range(total):
        v, desc, rays, dt = arm.sample(per_step)
        if i >= args.warmup:
            vals.append((rays, dt))
        last = desc
    rays = sum(r for r, _ in vals); secs = sum(d for _, d in vals)
    value = rays / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"scenes/{SCENE} at {args.res[0]}x{args.res[1]} (x4 supersampling, 468861-star "
                               "synthetic catalogue); each step = a bounded band of rows of that frame, no bloom",
                   "note": "GHC/stack are not installed here: this is the C port of the reference "
                           "(oracle/, gcc -O2 -ffp-contract=off, pthreads over rows), not the Haskell binary"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": arm.cores, "kind": "port", "sample": last},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from blackstar_b200 import starmap
    from blackstar_b200.dist import TiledFrame
    from blackstar_b200.render import Renderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: blackstar_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if "BENCH_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["BENCH_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)  # NCCL prints its version banner on stdout; we owe ONE JSON line
        dist.init_process_group("nccl", device_id=device)

    cfg = load_workload(args.res)
    W, H = cfg.scene.resolution
    ss = 4 if cfg.scene.supersampling else 1
    rays_per_frame = W * H * ss
    stars = starmap.synthetic_stars()  # N = 468 861, seed 20190412

    r = Renderer(devices=[local])
    r.set_stars(stars)
    if args.variant is not None:
        r.set_option("trace_variant", args.variant)
    frame = TiledFrame(r, cfg, rank, world, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- value: device-resident, K frames back to back
    frame.calibrate()            # N > 1: adaptive row tiles (untimed, like the warm-up)
    for _ in range(args.warmup):
        frame.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    frame.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        frame.step()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    launches = int(sum_over_ranks(float(frame.launches)))
    ms_per_step = ms_total / args.steps
    value = rays_per_frame / (ms_per_step * 1e-3) / 1e6

    # ---------------- dominant kernel (geodesic trace): live CUDA-event duration per launch
    trace_ms, steps_total, hits_total = [], 0, 0
    for _ in range(3):
        st = frame.step(want_stats=True)
        trace_ms.append(st["trace_ms"])
        steps_total, hits_total = st["steps"], st["star_hits"]
    barrier()
    k1_ms = sum(trace_ms) / len(trace_ms)              # rank 0's own launch (its tile)
    rk4_steps = int(sum_over_ranks(float(steps_total)))
    my_rows = frame.tiles[rank][1] - frame.tiles[rank][0]
    my_steps = steps_total
    k1_bytes = 16.0 * my_rows * W                      # one float4 store per OUTPUT pixel
    peaks, peak_kind = measured_peaks()
    fp64_peak = r.measure_fp64_peak() if rank == 0 else 0.0
    # bloom kernels, timed alone on rank 0
    bloom_ms = None
    if rank == 0 and cfg.scene.bloomStrength != 0:
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        scratch = torch.empty_like(frame.full)
        scratch.copy_(frame.full)
        torch.cuda.synchronize()
        reps = 5
        b0.record()
        for _ in range(reps):
            r.bloom_device(cfg.scene.bloomStrength, cfg.scene.bloomDivider, W, H, frame.full.data_ptr(), scratch.data_ptr())
        b1.record()
        torch.cuda.synchronize()
        bloom_ms = b0.elapsed_time(b1) / reps
        del scratch

    # ---------------- e2e: public API, host buffers, D2H inside the timed region
    # (a) synchronous: every frame is traced, gathered, bloomed and copied to pinned host memory
    #     before the next one starts -- the headline e2e;
    # (b) pipelined: same work, but the copy of frame i runs on a copy stream while frame i+1 is
    #     traced (double-buffered frames) -- what a batch / animation driver gets.
    host = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True) for _ in range(2)] if rank == 0 else None
    for _ in range(max(1, min(2, args.warmup))):
        frame.step_to_host(host[0] if rank == 0 else None)
    barrier()
    e0.record()
    for _ in range(args.steps):
        frame.step_to_host(host[0] if rank == 0 else None)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = rays_per_frame / (e2e_ms * 1e-3) / 1e6
    # (c) the image writeImg consumes: sRGB + 8 bit on the device, 3 bytes per pixel over PCIe
    host8 = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True) if rank == 0 else None
    frame.step_to_host_srgb8(host8)
    barrier()
    e0.record()
    for _ in range(args.steps):
        frame.step_to_host_srgb8(host8)
    e1.record()
    barrier()
    e2e8_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e8_value = rays_per_frame / (e2e8_ms * 1e-3) / 1e6
    frame.enable_double_buffering()
    for _ in range(2):
        frame.step_to_host_pipelined(host)
    frame.drain()
    barrier()
    e0.record()
    for _ in range(args.steps):
        frame.step_to_host_pipelined(host)
    frame.drain()
    e1.record()
    barrier()
    e2e_pipe_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_pipe_value = rays_per_frame / (e2e_pipe_ms * 1e-3) / 1e6

    # ---------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(cfg, stars)
        v, desc, _, _ = arm.sample(args.cpu_seconds)
        cpu = {"value": v, "unit": "Mrays/s", "cores": arm.cores, "kind": "port",
               "sample": desc + "; C port of the reference (GHC unavailable), all host threads"}

    if rank == 0:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "trace_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        k1_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9
        k1_tflops = FLOPS_PER_STEP * my_steps / (k1_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"scenes/{SCENE} at {W}x{H}: x4 supersampling ({rays_per_frame} rays/frame), "
                                   "468861-star synthetic catalogue (seed 20190412) in the k-d tree, bloom",
                       "parallelism": f"row tiles over {world} GPU(s), one NCCL gather to rank 0, bloom on rank 0",
                       "tiles": [list(t) for t in frame.tiles],
                       "l2": "each step writes a 268 MB frame (> 126 MB L2); the 22 MB star tree is L2-resident by design",
                       "trace_variant": args.variant},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 176, "d2h_bytes_per_step": W * H * 16,
                    "pipelined": {"value": e2e_pipe_value, "ms_per_step": e2e_pipe_ms,
                                  "note": "same bytes; D2H of frame i overlaps the trace of frame i+1"},
                    "srgb8": {"value": e2e8_value, "ms_per_step": e2e8_ms, "d2h_bytes_per_step": W * H * 3,
                              "note": "synchronous; sRGB + 8-bit map (writeImg, src/Raytracer.hs:23-32) on the "
                                      "device, the RGB8 image is what crosses PCIe"}},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "trace (geodesic RK4 + sky lookup + 2x2 supersample)",
                         "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak,
                         "traffic": traffic, "peak_kind": peak_kind, "launch_ms": k1_ms,
                         "note": "this kernel is FP64-issue-bound by construction (~35 kflop and 4 B of output "
                                 "per ray); see roofline_fp64 for the binding roofline"},
            "roofline_fp64": {"bound": "fp64", "achieved": k1_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": (k1_tflops / fp64_peak) if fp64_peak else None,
                              "flops_per_rk4_step": FLOPS_PER_STEP, "rk4_steps_per_frame": rk4_steps,
                              "peak_kind": "sustained DFMA micro-benchmark measured in this run (2 flops per FMA); "
                                           "1 ms bursts reach 36.6 (profiles/r01_fp64_pipe_probe.txt)",
                              "executed": {"dp_instr_per_rk4_step": DP_INSTR_PER_STEP,
                                           "pipe_frac": (DP_INSTR_PER_STEP * my_steps / (k1_ms * 1e-3)) / (fp64_peak * 1e12 / 2)
                                           if fp64_peak else None,
                                           "note": "the kernel executes 61.4 FP64 (+19.1 other) instructions per RK4 step (ncu, "
                                                   "profiles/r01_ncu_summary.json) where the reference as written "
                                                   "needs 156 flops, so the algorithmic frac can exceed 1; pipe_frac "
                                                   "is executed FP64 instructions / DFMA issue peak; an FP64 "
                                                   "instruction holds the issue port 2 cycles, any other 1, and "
                                                   "2*fp64+others reproduces the measured cycles to 3% (profiles/README.md)"}},
            "cpu_baseline": cpu,
        }
        if bloom_ms:
            bb = 2.0 * W * H * 16
            line["roofline_bloom"] = {"bound": "hbm", "kernel": "box3_transpose x2 (bloom)", "achieved": bb / (bloom_ms * 1e-3) / 1e9,
                                      "peak": hbm_peak, "unit": "GB/s", "frac": bb / (bloom_ms * 1e-3) / 1e9 / hbm_peak,
                                      "launch_ms": bloom_ms, "moved_bytes": 5.0 * W * H * 16,
                                      "note": "algorithmic bytes = read + write the frame once; the two-launch "
                                              "separable implementation moves 2.5x that"}
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
