"""Animation model and frame generation (row N2 of SURVEY.md section 8f).

Mirror of src/Animation.hs (keyframes -> one Config per frame by linear interpolation of the
camera) and of the `animate` executable's file naming (app/Animate.hs:53-62, Util.padZero),
plus ``render_animation``: frames are independent, so frame i goes to GPU i mod N and every GPU
renders, blooms and tone-maps whole frames -- no collective ("replicas only").
"""
from __future__ import annotations

import math
import os
import threading
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import yaml

from .config import Camera, Config, Scene, camera_from_dict, scene_from_dict


@dataclass
class Keyframe:
    """src/Animation.hs:15-17"""
    camera: Camera
    time: float


@dataclass
class Animation:
    """src/Animation.hs:21-25 (interpolation: only 'linear' exists, :29-34 maps anything to it)"""
    scene: Scene
    nFrames: int
    keyframes: List[Keyframe]
    interpolation: str = "linear"


def animation_from_dict(d: Dict[str, Any]) -> Animation:
    for k in ("scene", "nFrames", "interpolation", "keyframes"):
        if k not in d:
            raise ValueError(f"animation: key {k!r} not present")  # Generic FromJSON Animation
    kfs = []
    for kf in d["keyframes"]:
        if "camera" not in kf or "time" not in kf:
            raise ValueError("keyframe needs 'camera' and 'time'")
        kfs.append(Keyframe(camera_from_dict(kf["camera"]), float(kf["time"])))
    return Animation(scene_from_dict(d["scene"]), int(d["nFrames"]), kfs, "linear")


def load_animation(path: str) -> Animation:
    with open(path, "r", encoding="utf-8") as f:
        return animation_from_dict(yaml.safe_load(f))


def validate_keyframes(kfs: Sequence[Keyframe]) -> Optional[str]:
    """src/Animation.hs:38-43; returns the reference's error string or None."""
    if len(kfs) < 2:
        return "Must have at least two keyframes"
    if kfs[0].time == 0 and kfs[-1].time == 1:
        return None
    return "First keyframe must have time == 0, last time == 1"


def _lerp(t: float, a, b):
    """interpolationFunction Linear (src/Animation.hs:81-86): a + t * (b - a), componentwise."""
    if isinstance(a, tuple):
        return tuple(x + t * (y - x) for x, y in zip(a, b))
    return a + t * (b - a)


def interpolate(frames: Sequence[Keyframe], t: float) -> Camera:
    """src/Animation.hs:61-79: the first consecutive pair with time f1 <= t < time f2; past the
    end the last keyframe is paired with itself one time unit later (so t' = 0)."""
    f1 = f2 = None
    t2 = None
    for a, b in zip(frames, frames[1:]):
        if a.time <= t < b.time:
            f1, f2, t2 = a, b, b.time
            break
    if f1 is None:
        f1 = f2 = frames[-1]
        t2 = frames[-1].time + 1
    tp = (t - f1.time) / (t2 - f1.time)
    c1, c2 = f1.camera, f2.camera
    return Camera(position=_lerp(tp, c1.position, c2.position), lookAt=_lerp(tp, c1.lookAt, c2.lookAt),
                  upVec=_lerp(tp, c1.upVec, c2.upVec), fov=_lerp(tp, c1.fov, c2.fov))


def _check_n_frames(n_frames: int):
    # the reference divides by (nFrames - 1) and produces Infinity / NaN cameras for nFrames = 1
    # without complaint (src/Animation.hs:62-66); an explicit error is more useful than a NaN scene
    if n_frames < 2:
        raise ValueError("nFrames must be at least 2 (the reference interpolates over nFrames - 1 intervals)")


def generate_frames(anim: Animation) -> List[Config]:
    """src/Animation.hs:45-52: nFrames points i / (nFrames - 1), keyframes sorted by time (stable)."""
    _check_n_frames(anim.nFrames)
    stepsize = 1.0 / (anim.nFrames - 1)
    frames = sorted(anim.keyframes, key=lambda k: k.time)
    return [Config(scene=anim.scene, camera=interpolate(frames, i * stepsize)) for i in range(anim.nFrames)]


def pad_zero(max_val: int, val: int) -> str:
    """Util.padZero (src/Util.hs:43-48), including its quirks: nDigits 0 = floor(logBase 10 0) + 1
    underflows, the zero count goes negative and frame 0 is written WITHOUT padding; and GHC's
    ``logBase 10 x = log x / log 10`` is 2.9999999999999996 for 1000 (and short for 1e6, 1e9, ...), so
    such a value counts one digit less than it has (frame 1000 of 1001+ is named ``_01000``)."""
    def n_digits(x: int) -> int:
        return math.floor(math.log(float(x)) / math.log(10.0)) + 1
    if val <= 0 or max_val <= 0:
        return str(val)
    return "0" * max(0, n_digits(max_val) - n_digits(val)) + str(val)


def frame_filename(basename: str, n_frames: int, idx: int, ext: str = ".yaml") -> str:
    """app/Animate.hs:55-56"""
    return f"{basename}_{pad_zero(n_frames - 1, idx)}{ext}"


def config_to_yaml(cfg: Config) -> str:
    """What `encode frame` writes (ToJSON instances of src/ConfigFile.hs:45-46,53-54,58-84)."""
    s, c = cfg.scene, cfg.camera
    d = {"scene": {"safeDistance": 0.0, "stepSize": s.stepSize, "bloomStrength": s.bloomStrength,
                   "bloomDivider": s.bloomDivider, "starIntensity": s.starIntensity,
                   "starSaturation": s.starSaturation,
                   "diskColor": [360 * s.diskColor[0], s.diskColor[1], s.diskColor[2]],
                   "diskOpacity": s.diskOpacity, "diskInner": s.diskInner, "diskOuter": s.diskOuter,
                   "resolution": list(s.resolution), "supersampling": s.supersampling},
         "camera": {"position": list(c.position), "lookAt": list(c.lookAt), "upVec": list(c.upVec), "fov": c.fov}}
    return yaml.safe_dump(d, default_flow_style=None, sort_keys=False)


def write_frames(anim: Animation, basename: str, outdir: str) -> List[str]:
    """The `animate` executable: one scene YAML per frame (app/Animate.hs:53-62)."""
    os.makedirs(outdir, exist_ok=True)
    paths = []
    for idx, cfg in enumerate(generate_frames(anim)):
        p = os.path.join(outdir, frame_filename(basename, anim.nFrames, idx))
        with open(p, "w", encoding="utf-8") as f:
            f.write(config_to_yaml(cfg))
        paths.append(p)
    return paths


def render_animation(renderers: Sequence, anim: Animation, outdir: str, basename: str = "frame",
                     frames: Optional[Sequence[int]] = None, write_png: bool = True) -> Dict[int, Any]:
    """Render the animation with frames sharded over ``renderers`` (one 1-GPU Renderer each):
    frame i -> renderers[i % N].  Each GPU renders + blooms + tone-maps whole frames; PNGs are
    written by the worker that rendered them.  Returns {frame index: uint8 image}."""
    from .render import write_img
    cfgs = generate_frames(anim)
    todo = list(range(len(cfgs))) if frames is None else list(frames)
    os.makedirs(outdir, exist_ok=True)
    out: Dict[int, Any] = {}
    errors: List[BaseException] = []

    def work(k: int):
        try:
            r = renderers[k]
            for i in todo[k::len(renderers)]:
                img8 = r.do_render_srgb8(cfgs[i])
                if write_png:
                    write_img(img8, os.path.join(outdir, frame_filename(basename, anim.nFrames, i, ".png")))
                out[i] = img8
        except BaseException as e:  # surfaced after join
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(renderers))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out
