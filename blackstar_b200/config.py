"""Scene / camera data model and YAML loader.

Host-side mirror of the reference's config layer (which stays Haskell in a real
integration; GHC is not available in this image):

* ``Camera`` / ``Scene`` / ``Config``  -- src/ConfigFile.hs:16-38
* YAML defaults                        -- src/ConfigFile.hs:66-79
* hue given in degrees, stored / 360   -- src/ConfigFile.hs:48-51
* ``prepare_scene`` (``--preview``)    -- app/Main.hs:93-103
"""
from __future__ import annotations

import ctypes
import dataclasses
from dataclasses import dataclass, field
from typing import Any, Dict, Tuple

import yaml


@dataclass
class Camera:
    """src/ConfigFile.hs:34-37"""
    position: Tuple[float, float, float]
    lookAt: Tuple[float, float, float]
    upVec: Tuple[float, float, float]
    fov: float


@dataclass
class Scene:
    """src/ConfigFile.hs:20-31 with the defaults of :66-79.

    ``diskColor`` is HSI with the hue already divided by 360 (``:51``).
    ``safeDistance`` is never read from YAML (``:67``) and is overwritten by
    ``render`` (src/Raytracer.hs:59-60), so it is not a field here.
    """
    stepSize: float = 0.3
    bloomStrength: float = 0.4
    bloomDivider: int = 25
    starIntensity: float = 0.7
    starSaturation: float = 0.7
    diskColor: Tuple[float, float, float] = (0.16, 0.1, 0.95)
    diskOpacity: float = 0.0
    diskInner: float = 3.0
    diskOuter: float = 12.0
    resolution: Tuple[int, int] = (1280, 720)
    supersampling: bool = False


@dataclass
class Config:
    """src/ConfigFile.hs:16-18"""
    scene: Scene
    camera: Camera


def _vec3(v: Any, what: str) -> Tuple[float, float, float]:
    if not isinstance(v, (list, tuple)) or len(v) != 3:
        raise ValueError(f"{what}: expected [x, y, z], got {v!r}")  # aeson: pattern match failure
    return (float(v[0]), float(v[1]), float(v[2]))


def camera_from_dict(d: Dict[str, Any]) -> Camera:
    """Generic FromJSON Camera: all four fields are mandatory (src/ConfigFile.hs:61)."""
    for k in ("position", "lookAt", "upVec", "fov"):
        if k not in d:
            raise ValueError(f"camera: key {k!r} not present")
    return Camera(_vec3(d["position"], "position"), _vec3(d["lookAt"], "lookAt"),
                  _vec3(d["upVec"], "upVec"), float(d["fov"]))


def scene_from_dict(d: Dict[str, Any]) -> Scene:
    """FromJSON Scene (src/ConfigFile.hs:66-79); unknown keys are ignored, as aeson does
    (this is why ``diskHSV`` in animations/default-ani.yaml has no effect)."""
    if not isinstance(d, dict):
        raise ValueError("scene: expected Object")
    s = Scene()
    if d.get("stepSize") is not None:
        s.stepSize = float(d["stepSize"])
    if d.get("bloomStrength") is not None:
        s.bloomStrength = float(d["bloomStrength"])
    if d.get("bloomDivider") is not None:
        s.bloomDivider = int(d["bloomDivider"])
    if d.get("starIntensity") is not None:
        s.starIntensity = float(d["starIntensity"])
    if d.get("starSaturation") is not None:
        s.starSaturation = float(d["starSaturation"])
    if d.get("diskColor") is not None:
        h, sa, i = _vec3(d["diskColor"], "diskColor")
        s.diskColor = (h / 360, sa, i)  # src/ConfigFile.hs:51
    if d.get("diskOpacity") is not None:
        s.diskOpacity = float(d["diskOpacity"])
    if d.get("diskInner") is not None:
        s.diskInner = float(d["diskInner"])
    if d.get("diskOuter") is not None:
        s.diskOuter = float(d["diskOuter"])
    if d.get("resolution") is not None:
        r = d["resolution"]
        if not isinstance(r, (list, tuple)) or len(r) != 2:
            raise ValueError(f"resolution: expected [w, h], got {r!r}")
        s.resolution = (int(r[0]), int(r[1]))
    if d.get("supersampling") is not None:
        s.supersampling = bool(d["supersampling"])
    return s


def config_from_dict(d: Dict[str, Any]) -> Config:
    if "scene" not in d or "camera" not in d:
        raise ValueError("config needs 'scene' and 'camera'")  # Generic FromJSON Config
    return Config(scene=scene_from_dict(d["scene"]), camera=camera_from_dict(d["camera"]))


def load_config(path: str) -> Config:
    """Data.Yaml.decodeFileEither (app/Main.hs:85)."""
    with open(path, "r", encoding="utf-8") as f:
        return config_from_dict(yaml.safe_load(f))


def prepare_scene(cfg: Config, preview: bool) -> Config:
    """app/Main.hs:93-103: --preview = long side 300, supersampling off, bloom off."""
    if not preview:
        return cfg
    w, h = cfg.scene.resolution
    res = 300
    new_res = (res, res * h // w) if w >= h else (res * w // h, res)
    scn = dataclasses.replace(cfg.scene, resolution=new_res, supersampling=False, bloomStrength=0.0)
    return Config(scene=scn, camera=cfg.camera)


def with_resolution(cfg: Config, w: int, h: int) -> Config:
    """Resolution override used by the BASELINE.json configs."""
    return Config(scene=dataclasses.replace(cfg.scene, resolution=(int(w), int(h))), camera=cfg.camera)


# ---------------------------------------------------------------- C ABI structs
class CCamera(ctypes.Structure):
    """bsb_camera (include/blackstar_b200.h)"""
    _fields_ = [("pos", ctypes.c_double * 3), ("look_at", ctypes.c_double * 3),
                ("up", ctypes.c_double * 3), ("fov", ctypes.c_double)]


class CScene(ctypes.Structure):
    """bsb_scene (include/blackstar_b200.h)"""
    _fields_ = [("step_size", ctypes.c_double), ("bloom_strength", ctypes.c_double),
                ("star_intensity", ctypes.c_double), ("star_saturation", ctypes.c_double),
                ("disk_hsi", ctypes.c_double * 3), ("disk_opacity", ctypes.c_double),
                ("disk_inner", ctypes.c_double), ("disk_outer", ctypes.c_double),
                ("bloom_divider", ctypes.c_int32), ("width", ctypes.c_int32),
                ("height", ctypes.c_int32), ("supersampling", ctypes.c_int32)]


def to_c(cfg: Config) -> Tuple[CCamera, CScene]:
    cam, scn = cfg.camera, cfg.scene
    c = CCamera((ctypes.c_double * 3)(*cam.position), (ctypes.c_double * 3)(*cam.lookAt),
                (ctypes.c_double * 3)(*cam.upVec), cam.fov)
    s = CScene(scn.stepSize, scn.bloomStrength, scn.starIntensity, scn.starSaturation,
               (ctypes.c_double * 3)(*scn.diskColor), scn.diskOpacity, scn.diskInner, scn.diskOuter,
               scn.bloomDivider, scn.resolution[0], scn.resolution[1], 1 if scn.supersampling else 0)
    return c, s
