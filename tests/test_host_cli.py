"""The C++ host (host/blackstar: mirror of app/Main.hs over the C ABI).  CPU tests cover the
YAML loader (against the Python mirror of ConfigFile.hs), the preview override, the PNG writer and
the start-up error behaviour; the GPU test renders through it and compares with the Python path."""
import json
import os
import subprocess

import numpy as np
import pytest

from blackstar_b200 import config, starmap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "blackstar")

BLOCK_STYLE = """\
camera:
    # comments and block style, as the reference's own scene files are written
    position:   [0, 1, -20]  # The position of the camera
    lookAt:     [2, 0, 0]
    upVec:      [-0.2, 1, 0]
    fov:        1.5           # The tangent of the view angle

scene:
    resolution: [1920, 1080]   # [width, height]
    bloomStrength: 0.15
    diskColor: [180, 0.1, 1.05]  # H: 0..360
    diskOpacity: 0.95
    supersampling: true
    diskHSV: [1, 2, 3]   # unknown keys are ignored
"""


@pytest.fixture(scope="module")
def exe():
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    return EXE


def _dump(exe, path, *flags):
    r = subprocess.run([exe, *flags, "--dump-config", path], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return json.loads(r.stdout)


def _as_dict(cfg):
    s, c = cfg.scene, cfg.camera
    return {"camera": {"position": list(c.position), "lookAt": list(c.lookAt), "upVec": list(c.upVec), "fov": c.fov},
            "scene": {"stepSize": s.stepSize, "bloomStrength": s.bloomStrength, "bloomDivider": s.bloomDivider,
                      "starIntensity": s.starIntensity, "starSaturation": s.starSaturation,
                      "diskColor": list(s.diskColor), "diskOpacity": s.diskOpacity, "diskInner": s.diskInner,
                      "diskOuter": s.diskOuter, "resolution": list(s.resolution), "supersampling": s.supersampling}}


def test_yaml_loader_matches_python_mirror(exe, scenes_dir, tmp_path):
    files = [os.path.join(scenes_dir, f) for f in sorted(os.listdir(scenes_dir)) if f.endswith(".yaml")]
    p = tmp_path / "block.yaml"
    p.write_text(BLOCK_STYLE)
    files.append(str(p))
    for f in files:
        for flags, preview in (((), False), (("-p",), True)):
            got = _dump(exe, f, *flags)
            want = _as_dict(config.prepare_scene(config.load_config(f), preview))
            assert got == want, f


def test_yaml_errors_are_reported_not_fatal(exe, tmp_path):
    p = tmp_path / "bad.yaml"
    p.write_text("scene: {}\n")
    r = subprocess.run([exe, "--dump-config", str(p)], capture_output=True, text=True)
    assert r.returncode == 2 and "camera" in r.stdout
    p.write_text("camera: {position: [0,0,1], lookAt: [0,0,0], upVec: [0,1,0]}\nscene: {}\n")
    r = subprocess.run([exe, "--dump-config", str(p)], capture_output=True, text=True)
    assert r.returncode == 2 and "fov" in r.stdout


def test_png_writer(exe, tmp_path):
    from PIL import Image
    out = str(tmp_path / "t.png")
    assert subprocess.run([exe, "--selftest-png", out]).returncode == 0
    im = Image.open(out)
    assert im.mode == "RGB" and im.size == (67, 31)
    a = np.array(im)
    y, x = np.mgrid[0:31, 0:67]
    np.testing.assert_array_equal(a[..., 0], x * 255 // 66)
    np.testing.assert_array_equal(a[..., 1], y * 255 // 30)
    np.testing.assert_array_equal(a[..., 2], (x * 7 + y * 13) & 255)


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_parallel_png_writer(exe, tmp_path, threads):
    # pigz-style banded deflate: one zlib stream, any decoder must read it back exactly
    from PIL import Image
    out = str(tmp_path / "p.png")
    assert subprocess.run([exe, "--selftest-png-parallel", out, str(threads)]).returncode == 0
    a = np.array(Image.open(out))
    w, h = 1031, 517
    assert a.shape == (h, w, 3)
    y, x = np.mgrid[0:h, 0:w]
    np.testing.assert_array_equal(a[..., 0], (x * 3 + y) & 255)
    np.testing.assert_array_equal(a[..., 2], ((x ^ y) * 5) & 255)
    lcg = np.uint32(12345)
    g = np.empty(w * h, dtype=np.uint8)
    state = 12345
    for k in range(w * h):
        state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
        g[k] = state >> 24
    np.testing.assert_array_equal(a[..., 1].reshape(-1), g)


def test_refuses_to_start_without_star_map(exe, scenes_dir):
    # app/Main.hs:46-50
    r = subprocess.run([exe, "-s", "/no/such/stars", os.path.join(scenes_dir, "default.yaml")], capture_output=True, text=True)
    assert r.returncode == 1 and "Error decoding star tree" in r.stdout


@pytest.mark.gpu
def test_cli_render_equals_python_path(exe, scenes_dir, tmp_path):
    from PIL import Image
    from blackstar_b200.render import Renderer
    cat = starmap.synthetic_catalogue(60000, seed=12)
    smap = tmp_path / "ppm.bin"
    smap.write_bytes(cat)
    scene = tmp_path / "s.yaml"
    cfg = config.with_resolution(config.load_config(os.path.join(scenes_dir, "default-aa.yaml")), 320, 180)
    scene.write_text(open(os.path.join(scenes_dir, "default-aa.yaml")).read().replace("[1920, 1080]", "[320, 180]"))
    out = tmp_path / "out"
    r = subprocess.run([exe, "-f", "-s", str(smap), "-o", str(out), str(scene)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Starmap successfully read." in r.stdout and "Everything done. Thank you!" in r.stdout
    got = np.array(Image.open(out / "s.png"))
    with Renderer(devices=[0]) as rd:
        rd.set_stars_ppm(cat)
        want = rd.do_render_srgb8(cfg)
    np.testing.assert_array_equal(got, want)
    # batch directory + preview naming (app/Main.hs:64-77,86)
    r = subprocess.run([exe, "-p", "-f", "-s", str(smap), "-o", str(out), scenes_dir], capture_output=True, text=True)
    assert r.returncode == 0 and "Batch mode progress: 9/9" in r.stdout
    assert sorted(f for f in os.listdir(out) if f.startswith("prev-")) == sorted(
        "prev-" + f[:-5] + ".png" for f in os.listdir(scenes_dir) if f.endswith(".yaml"))
    assert Image.open(out / "prev-default.png").size == (300, 168)


@pytest.mark.parametrize("style", [None, False, True])
def test_yaml_loader_fuzz_against_pyyaml(exe, tmp_path, style):
    # random scene files written by PyYAML in block, mixed and flow style; the C++ loader must agree
    # with the Python mirror of src/ConfigFile.hs on every one
    import yaml
    rng = np.random.default_rng(5 if style is None else int(style) + 6)
    keys = ["stepSize", "bloomStrength", "bloomDivider", "starIntensity", "starSaturation", "diskColor",
            "diskOpacity", "diskInner", "diskOuter", "resolution", "supersampling"]
    for k in range(25):
        scene = {}
        for key in keys:
            if rng.random() < 0.6:
                if key == "bloomDivider":
                    scene[key] = int(rng.integers(1, 60))
                elif key == "diskColor":
                    scene[key] = [float(rng.uniform(0, 359)), float(rng.uniform(0, 1)), float(rng.uniform(0, 1.2))]
                elif key == "resolution":
                    scene[key] = [int(rng.integers(1, 5000)), int(rng.integers(1, 5000))]
                elif key == "supersampling":
                    scene[key] = bool(rng.integers(2))
                else:
                    scene[key] = float(rng.choice([rng.uniform(0, 3), int(rng.integers(0, 20)), 1e-3, 2.5e2]))
        scene["notAKey"] = [1, 2, {"deep": "x"}]
        cam = {"position": [float(x) for x in rng.normal(0, 20, 3)], "lookAt": [int(x) for x in rng.integers(-5, 5, 3)],
               "upVec": [float(x) for x in rng.normal(0, 1, 3)], "fov": float(rng.uniform(0.2, 4))}
        doc = {"camera": cam, "scene": scene} if k % 2 else {"scene": scene, "camera": cam}
        p = tmp_path / f"f{k}.yaml"
        p.write_text(yaml.safe_dump(doc, default_flow_style=style, sort_keys=bool(k % 3)))
        got = _dump(exe, str(p))
        want = _as_dict(config.load_config(str(p)))
        assert got == want, p.read_text()
