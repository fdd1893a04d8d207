"""Animation.hs / Animate.hs mirror (row N2): interpolation, frame naming, YAML round trip;
on the GPU: frame-sharded batch rendering equals frame-by-frame rendering."""
import os

import numpy as np
import pytest

from blackstar_b200 import animation, config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ANI = os.path.join(ROOT, "animations", "default-ani.yaml")


def test_default_animation_loads_and_ignores_diskHSV():
    a = animation.load_animation(ANI)
    assert a.nFrames == 375 and len(a.keyframes) == 2
    assert a.scene.diskColor == (0.16, 0.1, 0.95)       # diskHSV is not a key the parser reads (S10)
    assert animation.validate_keyframes(a.keyframes) is None


def test_generate_frames_endpoints_and_linearity():
    a = animation.load_animation(ANI)
    fr = animation.generate_frames(a)
    assert len(fr) == 375
    k0, k1 = a.keyframes
    assert fr[0].camera == k0.camera
    # t = 374 * (1/374) may round to 1 or just below it; either way the camera is the last keyframe
    np.testing.assert_allclose(fr[-1].camera.position, k1.camera.position, atol=1e-12)
    assert abs(fr[-1].camera.fov - k1.camera.fov) < 1e-12
    mid = fr[187]  # t = 187/374 = 0.5
    np.testing.assert_allclose(mid.camera.position, [(3 - 15) / 2, 2, -20], atol=1e-12)
    np.testing.assert_allclose(mid.camera.lookAt, [3, -1, 0], atol=1e-12)
    assert abs(mid.camera.fov - 1.75) < 1e-12
    assert all(f.scene is a.scene for f in fr)


def test_keyframe_validation_messages():
    a = animation.load_animation(ANI)
    assert animation.validate_keyframes(a.keyframes[:1]) == "Must have at least two keyframes"
    assert animation.validate_keyframes([]) == "Must have at least two keyframes"
    bad = [animation.Keyframe(a.keyframes[0].camera, 0.1), a.keyframes[1]]
    assert animation.validate_keyframes(bad) == "First keyframe must have time == 0, last time == 1"


def test_three_keyframes_unsorted_input():
    a = animation.load_animation(ANI)
    c0, c1 = a.keyframes[0].camera, a.keyframes[1].camera
    kfs = [animation.Keyframe(c0, 0.0), animation.Keyframe(c0, 1.0), animation.Keyframe(c1, 0.5)]
    anim = animation.Animation(a.scene, 5, kfs)
    fr = animation.generate_frames(anim)      # sorted by time: c0 @0, c1 @0.5, c0 @1
    np.testing.assert_allclose(fr[2].camera.position, c1.position)
    np.testing.assert_allclose(fr[1].camera.position, [(x + y) / 2 for x, y in zip(c0.position, c1.position)])
    np.testing.assert_allclose(fr[3].camera.position, fr[1].camera.position)


def test_pad_zero_and_its_quirk():
    assert animation.pad_zero(374, 7) == "007" and animation.pad_zero(374, 42) == "042"
    assert animation.pad_zero(374, 374) == "374" and animation.pad_zero(9, 3) == "3"
    assert animation.pad_zero(374, 0) == "0"          # the reference does not pad frame 0 (src/Util.hs:45)
    # logBase 10 1000 = log 1000 / log 10 = 2.9999999999999996 in GHC: 1000 counts as a 3-digit number
    assert animation.pad_zero(1500, 1000) == "01000" and animation.pad_zero(1500, 999) == "0999"
    assert animation.pad_zero(1000, 7) == "007"       # ... also as the maximum
    with pytest.raises(ValueError):
        a = animation.load_animation(ANI)
        a.nFrames = 1
        animation.generate_frames(a)
    assert animation.frame_filename("default-ani", 375, 1) == "default-ani_001.yaml"


def test_written_frames_round_trip(tmp_path):
    a = animation.load_animation(ANI)
    a.nFrames = 4
    paths = animation.write_frames(a, "ani", str(tmp_path))
    assert [os.path.basename(p) for p in paths] == ["ani_0.yaml", "ani_1.yaml", "ani_2.yaml", "ani_3.yaml"]
    want = animation.generate_frames(a)
    for p, w in zip(paths, want):
        got = config.load_config(p)
        assert got.scene == w.scene
        np.testing.assert_allclose(got.camera.position, w.camera.position, rtol=1e-15)
        assert got.camera.fov == w.camera.fov


@pytest.mark.gpu
def test_frame_sharded_batch_equals_single_renders(tmp_path):
    import torch
    from blackstar_b200 import starmap
    from blackstar_b200.render import Renderer
    a = animation.load_animation(ANI)
    a.nFrames = 6
    a.scene.resolution = (160, 90)
    stars = starmap.synthetic_stars(30000, seed=4)
    n = max(1, min(8, torch.cuda.device_count()))   # every GPU of the box: frame i -> GPU i mod n
    rs = [Renderer(devices=[k]) for k in range(n)]
    for r in rs:
        r.set_stars(stars)
    got = animation.render_animation(rs, a, str(tmp_path), "ani")
    assert sorted(got) == list(range(6))
    assert os.path.exists(tmp_path / "ani_0.png") and os.path.exists(tmp_path / "ani_5.png")
    for i, cfg in enumerate(animation.generate_frames(a)):
        np.testing.assert_array_equal(got[i], rs[0].do_render_srgb8(cfg))
    assert not np.array_equal(got[0], got[5])
    for r in rs:
        r.close()
