"""debug / evaluation frames (FixedwingBaseEnv.render, fixedwing_base_env.py:350-369).  The CUDA
ray caster (csrc/fw_render.cu) against the fp64 oracle's fwo_render on the same injected state: class masks equal except on
silhouette-edge pixels, depth-buffer values and colours equal on the agreeing pixels."""
import numpy as np
import pytest

import pyflyt_drone_b200 as fw

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("preset,size", [("waypoint_objlock", (128, 128)), ("objlock_duck", (160, 120)), ("waypoints_v3", (96, 64))])
def test_rendered_frame_matches_the_oracle(oracle_mod, preset, size):
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    cfg = getattr(fw, preset)()
    n = 6
    env = FixedwingVecEnv(n, config=cfg, seed=5)
    orc = oracle_mod.OracleVecEnv(cfg.as_dict(), n, seed=5)
    env.reset(); orc.reset()
    orc.rollout_random(25)
    env.set_state(orc.get_state())
    W, H = size
    edge = total = 0
    classes = set()
    for i in range(n):
        g, o = env.render_layers(i, W, H), orc.render(i, W, H)
        assert g["rgba"].shape == (H, W, 4) and g["rgba"].dtype == np.uint8 and (g["rgba"][..., 3] == 255).all()
        same = g["seg"] == o["seg"]
        edge += int((~same).sum()); total += same.size
        classes |= set(np.unique(o["seg"]).tolist())
        assert np.abs(g["depth"][same] - o["depth"][same]).max() < 2e-5
        # colours: one count of rounding, except where a checker / cap boundary falls between the fp32 and fp64 hit points
        dc = np.abs(g["rgba"][same].astype(int) - o["rgba"][same].astype(int)).max(axis=1)
        assert (dc > 1).mean() < 5e-3
    assert edge <= 2e-3 * total, (edge, total)
    assert 0 in classes or -1 in classes
    if preset == "waypoint_objlock":
        assert any(1 <= c < 64 for c in classes)       # the duck or an obstacle shows in some frame
    if preset == "waypoints_v3":
        assert any(c >= 64 for c in classes)           # a waypoint sphere shows in some frame
    env.close()


def test_gym_env_render_and_vecenv_images():
    from pyflyt_drone_b200.gym_env import FixedwingWaypointsEnv
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    e = FixedwingWaypointsEnv(render_mode="rgb_array", seed=1)
    e.reset(seed=1)
    img = e.render()
    assert img.shape == (480, 480, 4) and img.dtype == np.uint8
    assert len(np.unique(img.reshape(-1, 4), axis=0)) > 3          # sky, two ground tones, shading
    e.close()
    with pytest.raises(ValueError):
        FixedwingWaypointsEnv().render()
    v = FixedwingVecEnv(3, preset="waypoint_objlock", seed=2)
    v.reset()
    imgs = v.get_images()
    assert len(imgs) == 3 and imgs[0].shape == (128, 128, 4)
    layers = v.render_layers(1)
    assert layers["depth"].min() >= 0.0 and layers["depth"].max() <= 1.0 and layers["seg"].min() >= -1
    v.close()
