"""The opt-in two-envs-per-thread step kernels (csrc/fw_pack.cuh, ``packed_pairs=1``): same agent step as the default
one-env kernels on the packed fp32x2 instructions of sm_100a.  Held to the same bar as the default path -- single steps
within 1e-4 of the fp64 oracle, flags and sparse rewards exact -- and to the default kernels themselves."""
import numpy as np
import pytest

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.config import FLAG_TERM, FLAG_TRUNC
from pyflyt_drone_b200.vec_env import FixedwingVecEnv

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    return float((np.abs(got - ref).max(axis=1) / np.maximum(np.abs(ref).max(axis=1), 1.0)).max())


@pytest.mark.parametrize("preset,kw", [("waypoints_v3", dict(noise_ratio=0.0)),
                                       ("waypoints_v3", dict(noise_ratio=0.02, sparse_reward=0, angle_repr=1)),
                                       ("lowlevel", dict(noise_ratio=0.0))])
def test_packed_single_step_parity_with_the_oracle(oracle_mod, preset, kw):
    cfg = fw.make_config(preset, packed_pairs=1, **kw)
    n = 511                                                    # odd: the last thread's second lane has no env
    env = FixedwingVecEnv(n, config=cfg, seed=7)
    orc = oracle_mod.OracleVecEnv(cfg.as_dict(), n, seed=7)
    og, oc = env.reset(), orc.reset()
    assert _rel(og, oc) < 1e-4
    rng = np.random.default_rng(1)
    n_done = 0
    for k in range(140):                                       # random actions put the first aircraft into the ground after ~100 steps
        env.set_state(orc.get_state())
        a = rng.uniform(-1, 1, (n, env.act_dim)).astype(np.float32)
        og, rg, fg, tg = (x.copy() for x in env.step_arrays(a))
        oc, rc, fc, tc = orc.step(a.astype(np.float64))
        assert np.array_equal(fg.astype(np.int32), fc), (k, np.nonzero(fg != fc))
        done = (fc & (FLAG_TERM | FLAG_TRUNC)) != 0
        n_done += int(done.sum())
        ang = slice(3, 6)
        if cfg.angle_repr == 0:                               # yaw / roll wrap at +-pi
            for x, y in ((og, oc), (tg, tc)):
                d = x[:, ang] - y[:, ang]
                x[:, ang] = y[:, ang] + (d + np.pi) % (2 * np.pi) - np.pi
        assert _rel(og, oc) < 1e-4, k
        if done.any():
            assert _rel(tg[done], tc[done]) < 1e-4, k
        assert np.abs(rg - rc).max() <= 1e-4 * max(1.0, np.abs(rc).max()), k
        if preset == "waypoints_v3":
            np.testing.assert_array_equal(env.last_targets_reached.astype(np.int32), orc.last_targets_reached)
    assert n_done > 0 or preset == "lowlevel"          # the tracking env's episodes outlast this test
    env.close()


def test_packed_and_default_kernels_agree_over_a_rollout():
    """Free-running 64 random-action steps (episodes end and restart inside): flags identical, states equal to fp32
    rounding amplified over the horizon."""
    n = 4096
    a = FixedwingVecEnv(n, config=fw.make_config("physics_only", packed_pairs=1), seed=3)
    b = FixedwingVecEnv(n, config=fw.make_config("physics_only", packed_pairs=0), seed=3)
    for _ in range(8):
        ra, fa = a.step_random(8, with_outputs=True)
        rb, fb = b.step_random(8, with_outputs=True)
        assert bool((fa == fb).all()) and bool(((ra - rb).abs() < 1e-5).all())
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa["episode"], sb["episode"]) and np.array_equal(sa["step_count"], sb["step_count"])
    assert np.abs(sa["pos"] - sb["pos"]).max() < 5e-2 and np.abs(sa["quat"] - sb["quat"]).max() < 5e-3
    a.close(); b.close()
