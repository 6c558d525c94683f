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
    """Free-running random-action rollouts of the two kernel families from the same seed: the first steps agree to fp32
    rounding; over 160 steps (episodes end and restart inside; fp32 rounding differences grow along each flight) the
    episode statistics agree statistically."""
    n = 4096
    a = FixedwingVecEnv(n, config=fw.make_config("physics_only", packed_pairs=1), seed=3)
    b = FixedwingVecEnv(n, config=fw.make_config("physics_only", packed_pairs=0), seed=3)
    a.step_random(4); b.step_random(4)
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa["physics_steps"], sb["physics_steps"]) and np.array_equal(sa["episode"], sb["episode"])
    assert np.abs(sa["pos"] - sb["pos"]).max() < 1e-3 and np.abs(sa["quat"] - sb["quat"]).max() < 1e-4
    assert np.abs(sa["vel"] - sb["vel"]).max() < 1e-2 and np.abs(sa["act"] - sb["act"]).max() < 1e-5
    a.step_random(156); b.step_random(156)
    ea, eb = a.episode_stats(), b.episode_stats()
    assert ea["episodes"] > 1000 and abs(ea["episodes"] - eb["episodes"]) <= 0.03 * eb["episodes"], (ea, eb)
    assert abs(ea["length_sum"] / ea["episodes"] - eb["length_sum"] / eb["episodes"]) < 2.0
    assert abs(ea["collisions"] - eb["collisions"]) <= 0.05 * max(eb["collisions"], 1.0) + 20
    a.close(); b.close()
