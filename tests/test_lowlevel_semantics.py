"""FixedwingLowLevelEnv semantics in the fp64 oracle, cited against
/root/reference/envs/fixedwing_envs/fixedwing_lowlevel_env.py."""
import ctypes as C
import math

import numpy as np
import pytest

import pyflyt_drone_b200 as fw


@pytest.fixture(scope="module")
def fo(oracle_mod):
    return oracle_mod


def make(fo, n=6, seed=4, **over):
    cfg = fw.lowlevel(noise_ratio=0.0, **over)
    env = fo.OracleVecEnv(cfg.as_dict(), n, seed=seed)
    return cfg, env, env.reset()


def test_spaces_and_reset_observation(fo):
    cfg, env, obs = make(fo)
    assert env.obs_dim == 21 and env.act_dim == 6                     # :64-72
    # reset: Aviary state at the start pose, zero previous action, fresh target (:74-95); no warm-up steps
    assert np.allclose(obs[:, 0:6], 0.0) and np.allclose(obs[:, 6:9], [15.0, 0.0, 0.0]) and np.allclose(obs[:, 9:12], [0, 0, 10.0])
    assert np.allclose(obs[:, 12:18], 0.0)
    psi, h, v = obs[:, 18], obs[:, 19], obs[:, 20]
    assert (np.abs(psi) <= math.pi).all() and ((h >= 5) & (h <= 20)).all() and ((v >= 10) & (v <= 20)).all()
    assert len(np.unique(np.round(psi, 9))) == len(psi)               # per-env draws
    assert all(e.physics_steps == 0 and e.step_count == 0 for e in env.envs)


def test_one_aviary_step_per_env_step_and_mode_minus_one_mapping(fo):
    cfg, env, _ = make(fo)
    a = np.tile(np.array([0.3, -0.2, 0.1, -0.4, 0.25, 0.6]), (env.n, 1))
    obs, rew, flags, _ = env.step(a)
    e = env.envs[0]
    assert e.physics_steps == 2 and e.step_count == 1                 # env.step() once = 2 substeps at 240 Hz (:102-103)
    assert list(e.cmd[:]) == pytest.approx(list(a[0]))                # mode -1: the six channels go straight through
    assert np.allclose(obs[:, 12:18], a)                              # prev_action is this step's action (:99)
    k = cfg.dt / 0.05                                                 # first-order actuator lag over two substeps
    assert e.act[0] == pytest.approx(0.3 * (1 - (1 - k) ** 2), rel=1e-12)
    assert e.throttle == pytest.approx(0.6 * (1 - (1 - cfg.dt / 0.01) ** 2), rel=1e-12)   # thrust is NOT remapped here


def test_reward_is_the_tracking_error(fo):
    cfg, env, _ = make(fo)
    a = np.zeros((env.n, 6)); a[:, 5] = 0.7
    obs, rew, flags, _ = env.step(a)
    for i in range(env.n):
        yaw, alt = obs[i, 5], obs[i, 11]
        speed = np.linalg.norm(obs[i, 6:9])
        psi_ref, h_ref, v_ref = obs[i, 18:21]
        wrap = (psi_ref - yaw + math.pi) % (2 * math.pi) - math.pi
        expect = -(abs(wrap) + abs(h_ref - alt) + 0.5 * abs(v_ref - speed)) + 0.1      # :123-126
        assert rew[i] == pytest.approx(expect, rel=1e-12)
    assert (flags == 0).all()


def test_altitude_band_terminates_with_penalty_and_auto_resets(fo):
    cfg, env, _ = make(fo, n=2)
    st = env.get_state()
    st["pos"][0] = [3.0, 0.0, 0.9]          # below 1 m
    st["pos"][1] = [3.0, 0.0, 100.5]        # above 100 m
    env.set_state(st)
    tgt = [list(e.target_ref[:]) for e in env.envs]
    obs, rew, flags, term = env.step(np.zeros((2, 6)))
    assert list(flags) == [1 | 8, 1 | 8]                                 # terminated (+ our out-of-band info bit)
    for i in range(2):
        alt = term[i, 11]
        base = -(abs(((tgt[i][0] - term[i, 5] + math.pi) % (2 * math.pi)) - math.pi) + abs(tgt[i][1] - alt)
                 + 0.5 * abs(tgt[i][2] - np.linalg.norm(term[i, 6:9]))) + 0.1
        assert rew[i] == pytest.approx(base - 100.0, rel=1e-12)          # reward -= 100 (:130-132), not overwritten
        assert np.allclose(obs[i, 9:12], [0, 0, 10.0]) and env.envs[i].episode == 1 and env.envs[i].step_count == 0
        assert list(env.envs[i].target_ref[:]) != tgt[i]                # new target on reset


def test_truncation_on_the_2000th_step(fo):
    cfg, env, _ = make(fo, n=1)
    st = env.get_state()
    st["step_count"][0] = 1998
    env.set_state(st)
    a = np.zeros((1, 6)); a[0, 5] = 0.8
    _, _, f1, _ = env.step(a)
    assert f1[0] == 0                                                  # 1999 < 2000
    st = env.get_state(); st["pos"][0] = [0, 0, 10.0]; env.set_state(st)
    _, _, f2, _ = env.step(a)
    assert f2[0] == 2 and env.envs[0].episode == 1                     # `_episode_steps >= 2000` (:134-136)


def test_random_rollout_runs_and_uses_six_channels(fo):
    cfg, env, _ = make(fo, n=8)
    out = (C.c_double * 6)()
    fo.lib().fwo_random_action6(7, 3, 0, 5, out)
    a = np.array(out[:])
    assert (np.abs(a) <= 1).all() and len(np.unique(a)) == 6
    assert env.rollout_random(300) == 8 * 300
    assert all(0 <= e.step_count < 2000 for e in env.envs)
