"""State-machine traces of the env glue, run on the fp64 oracle (CPU).

Each test pins one subtlety of /root/reference/envs/fixedwing_envs/fixedwing_base_env.py:296-348 or of the
waypoint reward (/root/reference/envs/fixedwing_waypoint_objlock_env.py:278-300 and the upstream
FixedwingWaypointsEnv it was derived from): reward overwrite vs accumulate, the break test preceding
env.step(), the 3,602-nd call being the first truncated one and running exactly one inner iteration,
waypoint advance, zero padding of the flattened observation, SubprocVecEnv reset-on-done.
"""
import ctypes as C

import numpy as np
import pytest

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.config import FLAG_COLLISION, FLAG_COMPLETE, FLAG_OOB, FLAG_TERM, FLAG_TRUNC


@pytest.fixture(scope="module")
def fo(oracle_mod):
    return oracle_mod


def make(fo, n=1, **kw):
    cfg = fw.waypoints_v3(noise_ratio=0.0, **kw)
    env = fo.OracleVecEnv(cfg.as_dict(), n, seed=9)
    obs = env.reset()
    return cfg, env, obs


ZERO = np.zeros((1, 4))


def test_reset_runs_twenty_warmup_substeps_and_emits_28_floats(fo):
    cfg, env, obs = make(fo)
    e = env.envs[0]
    assert obs.shape == (1, 28)
    assert e.physics_steps == 20 and e.step_count == 0
    assert e.pos[0] == pytest.approx(20.0 * 20 / 240, rel=2e-2)      # ~1.66 m downrange
    assert np.all(obs[0, 12:16] == 0.0)                                # last action zeroed by begin_reset
    assert obs[0, 9:12] == pytest.approx(list(e.pos), abs=0)
    assert np.isinf(e.old_dist) and np.isfinite(e.new_dist)


def test_plain_step_reward_and_counters(fo):
    cfg, env, _ = make(fo)
    obs, rew, flags, _ = env.step(np.array([[0.1, -0.2, 0.3, 0.4]]))
    e = env.envs[0]
    assert rew[0] == pytest.approx(-0.1)           # sparse: only the per-step constant
    assert flags[0] == 0
    assert e.step_count == 1 and e.physics_steps == 28
    assert obs[0, 12:16] == pytest.approx([0.1, -0.2, 0.3, 0.4])   # raw action, thrust NOT remapped (:328)
    assert e.setpoint[3] == pytest.approx(0.4 / 2 + 0.5)            # aviary sees the remapped thrust (:330)


def test_dense_reward_accumulates_over_four_inner_iterations(fo):
    cfg, env, _ = make(fo, sparse_reward=0)
    e = env.envs[0]
    d_prev = e.new_dist
    # recompute by replaying substeps on a clone
    clone = fo.OEnv.from_buffer_copy(bytes(e))
    obs, rew, flags, _ = env.step(ZERO)
    L = fo.lib()
    clone.setpoint[3] = 0.5
    clone.last_action[:] = [0.0] * 6
    expect = -0.1
    for _ in range(4):
        clone.contact = 0
        L.fwo_substep(C.byref(env.cfg), C.byref(clone), 9)
        L.fwo_substep(C.byref(env.cfg), C.byref(clone), 9)
        d = np.linalg.norm(np.array(clone.targets[0][:]) - np.array(clone.pos[:]))
        expect += max(3.0 * (d_prev - d), 0.0) + 1.0 / d
        d_prev = d
    assert rew[0] == pytest.approx(expect, rel=1e-12)


def test_first_truncated_call_is_the_3602nd_and_runs_one_inner_iteration(fo):
    cfg, env, _ = make(fo)
    st = env.get_state()
    st["step_count"][0] = 3600       # the 3601-st call: step_count 3600 is NOT > max_steps
    st["pos"][0] = [0.0, 0.0, 50.0]
    env.set_state(st)
    _, rew, flags, _ = env.step(ZERO)
    e = env.envs[0]
    assert flags[0] == 0 and e.step_count == 3601 and e.physics_steps == 20 + 8
    ps = e.physics_steps
    _, rew, flags, term = env.step(ZERO)   # the 3602-nd call
    assert flags[0] == FLAG_TRUNC
    assert rew[0] == pytest.approx(-0.1)
    # vec semantics: env was reset after the truncated step
    assert e.step_count == 0 and e.episode == 1 and e.physics_steps == 20
    # the truncated step executed exactly one Aviary.step (2 substeps): terminal obs position moved ~2 substeps
    cfg2, env2, _ = make(fo)
    st2 = env2.get_state()
    st2["step_count"][0] = 3601
    st2["pos"][0] = [0.0, 0.0, 50.0]
    env2.set_state(st2)
    L = fo.lib()
    e2 = env2.envs[0]
    obs = np.zeros(28); r = C.c_double(); f = C.c_int32()
    a = np.zeros(4)
    L.fwo_step(C.byref(env2.cfg), C.byref(e2), 9, a.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p),
               C.byref(r), C.byref(f))
    assert f.value == FLAG_TRUNC and e2.physics_steps == 20 + 2 and e2.step_count == 3602


def test_collision_overwrites_reward_and_upstream_adds_dense_terms_on_top(fo):
    for early, sparse in ((0, 0), (1, 0), (0, 1)):
        cfg, env, _ = make(fo, sparse_reward=sparse, early_return_on_crash=early)
        st = env.get_state()
        st["pos"][0] = [0.0, 0.0, 0.05]
        st["vel"][0] = [20.0, 0.0, 0.0]
        env.set_state(st)
        _, rew, flags, term = env.step(ZERO)
        assert flags[0] & FLAG_TERM and flags[0] & FLAG_COLLISION
        if sparse or early:
            assert rew[0] == pytest.approx(-100.0)
        else:   # upstream Waypoints env has no early return: -100 then += dense terms of that iteration
            assert -100.0 < rew[0] < -99.0
        assert env.envs[0].episode == 1    # auto-reset happened
        assert term[0, 11] < 0.2           # terminal observation keeps the crash altitude


def test_break_precedes_env_step(fo):
    # a flag raised in inner iteration 0 means iterations 1..3 never run their substeps
    cfg, env, _ = make(fo)
    e = env.envs[0]
    st = env.get_state()
    st["pos"][0] = [0.0, 0.0, 0.05]
    env.set_state(st)
    L = fo.lib()
    obs = np.zeros(28); r = C.c_double(); f = C.c_int32(); a = np.zeros(4)
    L.fwo_step(C.byref(env.cfg), C.byref(e), 9, a.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p),
               C.byref(r), C.byref(f))
    assert f.value & FLAG_COLLISION and e.physics_steps == 20 + 2


def test_out_of_bounds_uses_strict_3d_norm(fo):
    cfg, env, _ = make(fo)
    st = env.get_state()
    st["pos"][0] = [60.0, 60.0, 55.0]   # |p| = 101.1 > 100
    st["vel"][0] = [1.0, 0.0, 0.0]
    env.set_state(st)
    _, rew, flags, _ = env.step(ZERO)
    assert flags[0] == FLAG_TERM | FLAG_OOB and rew[0] == pytest.approx(-100.0)
    cfg, env, _ = make(fo)
    st = env.get_state()
    st["pos"][0] = [50.0, 50.0, 50.0]   # |p| = 86.6
    st["vel"][0] = [1.0, 0.0, 0.0]
    env.set_state(st)
    _, rew, flags, _ = env.step(ZERO)
    assert flags[0] == 0


def test_waypoint_reach_advances_and_observation_shows_pre_advance_rows(fo):
    cfg, env, _ = make(fo)
    e = env.envs[0]
    st = env.get_state()
    tg = st["targets"].copy()
    # put target 0 two metres ahead of where the aircraft will be: reached (goal_reach 4) in iteration 0
    tg[0, 0] = st["pos"][0] + np.array([2.0, 0.0, 0.0])
    st["targets"] = tg
    env.set_state(st)
    obs, rew, flags, _ = env.step(ZERO)
    assert rew[0] == pytest.approx(100.0) or rew[0] == pytest.approx(100.0 - 0.0)
    assert e.target_idx == 1 and e.n_remaining == 7 and e.num_targets_reached == 1
    assert flags[0] == 0
    # later inner iterations recomputed the state: rows are now (target 1, target 2) in the body frame
    L = fo.lib()
    Rm = (C.c_double * 9)()
    L.fwo_quat_to_mat((C.c_double * 4)(*e.quat), Rm)
    R = np.array(Rm[:]).reshape(3, 3)
    d1 = R.T @ (tg[0, 1] - np.array(e.pos[:]))
    assert obs[0, 22:25] == pytest.approx(d1, abs=1e-6)


def test_last_waypoint_truncates_with_env_complete_and_zero_padding(fo):
    cfg, env, _ = make(fo, num_targets=1)
    e = env.envs[0]
    st = env.get_state()
    tg = st["targets"].copy()
    tg[0, 0] = st["pos"][0] + np.array([2.0, 0.0, 0.0])
    st["targets"] = tg
    env.set_state(st)
    obs, rew, flags, term = env.step(ZERO)
    assert flags[0] == FLAG_TRUNC | FLAG_COMPLETE
    assert rew[0] == pytest.approx(100.0)
    # terminal observation: computed before the advance -> row 0 is the reached target, row 1 zero padding
    assert np.linalg.norm(term[0, 22:25]) < 4.0
    assert np.all(term[0, 25:28] == 0.0)
    # obs returned to the agent is the fresh reset observation of the next episode
    assert e.episode == 1 and e.n_remaining == 1


def test_objlock_style_flags(fo):
    # in-tree variant: early return on crash, reaching the last waypoint does not truncate
    cfg, env, _ = make(fo, num_targets=1, early_return_on_crash=1, complete_truncates=0)
    e = env.envs[0]
    st = env.get_state()
    tg = st["targets"].copy()
    tg[0, 0] = st["pos"][0] + np.array([2.0, 0.0, 0.0])
    st["targets"] = tg
    env.set_state(st)
    obs, rew, flags, term = env.step(ZERO)
    assert flags[0] == FLAG_COMPLETE and e.n_remaining == 0
    assert np.all(obs[0, 22:28] == 0.0)


def test_vec_autoreset_and_rng_streams_are_independent_of_batch_split(fo):
    cfg = fw.waypoints_v3()
    a = fo.OracleVecEnv(cfg.as_dict(), 8, seed=4, env_id0=0)
    b = fo.OracleVecEnv(cfg.as_dict(), 4, seed=4, env_id0=4)
    oa, ob = a.reset(), b.reset()
    assert np.array_equal(oa[4:], ob)
    rng = np.random.default_rng(1)
    for _ in range(5):
        act = rng.uniform(-1, 1, (8, 4))
        ra = a.step(act)
        rb = b.step(act[4:])
        assert np.array_equal(ra[0][4:], rb[0]) and np.array_equal(ra[1][4:], rb[1])


def test_random_action_stream_matches_rollout(fo):
    cfg = fw.waypoints_v3()
    a = fo.OracleVecEnv(cfg.as_dict(), 3, seed=4)
    b = fo.OracleVecEnv(cfg.as_dict(), 3, seed=4)
    a.reset(); b.reset()
    a.rollout_random(6)
    L = fo.lib()
    for s in range(6):
        act = np.zeros((3, 4))
        for i in range(3):
            out = (C.c_double * 4)()
            L.fwo_random_action(4, i, b.envs[i].episode, b.envs[i].step_count, out)
            act[i] = out[:]
        assert np.all(np.abs(act) < 1.0)
        b.step(act)
    assert np.array_equal(a.get_state()["pos"], b.get_state()["pos"])
