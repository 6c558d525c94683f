"""CPU tests of the duck-only lock/strike task on the fp64 oracle: observation layout, vision
history and deltas, the lock counter, strike, crash and truncation rules of
/root/reference/envs/fixedwing_objlock_env.py:177-459 with the training configuration of
/root/reference/train/train_objlock.py:27-146, seen through the flattening of
/root/reference/envs/flatten_objlock_env.py:41-46."""
import math

import numpy as np
import pytest

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.config import FLAG_COLLISION, FLAG_COMPLETE, FLAG_OOB, FLAG_STRIKE, FLAG_TERM, FLAG_TRUNC

ATT = 22          # euler attitude block: 12 + action 4 + aux 6
VIS = ATT + 3     # first history row
ZERO = np.zeros((1, 4))


@pytest.fixture(scope="module")
def fo(oracle_mod):
    return oracle_mod


def make(fo, n=1, seed=2, **kw):
    cfg = fw.make_config("objlock_duck", wind={"enabled": False}, noise_ratio=0.0, **kw)
    env = fo.OracleVecEnv(cfg.as_dict(), n, seed=seed)
    obs = env.reset()
    return cfg, env, obs


def aim_at_duck(env, dist, height=30.0, i=0):
    """Level flight at `height`, `dist` metres (horizontally) short of the duck and heading straight at it."""
    st = env.get_state()
    d = st["duck"][i]
    yaw = 0.3
    st["pos"][i] = [d[0] - dist * math.cos(yaw), d[1] - dist * math.sin(yaw), height]
    st["quat"][i] = [0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2)]
    st["vel"][i] = [20 * math.cos(yaw), 20 * math.sin(yaw), 0.0]
    st["omega"][i] = 0.0
    return st


def test_preset_matches_the_training_script():
    cfg = fw.make_config("objlock_duck")
    assert cfg.obs_dim == 56 and cfg.task == 4
    assert cfg.start_pos == [0.0, 0.0, 100.0] and cfg.dome == 200.0 and cfg.max_steps == 1800
    assert cfg.cam_interval_substeps == 24 and cfg.cam_mode == 1 and cfg.cam_offset == [0.8, 0.0, 0.12]
    assert (cfg.lock_hold_steps, cfg.strike_dist, cfg.strike_reward, cfg.lock_step_reward, cfg.approach_scale) == \
        (5, 10.0, 400.0, 0.2, 0.1)
    assert (cfg.duck_dist_scale, cfg.lock_center_radius, cfg.centering_scale, cfg.visible_step_reward,
            cfg.area_reward_scale, cfg.lock_lost_penalty, cfg.approach_clip, cfg.lock_decay_steps) == \
        (1.0, 0.55, 3.0, 2.0, 5.0, 0.5, 2.0, 1)
    assert cfg.wind_mode == 2 and cfg.wind_base_hi == [10.0, 10.0, 0.10]
    assert fw.make_config("objlock_duck", vision_use_deltas=0, vision_hist_len=2, angle_repr=1).obs_dim == 23 + 3 + 18


def test_reset_observation_layout(fo):
    cfg, env, obs = make(fo, n=64)
    st = env.get_state()
    assert obs.shape == (64, 56)
    assert np.allclose(st["duck"][:, 2], 0.05) and np.all(np.abs(st["duck"][:, :2]) <= 100.0)
    assert st["duck"][:, 0].std() > 30 and st["duck"][:, 1].std() > 30            # U(-dome/2, dome/2)
    # target_vector = R^T (duck - pos): its norm is the distance to the duck
    assert np.allclose(np.linalg.norm(obs[:, ATT:ATT + 3], axis=1), np.linalg.norm(st["duck"] - st["pos"], axis=1))
    # no frame yet (first capture at physics step 24 > 20 warm-up steps): one history row holding the defaults of
    # _build_vision_vector(0.0) -- cx = cy = 0.5, steps_since_seen 60/60 -- older rows and deltas zero
    row0 = np.array([0, 0.5, 0.5, 0, 0, 1.0, 0, 0, 0])
    assert np.allclose(obs[:, VIS:VIS + 9], row0) and np.all(obs[:, VIS + 9:] == 0.0)
    assert np.all(st["ol_i"][:, 5] == 1) and np.all(st["ol_i"][:, 3] == 0)        # history filled 1, no camera frame
    assert np.all(st["ol_i"][:, 8] == 0)                                           # num_obstacles = 0


def test_history_shifts_once_per_inner_iteration_and_frames_arrive_every_24_substeps(fo):
    cfg, env, _ = make(fo)
    env.set_state(aim_at_duck(env, 120.0))
    # physics_steps is 20 after the warm-up; Aviary steps end at 22, 24 (capture), 26, 28 within the first env step
    obs, r, f, _ = env.step(ZERO)
    h = obs[0, VIS:VIS + 27].reshape(3, 9)
    assert np.all(h[:, 0] == 1.0)                         # rows of inner iterations 4, 3, 2 all saw the frame of step 24
    assert h[0, 5] == 0.0 and 0.4 < h[0, 1] < 0.6
    st = env.get_state()
    assert st["ol_i"][0, 5] == 3 and st["ol_i"][0, 3] == 1 and st["physics_steps"][0] == 28
    # deltas: rows 0 and 1 are copies of the same frame -> zero, but computed (both visible)
    assert np.all(obs[0, VIS + 27:] == 0.0)
    # frame refresh cadence: next capture at physics step 48 = inner iteration 2 of the fourth env step (46, 48, 50, 52)
    depth0 = h[0, 4]
    for _ in range(2):
        obs, *_ = env.step(ZERO)
        assert obs[0, VIS + 4] == depth0
    obs, *_ = env.step(ZERO)
    h = obs[0, VIS:VIS + 27].reshape(3, 9)
    assert h[0, 4] < depth0 and h[0, 4] == h[1, 4] == h[2, 4]
    assert env.get_state()["physics_steps"][0] == 52
    # a fresh frame and the row before it differ: the delta features carry base - prev of cx, cy, area, depth.
    # A one-inner-iteration env emits the observation of every compute_state, so the capture row is visible.
    cfg1, env1, _ = make(fo, inner_per_step=1)
    env1.set_state(aim_at_duck(env1, 120.0))
    seen = []
    for k in range(16):
        o, *_ = env1.step(ZERO)
        seen.append(o[0].copy())
    seen = np.array(seen)
    k_new = [k for k in range(1, 16) if seen[k, VIS + 4] != seen[k - 1, VIS + 4] and seen[k - 1, VIS] == 1.0]
    assert k_new == [13]                                   # captures at physics steps 24 (k=1) and 48 (k=13)
    k = k_new[0]
    d = seen[k, VIS + 27:VIS + 31]
    assert np.allclose(d, (seen[k, VIS + 1:VIS + 5].astype(np.float32) - seen[k - 1, VIS + 1:VIS + 5].astype(np.float32)))
    assert d[3] < 0 and d[2] > 0                           # closing in: depth shrinks, area grows
    assert np.all(seen[k + 1, VIS + 27:VIS + 31] == 0.0)


def test_fixed_camera_geometry(fo):
    """Duck straight ahead on the optical axis of the cockpit camera (tilt -5 deg = 5 deg nose-up) -> image centre."""
    cfg, env, _ = make(fo, inner_per_step=1, cam_interval_substeps=2)
    st = env.get_state()
    t = math.radians(cfg.cam_tilt_deg)
    f = np.array([math.cos(t), 0.0, -math.sin(t)])
    cam = np.array([0.0, 0.0, 50.0]) + np.array(cfg.cam_offset)
    st["pos"][0] = [0.0, 0.0, 50.0]; st["quat"][0] = [0, 0, 0, 1]; st["vel"][0] = [1e-3, 0, 0]; st["omega"][0] = 0
    st["duck"][0] = cam + 40.0 * f - [0, 0, cfg.duck_radius]
    st["act"][0] = 0
    env.set_state(st)
    e = env.envs[0]
    e.vel[0] = 0.0; e.pos[2] = 50.0
    obs, *_ = env.step(ZERO)
    v = obs[0, VIS:VIS + 9]
    # the aircraft drops ~g dt^2 in the two substeps; the projected centre stays within a pixel of (0.5, 0.5)
    assert v[0] == 1.0 and abs(v[1] - 0.5) < 1e-3 and abs(v[2] - 0.5) < 1e-3
    assert abs(v[4] - (40.0 - cfg.duck_radius)) < 0.05
    assert abs(v[3] - math.pi * (cfg.duck_radius / 40.0) ** 2 / 4) < 1e-4
    # duck behind the camera: not visible, steps_since_seen starts counting (per compute_state)
    st = env.get_state()
    st["duck"][0] = cam - 40.0 * f
    env.set_state(st)
    obs, *_ = env.step(ZERO)
    assert obs[0, VIS] == 0.0 and abs(obs[0, VIS + 5] - 1 / 60) < 1e-6
    assert obs[0, VIS + 1] == np.float32(v[1])            # _last_cx keeps the last sighting


def test_lock_counter_strike_and_reward_terms(fo):
    cfg, env, _ = make(fo)
    env.set_state(aim_at_duck(env, 30.0, height=5.0))
    e = env.envs[0]
    total_lock = 0
    for k in range(60):
        obs, r, f, _ = env.step(np.array([[0.0, 0.0, 0.0, 0.5]]))
        if f[0] & FLAG_TERM:
            break
        assert e.lock_steps <= cfg.lock_hold_steps
        total_lock = max(total_lock, e.lock_steps)
    assert total_lock == cfg.lock_hold_steps               # capped at the hold count (:333)
    assert f[0] == (FLAG_TERM | FLAG_COMPLETE | FLAG_STRIKE)
    assert r[0] > cfg.strike_reward                        # +400 on top of the shaping terms of that step
    st_obs = obs                                           # auto-reset: obs is the next episode's first observation
    assert np.allclose(st_obs[0, VIS:VIS + 9], [0, 0.5, 0.5, 0, 0, 1.0, 0, 0, 0])
    # a visible, centred duck far away: per inner iteration 1/max(d,2) + 2 + 5*area + 3*centre + 0.2 (+ approach)
    cfg2, env2, _ = make(fo, inner_per_step=1)
    env2.set_state(aim_at_duck(env2, 150.0, height=20.0))
    for _ in range(2):
        obs, r, f, _ = env2.step(ZERO)                     # second step consumes the frame captured at physics step 24
    e2 = env2.envs[0]
    v = obs[0, VIS:VIS + 9]
    assert v[0] == 1.0 and e2.lock_steps == 1
    dist = np.linalg.norm(obs[0, ATT:ATT + 3])
    dc = math.hypot(v[1] - 0.5, v[2] - 0.5)
    want = -0.1 + 1.0 / max(dist, 2.0) + 2.0 + 5.0 * v[3] + 3.0 * max(0.0, (0.55 - dc) / 0.55) + 0.2
    assert abs(r[0] - want) < 1e-9                         # first sighting: no previous estimate, no approach term
    obs, r, f, _ = env2.step(ZERO)
    dist = np.linalg.norm(obs[0, ATT:ATT + 3])
    want = -0.1 + 1.0 / max(dist, 2.0) + 2.0 + 5.0 * v[3] + 3.0 * max(0.0, (0.55 - dc) / 0.55) + 0.2
    assert abs(r[0] - want) < 1e-9 and e2.lock_steps == 2  # same frame again: approach difference is 0


def test_lock_decays_and_costs_a_penalty_when_the_duck_is_lost(fo):
    cfg, env, _ = make(fo, inner_per_step=1, cam_interval_substeps=2)
    env.set_state(aim_at_duck(env, 150.0, height=20.0))
    e = env.envs[0]
    for _ in range(3):
        env.step(ZERO)
    assert e.lock_steps == 3 and e.has_prev_dist == 1
    st = env.get_state()
    st["duck"][0, :2] = st["pos"][0, :2] - 100.0 * np.array([math.cos(0.3), math.sin(0.3)])     # now behind the aircraft
    env.set_state(st)
    obs, r, f, _ = env.step(ZERO)
    dist = np.linalg.norm(obs[0, ATT:ATT + 3])
    assert obs[0, VIS] == 0.0 and e.lock_steps == 2 and e.has_prev_dist == 0
    assert abs(r[0] - (-0.1 + 1.0 / dist - 0.5)) < 1e-9
    for _ in range(3):
        obs, r, f, _ = env.step(ZERO)
    dist = np.linalg.norm(obs[0, ATT:ATT + 3])
    assert e.lock_steps == 0 and abs(r[0] - (-0.1 + 1.0 / dist)) < 1e-9                      # nothing left to lose


def test_sparse_reward_never_locks_and_never_strikes(fo):
    cfg, env, _ = make(fo, sparse_reward=1)
    env.set_state(aim_at_duck(env, 60.0, height=12.0))
    e = env.envs[0]
    for _ in range(25):
        obs, r, f, _ = env.step(ZERO)
        if f[0] & (FLAG_TERM | FLAG_TRUNC):
            break
        assert r[0] == -0.1 and e.lock_steps == 0          # the lock counter only advances inside the dense branch (:300-356)
    assert not (f[0] & FLAG_STRIKE)


def test_crash_returns_exactly_minus_100_and_truncation_at_1800_steps(fo):
    cfg, env, _ = make(fo)
    st = env.get_state()
    st["pos"][0] = [0.0, 0.0, 0.05]
    st["vel"][0] = [20.0, 0.0, -5.0]
    env.set_state(st)
    obs, r, f, term = env.step(ZERO)
    assert f[0] == (FLAG_TERM | FLAG_COLLISION) and r[0] == -100.0
    st = env.get_state()
    st["pos"][0] = [199.5, 0.0, 50.0]
    env.set_state(st)
    obs, r, f, _ = env.step(ZERO)
    assert f[0] == (FLAG_TERM | FLAG_OOB) and r[0] == -100.0
    st = env.get_state()
    st["step_count"][0] = 1800
    st["pos"][0] = [0.0, 0.0, 150.0]
    env.set_state(st)
    obs, r, f, _ = env.step(ZERO)
    assert f[0] == 0
    obs, r, f, _ = env.step(ZERO)                           # step_count 1801 > max_steps: truncation, first inner iteration
    assert f[0] == FLAG_TRUNC


def test_obstacles_keep_clear_of_the_duck_and_the_start(fo):
    cfg, env, _ = make(fo, n=128, num_obstacles=20)
    st = env.get_state()
    n_obst = st["ol_i"][:, 8]
    assert n_obst.max() <= 20 and n_obst.min() >= 12
    for i in range(128):
        o = st["obst"][i, : n_obst[i]]
        assert np.all(np.hypot(o[:, 0] - st["duck"][i, 0], o[:, 1] - st["duck"][i, 1]) >= 10.0)     # :536-539
        assert np.all(o[:, 0] ** 2 + o[:, 1] ** 2 >= 100.0)                                         # :542-543
        assert np.all((o[:, 2] >= 10) & (o[:, 2] <= 30)) and np.all(np.abs(o[:, :2]) <= 100)
    # the obstacle penalty always uses the halved scale (:403): band depth 4 m of safe 10 m -> 0.5 * 0.6
    e = env.envs[0]
    before = e.reward
    cfg1, env1, _ = make(fo, inner_per_step=1, num_obstacles=1, sparse_reward=1)
    st = env1.get_state()
    st["pos"][0] = [0.0, 0.0, 20.0]; st["quat"][0] = [0, 0, 0, 1]; st["vel"][0] = [20.0, 0, 0]; st["omega"][0] = 0
    st["obst"][0] = 0; st["obst"][0, 0] = [40.0, 0.0, 60.0]; st["ol_i"][0, 8] = 1
    st["duck"][0] = [-150.0, 0.0, 0.05]
    env1.set_state(st)
    for _ in range(3):
        obs, r, f, _ = env1.step(ZERO)
    dmin = min(d for d in obs[0, VIS + 6:VIS + 9] if d > 0)
    assert dmin < cfg1.obst_safe or r[0] == -0.1
    if dmin < cfg1.obst_safe:
        assert abs(r[0] - (-0.1 - min(0.5 * (10.0 - dmin) / 10.0, 5.0))) < 1e-6
