"""CPU tests of the Waypoint+ObjLock task on the fp64 oracle: state machine and reward terms of
/root/reference/envs/fixedwing_waypoint_objlock_env.py:197-380 and the analytic camera that stands in for
PyBullet's rasteriser (documented approximation)."""
import ctypes as C
import math

import numpy as np
import pytest

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.config import FLAG_COLLISION, FLAG_COMPLETE, FLAG_STRIKE, FLAG_TERM


@pytest.fixture(scope="module")
def fo(oracle_mod):
    return oracle_mod


def make(fo, n=1, **kw):
    cfg = fw.waypoint_objlock(noise_ratio=0.0, wind={"enabled": False}, **kw)
    env = fo.OracleVecEnv(cfg.as_dict(), n, seed=2)
    obs = env.reset()
    return cfg, env, obs


ZERO = np.zeros((1, 4))


def level_state(env, pos, yaw=0.0, speed=20.0):
    st = env.get_state()
    st["pos"][0] = pos
    st["quat"][0] = [0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2)]
    st["vel"][0] = [speed * math.cos(yaw), speed * math.sin(yaw), 0.0]
    st["omega"][0] = 0.0
    return st


def test_reset_spawns_duck_under_last_waypoint_and_prunes_near_obstacles(fo):
    cfg, env, obs = make(fo, n=64)
    st = env.get_state()
    assert np.allclose(st["duck"][:, :2], st["targets"][:, -1, :2]) and np.allclose(st["duck"][:, 2], 0.05)
    n_obst = st["ol_i"][:, 8]
    assert n_obst.max() <= 20 and n_obst.min() >= 10
    for i in range(64):
        o = st["obst"][i, : n_obst[i]]
        assert np.all(o[:, 0] ** 2 + o[:, 1] ** 2 >= 100.0)            # x*x + y*y < 100 -> skipped (:472)
        assert np.all((o[:, 2] >= 10) & (o[:, 2] <= 30)) and np.all(np.abs(o[:, :2]) <= 50)
    # the warm-up captured one frame at physics step 12 and compute_state consumed it
    assert np.all(st["ol_i"][:, 3] == 1)
    assert obs.shape == (64, 28)


def test_duck_row_rides_after_the_remaining_waypoints(fo):
    cfg, env, _ = make(fo, num_targets=1)
    e = env.envs[0]
    st = level_state(env, [0.0, 0.0, 30.0])
    st["targets"][0, 0] = [60.0, 0.0, 30.0]
    st["duck"][0] = [60.0, 0.0, 0.05]
    st["obst"][0] = 0; st["ol_i"][0, 8] = 0
    env.set_state(st)
    obs, rew, flags, _ = env.step(ZERO)
    # one waypoint left: rows = (waypoint, duck)
    assert obs[0, 22] > 50 and abs(obs[0, 24]) < 3
    assert obs[0, 25] > 50 and obs[0, 27] < -25
    assert rew[0] > -0.1          # dense shaping: progress + 1/dist


def test_duck_visibility_geometry(fo):
    cfg, env, _ = make(fo, num_targets=1)
    e = env.envs[0]
    st = level_state(env, [0.0, 0.0, 20.0])
    st["targets"][0, 0] = [90.0, 0.0, 20.0]
    st["duck"][0] = [40.0, 0.0, 0.05]          # ahead and below: inside the 90-degree frustum of the chase camera
    st["obst"][0] = 0; st["ol_i"][0, 8] = 0
    st["physics_steps"][0] = 22                 # next Aviary.step ends on physics step 24: a capture step
    env.set_state(st)
    env.step(ZERO)
    assert e.frame_visible == 1
    assert 0.4 < e.frame_cx < 0.6 and e.frame_cy > 0.5       # centred, below the horizon line
    assert 30 < e.frame_depth < 50 and 0 < e.frame_area < 0.01
    assert e.vision[0] == 1.0 and e.steps_since_seen == 0
    # an obstacle between camera and duck hides it
    st = level_state(env, [0.0, 0.0, 20.0])
    st["physics_steps"][0] = 22
    st["obst"][0, 0] = [20.0, 0.0, 30.0]; st["ol_i"][0, 8] = 1
    env.set_state(st)
    env.step(ZERO)
    assert e.frame_visible == 0 and e.steps_since_seen >= 1


def test_obstacle_bands_and_penalty(fo):
    cfg, env, _ = make(fo, num_targets=1)
    e = env.envs[0]
    st = level_state(env, [0.0, 0.0, 15.0])
    st["targets"][0, 0] = [90.0, 0.0, 15.0]
    st["duck"][0] = [-80.0, 0.0, 0.05]
    st["obst"][0] = 0
    st["obst"][0, 0] = [9.0, 0.0, 30.0]; st["ol_i"][0, 8] = 1     # cylinder dead ahead, ~6 m in front after the step
    st["physics_steps"][0] = 22
    env.set_state(st)
    obs, rew, flags, _ = env.step(ZERO)
    assert e.frame_dc < e.frame_dl and e.frame_dc < e.frame_dr
    assert e.frame_dc < 20.0
    # penalty = scale * (d_safe - d)/d_safe capped at max_pen, applied every inner iteration
    d = min(x for x in (e.vision[6], e.vision[7], e.vision[8]) if x > 0)
    if d < cfg.obst_safe:
        assert rew[0] < -0.1


def test_obstacle_collision_terminates_with_minus_100(fo):
    cfg, env, _ = make(fo, num_targets=1)
    st = level_state(env, [0.0, 0.0, 15.0])
    st["obst"][0] = 0
    st["obst"][0, 0] = [2.5, 0.0, 30.0]; st["ol_i"][0, 8] = 1
    env.set_state(st)
    _, rew, flags, _ = env.step(ZERO)
    assert flags[0] == FLAG_TERM | FLAG_COLLISION and rew[0] == pytest.approx(-100.0)


def test_phase_switch_lock_and_strike(fo):
    cfg, env, _ = make(fo, num_targets=1, goal_reach=8.0)
    e = env.envs[0]
    st = level_state(env, [0.0, 0.0, 40.0])
    st["targets"][0, 0] = [3.0, 0.0, 40.0]          # reached in the first inner iteration
    st["duck"][0] = [50.0, 0.0, 0.05]
    st["obst"][0] = 0; st["ol_i"][0, 8] = 0
    env.set_state(st)
    _, rew, flags, _ = env.step(ZERO)
    assert e.n_remaining == 0 and flags[0] == 0      # last waypoint does NOT terminate or truncate (:297-300)
    assert e.post_waypoints == 1
    # dive at the duck: pitch the nose down so the duck sits near the image centre, keep stepping
    total, struck = 0.0, False
    for k in range(120):
        act = np.array([[0.0, 0.35, 0.0, 0.0]])
        _, rew, flags, _ = env.step(act)
        total += rew[0]
        if flags[0] & FLAG_STRIKE:
            struck = True
            assert flags[0] & FLAG_TERM and flags[0] & FLAG_COMPLETE and rew[0] > 190
            break
        if flags[0] & FLAG_TERM:
            break
    assert e.episode == 1          # the episode ended one way or another and auto-reset
    # phase bookkeeping is monotone while alive: duck_phase implies >= 2 consecutive sightings
    cfg2, env2, _ = make(fo, num_targets=1)
    e2 = env2.envs[0]
    st2 = env2.get_state()
    st2["ol_i"][0] = [0, 0, 1, 1, 1, 1, 0, 0, 0]     # post waypoints, valid visible frame, one sighting so far
    st2["ol_f"][0] = [0.5, 0.5, 0.01, 30.0, 0.5, 0.55, 0.01, 30.0, 0, 0, 0, 0]
    st2["target_idx"][0] = 1
    st2["targets"] = st2["targets"]
    st2 = {**st2, **level_state(env2, [0.0, 0.0, 40.0])}
    st2["target_idx"][0] = 1
    st2["duck"][0] = [40.0, 0.0, 0.05]
    st2["obst"][0] = 0; st2["ol_i"][0, 8] = 0
    env2.set_state(st2)
    _, rew, flags, _ = env2.step(ZERO)
    assert e2.duck_phase == 1 and e2.lock_steps >= 1
    assert rew[0] > -0.1           # dense 1/max(depth,2) + lock reward




def test_oracle_debug_frame_geometry(fo):
    """fwo_render: a duck placed straight ahead of a level aircraft projects to the image centre column,
    its silhouette area agrees with the pin-hole estimate pi (R / z)^2 / 4 the vision features use, the ground fills the
    lower half and depth-buffer values grow towards the horizon."""
    cfg = fw.waypoint_objlock(num_obstacles=0)
    o = fo.OracleVecEnv(cfg.as_dict(), 1, seed=1)
    o.reset()
    st = o.get_state()
    st["pos"][0] = [0.0, 0.0, 20.0]; st["quat"][0] = [0.0, 0.0, 0.0, 1.0]; st["vel"][0] = [20.0, 0.0, 0.0]; st["omega"][0] = 0.0
    st["duck"][0] = [60.0, 0.0, 0.0]
    o.set_state(st)
    W = 256
    f = o.render(0, W, W)
    seg, depth = f["seg"], f["depth"]
    ys, xs = np.nonzero(seg == 1)
    assert len(xs) > 20
    assert abs(xs.mean() - (W - 1) / 2) < 1.0                       # centred left-right
    cam = np.array([0.0, 0.0, 20.0]) + np.array(cfg.cam_offset)
    ctr = np.array([60.0, 0.0, cfg.duck_radius])
    fwd = -np.array(cfg.cam_offset) / np.linalg.norm(cfg.cam_offset)
    right = np.cross(fwd, [0.0, 0.0, 1.0]); right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    zc, yc = float((ctr - cam) @ fwd), float((ctr - cam) @ upv)
    assert ys.mean() / W == pytest.approx(0.5 - 0.5 * yc / zc, abs=0.01)      # the row the vision features report as cy
    area = np.pi * (cfg.duck_radius / zc) ** 2 / 4.0
    assert len(xs) / (W * W) == pytest.approx(area, rel=0.15)
    # ground (or a waypoint sphere in front of it) along the bottom row, sky (or a waypoint sphere) along the top
    assert ((seg[-1] == 0) | (seg[-1] >= 64)).all() and ((seg[0] == -1) | (seg[0] >= 64)).all() and (seg[-1] == 0).any()
    col = depth[:, 3]
    g = col[seg[:, 3] == 0]
    assert (np.diff(g) <= 1e-12).all() and g.max() <= 1.0           # rows further down are nearer
