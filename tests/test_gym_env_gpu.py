"""GPU tests of the single-env gymnasium-style view (seam B2) and the FlattenWaypointEnv wrapper."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_waypoints_env_matches_oracle_through_the_gym_surface(oracle_mod):
    import pyflyt_drone_b200 as fw
    from pyflyt_drone_b200.gym_env import FlattenWaypointEnv, make
    env = make("PyFlyt/Fixedwing-Waypoints-v3", sparse_reward=True, num_targets=8, goal_reach_distance=4,
               angle_representation="euler", flight_dome_size=100.0, max_duration_seconds=120.0, agent_hz=30, seed=7)
    assert env.observation_space["attitude"].shape == (22,)
    flat = FlattenWaypointEnv(env, context_length=2)
    assert flat.observation_space.shape == (28,)
    obs, info = flat.reset(seed=7)
    cfg = fw.waypoints_v3(context_len=4)
    orc = oracle_mod.OracleVecEnv(cfg.as_dict(), 1, seed=7)
    oc = orc.reset()
    assert obs.shape == (28,) and np.abs(obs - np.concatenate([oc[0, :22], oc[0, 22:28]])).max() < 1e-3
    assert len(env.unwrapped.waypoints.targets) == 8 and info["num_targets_reached"] == 0
    d = env.state
    assert d["target_deltas"].shape == (8, 3)
    rng = np.random.default_rng(0)
    for _ in range(400):
        a = rng.uniform(-1, 1, 4)
        o, r, term, trunc, info = flat.step(a)
        oo, ro, fo, to = orc.step(a.reshape(1, 4))
        assert (term, trunc) == (bool(fo[0] & 1), bool(fo[0] & 2))
        assert r == pytest.approx(ro[0], abs=1e-4)
        if term or trunc:
            assert np.abs(o[:22] - to[0, :22]).max() < 2e-2      # terminal observation, not the auto-reset one
            assert info["collision"] or info["out_of_bounds"] or info["env_complete"] or trunc
            with pytest.raises(RuntimeError):
                flat.step(a)
            break
    else:
        pytest.fail("episode did not end under random actions")
    flat.close()


def test_gym_env_validation_errors_match_reference():
    from pyflyt_drone_b200.gym_env import FixedwingWaypointsEnv
    with pytest.raises(ValueError, match="agent_hz"):
        FixedwingWaypointsEnv(agent_hz=50)
    with pytest.raises(ValueError, match="angle_representation"):
        FixedwingWaypointsEnv(angle_representation="matrix")
    env = FixedwingWaypointsEnv(num_targets=2)
    with pytest.raises(NotImplementedError, match="wind"):
        env.env.register_wind_field_function(lambda t, p: p * 0)
    with pytest.raises(RuntimeError):
        env.step(np.zeros(4))
    env.set_wind_config({"enabled": True, "mode": "constant", "wind_enu_mps": [1.0, 0.0, 0.0]})
    obs, _ = env.reset(seed=1)
    assert obs["attitude"].shape == (23,)
    env.close()


def test_objlock_env_dict_observation():
    from pyflyt_drone_b200.gym_env import FixedwingWaypointsEnv
    env = FixedwingWaypointsEnv(task="objlock", num_targets=3, goal_reach_distance=8.0, angle_representation="euler",
                                num_obstacles=20, obstacle_safe_distance_m=5.0, duck_strike_distance_m=8.0)
    obs, info = env.reset(seed=3)
    assert obs["target_deltas"].shape == (4, 3)            # three waypoints + the duck row
    assert obs["duck_vision"].shape == (9,) and obs["duck_vision"].dtype == np.float32
    assert "duck_strike" in info
    o, r, term, trunc, info = env.step(np.zeros(4))
    assert np.isfinite(r)
    env.close()


def test_lowlevel_gym_view_matches_the_reference_interface():
    """envs/fixedwing_envs/fixedwing_lowlevel_env.py: Box(21) obs, Box(6) action, info["target"], 2000-step truncation."""
    from pyflyt_drone_b200.gym_env import FixedwingLowLevelEnv
    env = FixedwingLowLevelEnv(seed=3)
    assert env.observation_space.shape == (21,) and env.action_space.shape == (6,)
    obs, info = env.reset(seed=3)
    assert obs.shape == (21,) and np.allclose(obs[9:12], [0, 0, 10.0]) and np.allclose(info["target"], obs[18:21])
    a = np.array([0.1, -0.1, 0.2, 0.0, 0.0, 0.7])
    total, steps = 0.0, 0
    while True:
        obs, r, term, trunc, info = env.step(a)
        total += r; steps += 1
        assert np.allclose(obs[12:18], a, atol=1e-6) and np.allclose(obs[18:21], info["target"])
        if term or trunc:
            break
        assert 1.0 <= obs[11] <= 100.0
    assert steps <= 2000 and (trunc == (steps == 2000) or term)
    if term:
        assert obs[11] < 1.0 or obs[11] > 100.0
    with pytest.raises(RuntimeError):
        env.step(a)
    env.close()
