"""Committed regression fixtures (tests/golden/oracle_traj_v1.npz, written by scripts/make_golden.py).

The fixtures were produced by this repo's fp64 oracle, NOT by PyFlyt/pybullet (absent offline; parity stays
unpinned).  CPU: the oracle still reproduces them.  GPU: the CUDA path, through the C ABI,
reproduces every stored step from the stored pre-step state within the north-star tolerance (1e-4 relative per
step, flags exact).
"""
import os

import numpy as np
import pytest

import pyflyt_drone_b200 as fw

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_traj_v1.npz")
CASES = {"sparse_euler": dict(noise_ratio=0.0),
         "dense_quat": dict(noise_ratio=0.0, sparse_reward=0, angle_repr=1, goal_reach=30.0)}
N, SEED, ID0 = 24, 11, 5
RTOL = 1e-4


def _pre(g, name, t):
    keys = [k.split("/pre/")[1] for k in g.files if k.startswith(f"{name}/pre/")]
    return {k: g[f"{name}/pre/{k}"][t] for k in keys}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(oracle_mod, name):
    g = np.load(GOLD)
    cfg = fw.waypoints_v3(**CASES[name])
    orc = oracle_mod.OracleVecEnv(cfg.as_dict(), N, seed=SEED, env_id0=ID0)
    assert np.abs(orc.reset() - g[f"{name}/obs0"]).max() < 1e-9
    acts = g[f"{name}/actions"]
    for t in range(acts.shape[0]):
        o, r, f, _ = orc.step(acts[t].astype(np.float64))
        assert np.array_equal(f, g[f"{name}/flags"][t]), t
        assert np.abs(o - g[f"{name}/obs"][t]).max() < 1e-9, t
        assert np.abs(r - g[f"{name}/rew"][t]).max() < 1e-9, t
    assert (g[f"{name}/flags"] != 0).any(), "fixture must contain terminations"


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_golden_steps(name):
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    g = np.load(GOLD)
    cfg = fw.waypoints_v3(**CASES[name])
    env = FixedwingVecEnv(N, config=cfg, seed=SEED, env_id0=ID0)
    o0 = env.reset()
    euler = name == "sparse_euler"
    a = 12 if euler else 13
    bounds = ([0, 3, 6, 9, 12] if euler else [0, 3, 7, 10, 13]) + [a + 4, a + 10, a + 13, a + 16]

    def err(got, ref):
        """worst norm-wise relative error over the observation groups (ang vel, attitude, lin vel, pos, action,
        aux, two target deltas), reference norm floored at 1 -- the measure tests/test_parity_gpu.py uses"""
        d = got - ref
        if euler:
            d[:, 3:6] = (d[:, 3:6] + np.pi) % (2 * np.pi) - np.pi      # euler angles wrap at +-pi
        worst = 0.0
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            scale = np.maximum(np.abs(ref[:, lo:hi]).max(axis=1), 1.0)
            worst = max(worst, float((np.abs(d[:, lo:hi]).max(axis=1) / scale).max()))
        return worst

    assert err(o0.astype(np.float64), g[f"{name}/obs0"]) < RTOL
    acts = g[f"{name}/actions"]
    worst = 0.0
    for t in range(acts.shape[0]):
        env.set_state(_pre(g, name, t))
        o, r, f, _ = env.step_arrays(acts[t])
        assert np.array_equal(f.astype(np.uint8), g[f"{name}/flags"][t]), t
        worst = max(worst, err(o.astype(np.float64), g[f"{name}/obs"][t]))
        rc = g[f"{name}/rew"][t]
        assert np.abs(r - rc).max() <= 1e-4 * max(1.0, np.abs(rc).max()), t
    assert worst < RTOL, worst
    env.close()


# ---------------------------------------------------------------- task heads (oracle_traj_v2_heads.npz)
GOLD2 = os.path.join(os.path.dirname(__file__), "golden", "oracle_traj_v2_heads.npz")
HEADS = {"lowlevel": dict(bounds=[0, 3, 6, 9, 12, 18, 21], euler=True),
         "objlock_duck": dict(bounds=[0, 3, 6, 9, 12, 16, 22, 25], euler=True)}


def _pre2(g, name, t):
    keys = [k.split("/pre/")[1] for k in g.files if k.startswith(f"{name}/pre/")]
    return {k: g[f"{name}/pre/{k}"][t] for k in keys}


@pytest.mark.parametrize("name", sorted(HEADS))
def test_oracle_reproduces_golden_task_heads(oracle_mod, name):
    g = np.load(GOLD2)
    cfg = fw.make_config(name, noise_ratio=0.0)
    orc = oracle_mod.OracleVecEnv(cfg.as_dict(), N, seed=SEED, env_id0=ID0)
    assert np.abs(orc.reset() - g[f"{name}/obs0"]).max() < 1e-9
    orc.set_state(_pre2(g, name, 0))                    # the generator edits the reset state (scripts/make_golden.py)
    acts = g[f"{name}/actions"]
    for t in range(acts.shape[0]):
        o, r, f, _ = orc.step(acts[t].astype(np.float64))
        assert np.array_equal(f, g[f"{name}/flags"][t]), t
        assert np.abs(o - g[f"{name}/obs"][t]).max() < 1e-9, t
        assert np.abs(r - g[f"{name}/rew"][t]).max() < 1e-9, t
    flags = np.unique(g[f"{name}/flags"])
    assert len(flags) > 1, "fixture must contain terminations"
    if name == "objlock_duck":
        assert 49 in flags                              # TERM | COMPLETE | STRIKE


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(HEADS))
def test_cuda_path_reproduces_golden_task_heads(name):
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    g = np.load(GOLD2)
    cfg = fw.make_config(name, noise_ratio=0.0)
    env = FixedwingVecEnv(N, config=cfg, seed=SEED, env_id0=ID0)
    o0 = env.reset()
    bounds = HEADS[name]["bounds"]

    def err(got, ref):
        d = got - ref
        d[:, 3:6] = (d[:, 3:6] + np.pi) % (2 * np.pi) - np.pi
        worst = 0.0
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            scale = np.maximum(np.abs(ref[:, lo:hi]).max(axis=1), 1.0)
            worst = max(worst, float((np.abs(d[:, lo:hi]).max(axis=1) / scale).max()))
        return worst

    assert err(o0.astype(np.float64), g[f"{name}/obs0"]) < RTOL
    acts = g[f"{name}/actions"]
    worst, vis_bad, vis_n = 0.0, 0, 0
    for t in range(acts.shape[0]):
        env.set_state(_pre2(g, name, t))
        o, r, f, _ = env.step_arrays(acts[t])
        assert np.array_equal(f.astype(np.uint8), g[f"{name}/flags"][t]), t
        ref = g[f"{name}/obs"][t]
        worst = max(worst, err(o.astype(np.float64), ref))
        rc = g[f"{name}/rew"][t]
        assert (np.abs(r - rc) <= 2e-4 * np.maximum(1.0, np.abs(rc))).all(), t
        if name == "objlock_duck":                      # vision history + deltas: equal to rounding
            bad = np.abs(o[:, 25:] - ref[:, 25:]) > 2e-4 * np.maximum(np.abs(ref[:, 25:]), 1.0)
            vis_bad += int(bad.sum()); vis_n += bad.size
    assert worst < RTOL, worst
    assert vis_bad <= 0.01 * max(vis_n, 1)
    env.close()
