"""GPU parity of the duck-only lock/strike task (FixedwingObjLockEnv + FlattenObjLockEnv) against the
fp64 oracle: reset, single steps from injected state (random flight and a lock/strike approach), the vision history,
a free-running random rollout, and the single-env gym view.  Tolerances as in test_objlock_gpu.py."""
import math

import numpy as np
import pytest

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.config import FLAG_STRIKE, FLAG_TERM, FLAG_TRUNC

pytestmark = pytest.mark.gpu
RTOL = 1e-4
ATT = 22
GROUPS = {"ang_vel": slice(0, 3), "ang_pos": slice(3, 6), "lin_vel": slice(6, 9), "lin_pos": slice(9, 12),
          "action": slice(12, 16), "aux": slice(16, 22), "target_vector": slice(22, 25)}
HIST = slice(25, 56)


@pytest.fixture(scope="module")
def fo(oracle_mod):
    return oracle_mod


def row_err(got, ref):
    """norm-wise relative error per row, worst group"""
    worst = np.zeros(len(ref))
    for k, sl in GROUPS.items():
        d = np.abs(got[:, sl] - ref[:, sl]).max(axis=1)
        worst = np.maximum(worst, d / np.maximum(np.abs(ref[:, sl]).max(axis=1), 1.0))
    return worst


def group_err(got, ref):
    out = {}
    for k, sl in GROUPS.items():
        d = np.abs(got[:, sl] - ref[:, sl]).max(axis=1)
        out[k] = float((d / np.maximum(np.abs(ref[:, sl]).max(axis=1), 1.0)).max()) if len(d) else 0.0
    return out


def angle_safe(og, oc):
    og = og.copy()
    d = og[:, 3:6] - oc[:, 3:6]
    og[:, 3:6] = oc[:, 3:6] + (d + np.pi) % (2 * np.pi) - np.pi
    return og


def pair(fo, n, cfg, seed=11):
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    return FixedwingVecEnv(n, config=cfg, seed=seed), fo.OracleVecEnv(cfg.as_dict(), n, seed=seed)


def hist_ok(og, oc):
    """vision part of the observation: equal to rounding (the features are float32 in the reference too)"""
    return np.abs(og[:, HIST] - oc[:, HIST]) <= 2e-4 * np.maximum(np.abs(oc[:, HIST]), 1.0)


@pytest.mark.parametrize("num_obstacles", [0, 12])
def test_duck_reset_parity(fo, num_obstacles):
    cfg = fw.make_config("objlock_duck", num_obstacles=num_obstacles)
    env, orc = pair(fo, 300, cfg)
    og, oc = env.reset(), orc.reset()
    assert og.shape == (300, 56)
    assert max(group_err(og, oc).values()) < RTOL
    assert hist_ok(og, oc).all()
    sg, sc = env.get_state(), orc.get_state()
    assert np.array_equal(sg["ol_i"], sc["ol_i"])
    n = sc["ol_i"][:, 8]
    for i in range(300):
        assert np.abs(sg["obst"][i, : n[i]] - sc["obst"][i, : n[i]]).max(initial=0.0) < 1e-4
    assert np.abs(sg["duck"] - sc["duck"]).max() < 1e-4
    assert np.abs(sg["wind"] - sc["wind"]).max() < 1e-5
    assert np.abs(sg["vis_hist"] - sc["vis_hist"]).max() < 1e-6
    env.close()


@pytest.mark.parametrize("scenario", ["random_flight", "approach"])
def test_duck_single_step_parity(fo, scenario):
    cfg = fw.make_config("objlock_duck", noise_ratio=0.0, num_obstacles=8 if scenario == "random_flight" else 0)
    N = 512
    env, orc = pair(fo, N, cfg)
    env.reset(); orc.reset()
    rng = np.random.default_rng(5)
    if scenario == "approach":
        # low, 25-60 m short of the duck and heading at it: frames with the duck in view, locks and strikes
        st = orc.get_state()
        ang = rng.uniform(-np.pi, np.pi, N)
        dist = rng.uniform(25, 60, N)
        st["pos"][:, 0] = st["duck"][:, 0] - dist * np.cos(ang)
        st["pos"][:, 1] = st["duck"][:, 1] - dist * np.sin(ang)
        st["pos"][:, 2] = rng.uniform(4, 9, N)
        st["quat"][:] = np.stack([0 * ang, 0 * ang, np.sin(ang / 2), np.cos(ang / 2)], 1)
        st["vel"][:] = np.stack([20 * np.cos(ang), 20 * np.sin(ang), 0 * ang], 1)
        st["omega"][:] = 0
        orc.set_state(st)
    worst, mism, events = {}, 0, dict(done=0, lock=0, strike=0, visible=0, deltas=0)
    hist_bad, hist_n, rew_bad, rew_n, rows = 0, 0, 0, 0, []
    for k in range(60):
        env.set_state(orc.get_state())
        a = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
        if scenario == "approach":
            a[:, :3] *= 0.1
            a[:, 3] = 0.0
        og, rg, fg, tg = env.step_arrays(a)
        og, rg, fg, tg = og.copy(), rg.copy(), fg.copy().astype(np.int32), tg.copy()
        oc, rc, fc, tc = orc.step(a.astype(np.float64))
        sg, sc = env.get_state(), orc.get_state()
        mism += int((fg != fc).sum()) + int((sg["ol_i"] != sc["ol_i"]).sum())
        done = (fc & (FLAG_TERM | FLAG_TRUNC)) != 0
        ok = fg == fc
        e = group_err(angle_safe(og, oc)[ok], oc[ok])
        rows.append(row_err(angle_safe(og, oc)[ok], oc[ok]))
        for g, v in e.items():
            worst[g] = max(worst.get(g, 0.0), v)
        # terminal observations of finished episodes carry the pre-reset history
        if done.any():
            both = ok & done
            e2 = group_err(angle_safe(tg, tc)[both], tc[both])
            rows.append(row_err(angle_safe(tg, tc)[both], tc[both]))
            for g, v in e2.items():
                worst[g] = max(worst.get(g, 0.0), v)
            hist_bad += int((~hist_ok(tg[both], tc[both])).sum()); hist_n += int(both.sum()) * 31
        rew_bad += int((np.abs(rg - rc)[ok] > 2e-4 * np.maximum(1.0, np.abs(rc[ok]))).sum()); rew_n += int(ok.sum())
        # vision entries: continuous in the state except where a pixel-column ray grazes a silhouette edge
        hb = ~hist_ok(og[ok], oc[ok])
        hist_bad += int(hb.sum()); hist_n += hb.size
        events["done"] += int(done.sum()); events["lock"] += int((sc["ol_i"][:, 6] > 0).sum())
        events["strike"] += int(((fc & FLAG_STRIKE) != 0).sum()); events["visible"] += int(sc["ol_i"][:, 4].sum())
        events["deltas"] += int((np.abs(oc[:, 52:56]).max(axis=1) > 0).sum())
    print(f"\n[duck/{scenario}] worst group rel err " + ", ".join(f"{g}={v:.1e}" for g, v in worst.items())
          + f"; flag/state-machine mismatches {mism}; vision entries off by more than rounding {hist_bad}/{hist_n}; "
            f"rewards off by more than 2e-4: {rew_bad}/{rew_n}; events {events}")
    assert hist_bad <= 0.01 * hist_n
    assert rew_bad <= 0.002 * rew_n
    # The +-10 m/s wind of this preset keeps the tail surfaces near their +-9 degree stall angles, where the
    # Khan-Nahon model is discontinuous: a step whose angle of attack lands within rounding of the boundary takes
    # the other branch in fp32 (measured with scripts/duck_parity_probe.py: 4 of 30,720 steps, 2e-4..8e-4, none with
    # the wind off).  Everything else must meet the single-step tolerance.
    rows = np.concatenate(rows)
    flips = int((rows >= RTOL).sum())
    print(f"steps past the {RTOL:g} tolerance (stall-boundary branch flips): {flips}/{rows.size}, worst {rows.max():.1e}")
    assert flips <= 5e-4 * rows.size and rows.max() < 5e-3
    assert mism <= 2
    if scenario == "approach":
        # (with four inner iterations per step and a 20-substep warm-up a fresh frame always arrives in the second inner
        # iteration, so the delta features are back to zero by the time the step's observation is emitted: the deltas
        # are exercised by the one-inner-iteration test below)
        assert events["visible"] > 0 and events["lock"] > 0 and events["strike"] > 0
    env.close()


def test_duck_history_persists_across_launches(fo):
    """Free-running (no state re-injection): the history planes written by one launch are what the next one shifts."""
    cfg = fw.make_config("objlock_duck", noise_ratio=0.0, wind={"enabled": False}, inner_per_step=1)
    N = 128
    env, orc = pair(fo, N, cfg, seed=4)
    env.reset(); orc.reset()
    st = orc.get_state()
    rng = np.random.default_rng(9)
    ang = rng.uniform(-np.pi, np.pi, N)
    st["pos"][:, 0] = st["duck"][:, 0] - 80 * np.cos(ang); st["pos"][:, 1] = st["duck"][:, 1] - 80 * np.sin(ang)
    st["pos"][:, 2] = 15.0
    st["quat"][:] = np.stack([0 * ang, 0 * ang, np.sin(ang / 2), np.cos(ang / 2)], 1)
    st["vel"][:] = np.stack([20 * np.cos(ang), 20 * np.sin(ang), 0 * ang], 1); st["omega"][:] = 0
    orc.set_state(st); env.set_state(orc.get_state())
    a = np.zeros((N, 4), np.float32); a[:, 3] = 0.2
    nz_delta = 0
    for k in range(30):
        og, rg, fg, _ = env.step_arrays(a)
        oc, rc, fc, _ = orc.step(a.astype(np.float64))
        assert np.array_equal(fg.astype(np.int32), fc)
        assert (np.abs(og[:, HIST] - oc[:, HIST]) <= 1e-3 * np.maximum(np.abs(oc[:, HIST]), 1.0)).mean() > 0.995
        assert np.abs(rg - rc).max() < 5e-3
        nz = np.abs(oc[:, 52:56]).max(axis=1) > 0
        nz_delta += int(nz.sum())
        assert np.array_equal(nz, np.abs(og[:, 52:56]).max(axis=1) > 0)
    assert nz_delta >= N                                   # the step after each fresh frame carries non-zero deltas
    sg, sc = env.get_state(), orc.get_state()
    assert np.array_equal(sg["ol_i"], sc["ol_i"]) and (sc["ol_i"][:, 5] == 3).all()
    env.close()


def test_duck_random_rollout_matches_oracle_counters(fo):
    cfg = fw.make_config("objlock_duck", num_obstacles=6)
    N = 256
    env, orc = pair(fo, N, cfg, seed=21)
    env.reset(); orc.reset()
    env.step_random(16)
    orc.rollout_random(16)
    sg, sc = env.get_state(), orc.get_state()
    agree = float((sg["episode"] == sc["episode"]).mean())
    print(f"\n[duck] 16-step random rollout: episode counters agree on {agree * 100:.1f}% of envs")
    assert agree > 0.97
    ok = sg["episode"] == sc["episode"]
    assert np.abs(sg["pos"][ok] - sc["pos"][ok]).max() < 5e-3
    assert (sg["ol_i"][ok] == sc["ol_i"][ok]).mean() > 0.99
    env.close()


def test_duck_gym_view_and_vec_infos():
    from pyflyt_drone_b200.gym_env import FixedwingObjLockEnv, FlattenObjLockEnv
    env = FlattenObjLockEnv(FixedwingObjLockEnv(angle_representation="euler", flight_dome_size=200.0,
                                                max_duration_seconds=60.0, num_obstacles=0, duck_lock_hold_steps=5,
                                                duck_strike_distance_m=10.0, duck_strike_reward=400.0,
                                                duck_lock_step_reward=0.2, duck_approach_reward_scale=0.1,
                                                duck_global_scaling=60.0, duck_camera_capture_interval_steps=12,
                                                render_resolution=(480, 480), seed=3))
    assert env.observation_space.shape == (56,)
    obs, info = env.reset(seed=5)
    assert obs.shape == (56,) and obs.dtype == np.float32 and info["duck_strike"] is False and info["is_success"] is False
    d = env.env.state
    assert set(d) == {"attitude", "target_vector", "duck_vision"} and d["duck_vision"].shape == (31,)
    assert abs(np.linalg.norm(d["target_vector"]) - np.linalg.norm(env.unwrapped.duck_pos - obs[9:12])) < 1e-2
    for _ in range(5):
        obs, r, term, trunc, info = env.step(np.zeros(4))
        assert obs.shape == (56,) and isinstance(r, float)
    env.close()
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    v = FixedwingVecEnv(8, preset="objlock_duck", seed=1)
    v.reset()
    obs, rew, dones, infos = v.step(np.zeros((8, 4), np.float32))
    assert obs.shape == (8, 56) and set(infos[0]) >= {"collision", "out_of_bounds", "env_complete", "duck_strike", "is_success"}
    v.close()


def test_ppo_on_the_duck_env_wide_observation():
    """train/train_objlock.py on the device: the 56-float observation runs through the 64-wide builds of the tcgen05 forward
    and of the fused gradient kernel, and through the value / bootstrap / running-moment kernels built for observations up
    to 64 floats wide.  The CUDA-core forward must agree with the fp32 torch towers to fp32 rounding, the tcgen05 forward
    within the TF32 tolerance, the running moments with NumPy, and a short run must improve the return."""
    import torch
    from pyflyt_drone_b200.ppo import PPO
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    env = FixedwingVecEnv(1024, preset="objlock_duck", seed=3)
    assert PPO("MlpPolicy", env, n_steps=32, batch_size=8192, seed=3).tensor_core_forward is True
    m = PPO("MlpPolicy", env, n_steps=32, batch_size=8192, n_epochs=4, seed=3, use_cuda_graph=False, tensor_core_forward=False)
    assert m.d == 56 and m.a == 4 and m.update == "kernel" and m.tensor_core_forward is False
    assert m.policy.count == m.policy.theta.numel() == int(m.lib.ppo_param_count(56))
    with torch.no_grad():
        m.policy.theta.add_(0.05 * torch.randn(m.policy.count, device=m.device, generator=m._gen))
    m.collect_rollouts()
    torch.cuda.synchronize()
    b = m.buf
    obs, act = b["obs"].view(-1, 56), b["act"].view(-1, 4)
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        v, lp, _ = m.policy.evaluate_actions(obs, act)
    dl, dv = (lp - b["logp"].view(-1)).abs(), (v - b["val"].view(-1)).abs()
    assert float(dl.max()) < 2e-4 and float(dl.mean()) < 2e-5, (float(dl.max()), float(dl.mean()))
    assert float(dv.max()) < 2e-4, float(dv.max())
    assert float(obs.abs().max()) <= 10.0 + 1e-6                      # VecNormalize clip on all 56 columns
    # running moments of a 56-column batch (two 32-column passes) against NumPy
    from pyflyt_drone_b200 import _lib
    from pyflyt_drone_b200.ppo import DeviceVecNormalize, _p, _stream
    vn = DeviceVecNormalize(56, 3000, m.device)
    x = (np.random.default_rng(0).normal(size=(3000, 56)) * np.linspace(2.0, 30, 56) + np.linspace(-50, 50, 56)).astype(np.float32)
    _lib.check(m.lib.ppo_moments_update(_p(torch.from_numpy(x).to(m.device)), 3000, 56, _p(vn.obs_stats), _p(vn.obs_scratch),
                                        _p(vn.obs_accum), _stream()))
    got = vn.obs_stats.cpu().numpy()
    bm, bv = x.astype(np.float64).mean(0), x.astype(np.float64).var(0)
    tot = 1e-4 + 3000
    assert np.allclose(got[:56], bm * 3000 / tot, rtol=1e-6, atol=1e-6)
    assert np.allclose(got[56:112], (1e-4 + bv * 3000 + bm ** 2 * 1e-4 * 3000 / tot) / tot, rtol=1e-4)
    # the 64-wide tcgen05 forward against the CUDA-core one on the same rows and noise stream (TF32 tolerance, as for D = 28)
    raw = (torch.randn(5000, 56, device=m.device, generator=m._gen) * 20).contiguous()
    stats = m.vecnorm.obs_stats.clone()
    outs = {}
    for name, fn in (("tc", m.lib.ppo_policy_forward_tc_a), ("cc", m.lib.ppo_policy_forward_a)):
        o = dict(obs_norm=torch.zeros(5000, 56, device=m.device), act_env=torch.zeros(5000, 4, device=m.device),
                 act_raw=torch.zeros(5000, 4, device=m.device), logp=torch.zeros(5000, device=m.device),
                 val=torch.zeros(5000, device=m.device))
        _lib.check(fn(_p(m.policy.theta), 56, 4, _p(raw), _p(stats), 10.0, 5000, 77, 5, 3, None, 0, _p(o["obs_norm"]),
                      _p(o["act_env"]), _p(o["act_raw"]), _p(o["logp"]), _p(o["val"]), _stream()))
        outs[name] = o
    torch.cuda.synchronize()
    assert torch.equal(outs["tc"]["obs_norm"], outs["cc"]["obs_norm"]) and torch.equal(outs["tc"]["logp"], outs["cc"]["logp"])
    assert float((outs["tc"]["val"] - outs["cc"]["val"]).abs().max()) < 1e-2
    assert float((outs["tc"]["act_raw"] - outs["cc"]["act_raw"]).abs().max()) < 1e-2
    # a six-channel policy has no 64-wide build: the entry point refuses instead of mis-staging
    assert m.lib.ppo_policy_forward_tc_a(_p(m.policy.theta), 56, 6, _p(obs), None, 10.0, 1024, 0, 0, 0, None, 0, None,
                                         _p(m.act_env), None, None, _p(b["val"][0]), _stream()) == -1
    m2 = PPO("MlpPolicy", env, n_steps=32, batch_size=8192, n_epochs=4, seed=3)
    r0, _, l0, _ = m2.evaluate_policy(n_eval_episodes=512)
    m2.learn(30 * 32 * 1024)
    r1, _, l1, _ = m2.evaluate_policy(n_eval_episodes=512)
    print(f"\n[duck ppo] raw return per episode {r0:.1f} -> {r1:.1f}, episode length {l0:.0f} -> {l1:.0f}")
    assert r1 > r0, (r0, r1, l0, l1)
    env.close()
