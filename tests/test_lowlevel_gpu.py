"""GPU parity of the low-level tracking task (FixedwingLowLevelEnv) against the fp64 oracle, through the
C ABI.  Same bar as tests/test_parity_gpu.py: per-step state within 1e-4 relative (norm-wise per observation group,
floor 1) from an identical injected state, flags exact, reward within 1e-4 of its scale."""
import numpy as np
import pytest

import pyflyt_drone_b200 as fw

pytestmark = pytest.mark.gpu
RTOL = 1e-4
GROUPS = {"ang_vel": slice(0, 3), "euler": slice(3, 6), "lin_vel": slice(6, 9), "pos": slice(9, 12),
          "prev_action": slice(12, 18), "target": slice(18, 21)}


@pytest.fixture(scope="module")
def fo(oracle_mod):
    return oracle_mod


def pair(fo, n, cfg, seed=7, env_id0=0):
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    return FixedwingVecEnv(n, config=cfg, seed=seed, env_id0=env_id0), fo.OracleVecEnv(cfg.as_dict(), n, seed=seed, env_id0=env_id0)


def group_err(got, ref):
    d = got - ref
    d[:, 3:6] = (d[:, 3:6] + np.pi) % (2 * np.pi) - np.pi
    out = {}
    for k, sl in GROUPS.items():
        scale = np.maximum(np.abs(ref[:, sl]).max(axis=1), 1.0)
        out[k] = float((np.abs(d[:, sl]).max(axis=1) / scale).max())
    return out


def test_lowlevel_reset_parity(fo):
    env, orc = pair(fo, 300, fw.lowlevel())
    assert env.obs_dim == 21 and env.act_dim == 6 and env.action_space.shape == (6,)
    e = group_err(env.reset().astype(np.float64), orc.reset())
    assert max(e.values()) < RTOL, e
    env.close()


@pytest.mark.parametrize("variant", ["plain", "noise", "wind"])
def test_lowlevel_single_step_parity_from_injected_state(fo, variant):
    wind = None
    kw = dict(noise_ratio=0.0)
    if variant == "noise":
        kw = dict(noise_ratio=0.02)
    if variant == "wind":
        wind = dict(enabled=True, mode="gust_sine", wind_enu_mps_range=[[-5, 5], [-5, 5], [-0.5, 0.5]],
                    gust_amp_enu_mps_range=[[0, 3], [0, 3], [0, 0.3]], gust_freq_hz=0.2, randomize_on_reset=True,
                    randomize_gust_phase=True)
    cfg = fw.lowlevel(wind=wind, **kw)
    N = 512
    env, orc = pair(fo, N, cfg)
    env.reset(); orc.reset()
    rng = np.random.default_rng(2)
    worst, n_done = {}, 0
    for k in range(60):
        st = orc.get_state()
        if k % 20 == 10:                      # push a few envs to the band edges / the step limit
            st["pos"][:8, 2] = 1.02
            st["pos"][8:16, 2] = 99.9
            st["step_count"][16:24] = 1999
            orc.set_state(st)
            st = orc.get_state()
        env.set_state(st)
        a = rng.uniform(-1, 1, (N, 6)).astype(np.float32)
        og, rg, fg, tg = env.step_arrays(a)
        og, rg, fg, tg = og.astype(np.float64), rg.astype(np.float64), fg.astype(np.int32), tg.astype(np.float64)
        oc, rc, fc, tc = orc.step(a.astype(np.float64))
        assert np.array_equal(fg, fc), (k, np.nonzero(fg != fc))
        done = (fc & 3) != 0
        n_done += int(done.sum())
        e = group_err(og, oc)
        if done.any():
            et = group_err(tg[done], tc[done])
            e = {g: max(e[g], et[g]) for g in e}
        for g, v in e.items():
            worst[g] = max(worst.get(g, 0.0), v)
        assert np.abs(rg - rc).max() <= 1e-4 * max(1.0, np.abs(rc).max()), k
    print(f"\n[lowlevel/{variant}] worst norm-wise relative error per group: "
          + ", ".join(f"{g}={v:.2e}" for g, v in worst.items()) + f"; done events {n_done}")
    assert n_done > 0 and max(worst.values()) < RTOL, worst
    env.close()


def test_lowlevel_random_action_lane_and_tensor_lane(fo):
    import torch
    cfg = fw.lowlevel(noise_ratio=0.02)
    env, orc = pair(fo, 300, cfg, seed=5, env_id0=1000)
    env.reset(); orc.reset()
    env.step_random(40)
    orc.rollout_random(40)
    sg, sc = env.get_state(), orc.get_state()
    assert np.array_equal(sg["episode"], sc["episode"]) and np.array_equal(sg["step_count"], sc["step_count"])
    assert np.abs(sg["pos"] - sc["pos"]).max() < 2e-3 and np.abs(sg["quat"] - sc["quat"]).max() < 1e-3
    # device-tensor lane == host lane, bit for bit
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    a_env, b_env = FixedwingVecEnv(200, config=cfg, seed=1), FixedwingVecEnv(200, config=cfg, seed=1)
    assert np.array_equal(a_env.reset(), b_env.reset_tensor().cpu().numpy())
    rng = np.random.default_rng(0)
    for _ in range(5):
        a = rng.uniform(-1, 1, (200, 6)).astype(np.float32)
        o1, r1, f1, _ = a_env.step_arrays(a)
        out = b_env.step_tensor(torch.from_numpy(a).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(o1, out[0].cpu().numpy()) and np.array_equal(r1, out[1].cpu().numpy())
        assert np.array_equal(f1, out[2].cpu().numpy())
    a_env.close(); b_env.close(); env.close()


def test_ppo_on_the_lowlevel_env_six_channel_policy():
    """train/train_lowlevel_cmd.py on the device: rollouts with the forward kernels compiled for a 6-channel Gaussian
    policy (CUDA-core fp32 and tcgen05 TF32), update through the fused tcgen05 gradient kernel compiled for action width 6.
    The forward kernels must agree
    with the fp32 torch towers, and a short run must improve the return."""
    import torch
    from pyflyt_drone_b200.ppo import PPO
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    env = FixedwingVecEnv(1024, preset="lowlevel", seed=3)
    for tc, tol_max, tol_mean in ((False, 2e-4, 2e-5), (True, 3e-2, 3e-3)):          # fp32 kernel; TF32 + MUFU.TANH kernel
        m = PPO("MlpPolicy", env, n_steps=32, batch_size=8192, n_epochs=4, seed=3, use_cuda_graph=False,
                tensor_core_forward=tc)
        assert m.a == 6 and m.update == "kernel" and m.tensor_core_forward == tc and m.policy.count == m.policy.theta.numel()
        with torch.no_grad():      # move the towers away from the near-zero initial action head
            m.policy.theta.add_(0.05 * torch.randn(m.policy.count, device=m.device, generator=m._gen))
        m.collect_rollouts()
        torch.cuda.synchronize()
        b = m.buf
        obs, act = b["obs"].view(-1, 21), b["act"].view(-1, 6)
        with torch.no_grad():
            v, lp, _ = m.policy.evaluate_actions(obs, act)
        dl, dv = (lp - b["logp"].view(-1)).abs(), (v - b["val"].view(-1)).abs()
        assert float(dl.max()) < tol_max and float(dl.mean()) < tol_mean, (tc, float(dl.max()), float(dl.mean()))
        assert float(dv.max()) < tol_max, (tc, float(dv.max()))
        assert float(m.act_env.abs().max()) <= 1.0 and float(b["act"].std()) > 0.5       # sampled, clipped for the env
    m = PPO("MlpPolicy", env, n_steps=32, batch_size=8192, n_epochs=4, seed=3)
    b = m.buf
    r0, _, l0, _ = m.evaluate_policy(n_eval_episodes=512)
    m.learn(40 * 32 * 1024)
    r1, _, l1, _ = m.evaluate_policy(n_eval_episodes=512)
    det = m.predict(m._obs, deterministic=True)
    assert det.shape == (1024, 6)
    print(f"\n[lowlevel ppo] raw return per episode {r0:.1f} -> {r1:.1f}, episode length {l0:.0f} -> {l1:.0f}")
    # PPO maximises the return: with a tracking penalty of several units per step against a one-off -100 for leaving
    # the altitude band, the first thing 1.3 M steps teach is to end hopeless episodes early -- the return goes up
    assert r1 > r0 + 500.0, (r0, r1, l0, l1)
    env.close()
