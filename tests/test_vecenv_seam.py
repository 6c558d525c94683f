"""The B1 seam (SURVEY section 8 b): what stable-baselines3 code on either side of a VecEnv expects from it.

The reference trains through ``VecNormalize(SubprocVecEnv(...))`` and evaluates through ``evaluate_policy(model, eval_env,
callback=WaypointEvalCallback._log_success_callback)`` (/root/reference/train/train_Fixedwing_Waypoints_v3.py:124-172,
251-260).  stable_baselines3 is not installable here, so this file restates the two pieces of SB3 that touch the seam
-- a ``VecEnvWrapper`` that passes observations through, and the per-env loop of ``evaluate_policy`` that hands
``locals()`` to the callback after every env's step -- and drives them over ``FixedwingVecEnv``:

  * on the CPU with the device calls of the class replaced by the fp64 oracle (tests may use it), which exercises every
    host-side line of the seam: ``step_async/step_wait``, the info dicts, ``get_attr/set_attr/env_method``, ``seed``;
  * on the GPU (``-m gpu``) with the real thing, checked against the oracle.
"""
from __future__ import annotations

import numpy as np
import pytest

import pyflyt_drone_b200 as fw
from oracle import fw_oracle as fo
from pyflyt_drone_b200.compat import spaces
from pyflyt_drone_b200.compat.vec_env import VecEnv, VecEnvWrapper
from pyflyt_drone_b200.ppo import PPO
from pyflyt_drone_b200.vec_env import FixedwingVecEnv


class OracleBackedVecEnv(FixedwingVecEnv):
    """FixedwingVecEnv with the two calls into libfwsim.so (create + step/reset) answered by the fp64 oracle."""

    def _open(self) -> None:
        self._orc = fo.OracleVecEnv(self.cfg.as_dict(), self.num_envs, seed=self._seed, env_id0=self.env_id0)
        self._h = None
        self.obs_dim, self.act_dim = self._orc.obs_dim, self._orc.act_dim
        self.observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=(self.obs_dim,), dtype=np.float32)
        self.action_space = spaces.Box(low=-1.0, high=1.0, shape=(self.act_dim,), dtype=np.float32)
        n, D = self.num_envs, max(self.obs_dim, 1)
        self._h_act = np.zeros((n, self.act_dim), np.float32)
        self._h_obs = np.zeros((n, D), np.float32)
        self._h_rew = np.zeros(n, np.float32)
        self._h_flags = np.zeros(n, np.uint8)
        self._h_term = np.zeros((n, D), np.float32)
        self._h_tidx = np.zeros(n, np.uint8)

    def reset(self):
        self._h_obs[:, : self.obs_dim] = self._orc.reset()
        self.reset_infos = [{} for _ in range(self.num_envs)]
        return self._h_obs[:, : self.obs_dim].copy()

    def step_arrays(self, actions, want_terminal_obs=True):
        obs, rew, flags, term = self._orc.step(np.asarray(actions, np.float64))
        self._h_obs[:, : self.obs_dim] = obs
        self._h_rew[:] = rew
        self._h_flags[:] = flags
        self._h_term[:, : self.obs_dim] = term
        self._h_tidx[:] = self._orc.last_targets_reached
        return self._h_obs[:, : self.obs_dim], self._h_rew, self._h_flags, self._h_term[:, : self.obs_dim]

    def seed(self, seed=None):
        if seed is not None:
            self._seed = int(seed)
            self._open()
        return [self._seed + self.env_id0 + i for i in range(self.num_envs)]

    def fault_count(self) -> int:
        return 0

    def close(self):
        self._closed = True


class PassThroughNormalize(VecEnvWrapper):
    """The shape of SB3's VecNormalize as a wrapper: reset/step_wait forward to ``venv`` (statistics left out)."""

    def reset(self):
        return self.venv.reset()

    def step_wait(self):
        obs, rew, dones, infos = self.venv.step_wait()
        for i in np.nonzero(dones)[0]:                       # VecNormalize normalises infos[i]["terminal_observation"] too
            assert "terminal_observation" in infos[i]
        return obs, rew, dones, infos


class ReachRateCallback:
    """WaypointEvalCallback._log_success_callback restated (train_Fixedwing_Waypoints_v3.py:130-140)."""

    def __init__(self):
        self.buffer: list[int] = []

    def __call__(self, locals_: dict, globals_: dict) -> None:
        info = locals_.get("info")
        if isinstance(info, (list, tuple)) and len(info) > 0:
            info = info[0]
        if locals_.get("done") and isinstance(info, dict):
            n = info.get("num_targets_reached")
            if n is not None:
                self.buffer.append(int(n))


def sb3_style_evaluate(policy_fn, env, n_eval_episodes, callback):
    """The loop of stable_baselines3.common.evaluation.evaluate_policy: per-env episode quotas, and after every env's step
    ``callback(locals(), globals())`` with that env's ``reward``, ``done`` and ``info`` in scope."""
    n_envs = env.num_envs
    episode_rewards, episode_lengths = [], []
    episode_counts = np.zeros(n_envs, dtype="int")
    episode_count_targets = np.array(PPO.episode_quotas(n_eval_episodes, n_envs), dtype="int")
    current_rewards, current_lengths = np.zeros(n_envs), np.zeros(n_envs, dtype="int")
    observations = env.reset()
    while (episode_counts < episode_count_targets).any():
        actions = policy_fn(observations)
        observations, rewards, dones, infos = env.step(actions)
        current_rewards += rewards
        current_lengths += 1
        for i in range(n_envs):
            if episode_counts[i] < episode_count_targets[i]:
                reward, done, info = rewards[i], dones[i], infos[i]      # noqa: F841 - read by the callback via locals()
                callback(locals(), globals())
                if dones[i]:
                    episode_rewards.append(current_rewards[i]); episode_lengths.append(current_lengths[i])
                    episode_counts[i] += 1
                    current_rewards[i] = 0; current_lengths[i] = 0
    return episode_rewards, episode_lengths


def _seam_checks(make_env):
    """Shared by the CPU (oracle-backed) and GPU runs: targets are 'reached' on every inner iteration (goal radius larger
    than the dome), so each episode reaches all 8 waypoints within two steps and ends by truncation with
    info["num_targets_reached"] == 8 -- the value a post-reset read would report as 0."""
    n = 6
    env = make_env(n, goal_reach=1.0e4, noise_ratio=0.0)
    assert isinstance(env, VecEnv)
    wrapped = PassThroughNormalize(env)
    assert wrapped.num_envs == n and wrapped.observation_space.shape == (28,)
    cb = ReachRateCallback()
    rng = np.random.default_rng(0)
    rewards, lengths = sb3_style_evaluate(lambda o: rng.uniform(-1, 1, (n, 4)).astype(np.float32), wrapped, 9, cb)
    assert len(rewards) == 9 and sorted(PPO.episode_quotas(9, n)) == [1, 1, 1, 2, 2, 2]
    assert cb.buffer == [8] * 9, cb.buffer                      # every finished episode reported all 8 targets
    assert all(ln == 2 for ln in lengths)
    # info keys of the reference env on a finished episode, plus SB3's two
    obs, rew, dones, infos = wrapped.step(np.zeros((n, 4), np.float32))
    obs, rew, dones, infos = wrapped.step(np.zeros((n, 4), np.float32))
    assert dones.all()
    for d in infos:
        assert {"out_of_bounds", "collision", "env_complete", "num_targets_reached", "terminal_observation",
                "TimeLimit.truncated"} <= set(d)
        assert d["env_complete"] and d["num_targets_reached"] == 8 and d["TimeLimit.truncated"]
        assert d["terminal_observation"].shape == (28,)
    # get_attr / set_attr / env_method / env_is_wrapped as SB3 code calls them through a wrapper
    assert wrapped.get_attr("num_targets") == [8] * n
    assert wrapped.get_attr("render_mode", indices=[0, 2]) == [None, None]
    wrapped.set_attr("curriculum_level", 3, indices=[1, 4])
    assert wrapped.get_attr("curriculum_level") == [None, 3, None, None, 3, None]
    with pytest.raises(AttributeError):
        wrapped.set_attr("goal_reach", 2.0)                      # compiled into the batch
    with pytest.raises(AttributeError):
        wrapped.get_attr("no_such_attribute")
    with pytest.raises(AttributeError):
        wrapped.env_method("no_such_method")
    assert wrapped.env_is_wrapped(PassThroughNormalize) == [False] * n
    assert wrapped.env_method("fault_count", indices=[0, 1]) == [0, 0]
    # seed(s) always rebuilds: the first episode after it is the same every time, also for an unchanged seed
    a = wrapped.seed(7); first = wrapped.reset().copy()
    wrapped.step(np.zeros((n, 4), np.float32))
    b = wrapped.seed(7); again = wrapped.reset().copy()
    assert a == b == [7 + i for i in range(n)]
    np.testing.assert_array_equal(first, again)
    wrapped.close()


def test_seam_on_cpu_with_the_oracle_behind_the_class():
    _seam_checks(lambda n, **over: OracleBackedVecEnv(n, config=fw.make_config("waypoints_v3", **over), seed=3))


def test_infos_carry_pre_reset_targets_reached_in_full_and_lazy_mode():
    cfg = fw.make_config("waypoints_v3", noise_ratio=0.0, goal_reach=1.0e4)
    for mode in ("full", "lazy"):
        env = OracleBackedVecEnv(4, config=cfg, seed=1, info_mode=mode)
        env.reset()
        _, _, dones, infos = env.step(np.zeros((4, 4), np.float32))
        assert not dones.any()
        got = [d.get("num_targets_reached") for d in infos]
        assert got == ([4] * 4 if mode == "full" else [None] * 4)     # lazy: one shared blank dict for running envs ...
        np.testing.assert_array_equal(env.last_targets_reached, [4] * 4)   # ... the array view always has the value
        _, _, dones, infos = env.step(np.zeros((4, 4), np.float32))
        assert dones.all() and [d["num_targets_reached"] for d in infos] == [8] * 4


def test_episode_quotas_match_sb3():
    assert PPO.episode_quotas(20, 4) == [5, 5, 5, 5]
    assert PPO.episode_quotas(10, 4) == [2, 2, 3, 3]
    assert PPO.episode_quotas(3, 8) == [0, 0, 0, 0, 0, 1, 1, 1]
    assert sum(PPO.episode_quotas(100, 64)) == 100


@pytest.mark.gpu
def test_seam_on_gpu():
    _seam_checks(lambda n, **over: FixedwingVecEnv(n, config=fw.make_config("waypoints_v3", **over), seed=3))


@pytest.mark.gpu
def test_targets_reached_matches_oracle_on_both_lanes():
    import torch
    cfg = fw.make_config("waypoints_v3", noise_ratio=0.0, goal_reach=60.0)      # reachable, but not at once
    n = 256
    env, orc = FixedwingVecEnv(n, config=cfg, seed=5), fo.OracleVecEnv(cfg.as_dict(), n, seed=5)
    dev = FixedwingVecEnv(n, config=cfg, seed=5)
    env.reset(); orc.reset(); dev.reset_tensor()
    rng = np.random.default_rng(1)
    seen = 0
    for _ in range(60):
        a = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
        _, _, flags, _ = env.step_arrays(a)
        _, _, fo_, _ = orc.step(a.astype(np.float64))
        np.testing.assert_array_equal(flags.astype(np.int32), fo_)
        np.testing.assert_array_equal(env.last_targets_reached.astype(np.int32), orc.last_targets_reached)
        dev.step_tensor(torch.from_numpy(a).cuda())
        np.testing.assert_array_equal(dev.targets_reached_tensor().cpu().numpy(), env.last_targets_reached)
        done = (flags & 3) != 0
        seen += int((env.last_targets_reached[done] > 0).sum())
    assert seen > 0, "no finished episode had reached a target: the pre-reset value was never exercised"
    assert env.fault_count() == 0
    env.close(); dev.close()


@pytest.mark.gpu
def test_non_finite_state_is_flagged_counted_and_reset():
    cfg = fw.make_config("waypoints_v3", noise_ratio=0.0)
    n = 64
    env = FixedwingVecEnv(n, config=cfg, seed=2)
    env.reset()
    st = env.get_state()
    st["vel"][5, 0] = np.nan
    st["pos"][9, 2] = np.inf
    env.set_state(st)
    obs, rew, flags, term = env.step_arrays(np.zeros((n, 4), np.float32))
    bad = np.zeros(n, bool); bad[[5, 9]] = True
    assert ((flags & fw.config.FLAG_FAULT) != 0).tolist() == bad.tolist()
    assert ((flags[bad] & fw.config.FLAG_TERM) != 0).all() and (rew[bad] == 0).all()
    assert np.isfinite(obs).all() and np.isfinite(term[bad]).all() and np.isfinite(rew).all()
    assert env.fault_count() == 2
    after = env.get_state()
    assert np.isfinite(after["pos"]).all() and np.isfinite(after["vel"]).all()
    assert (after["step_count"][bad] == 0).all()               # fresh episodes
    infos = env._make_infos(flags, (flags & 3) != 0, term)
    assert infos[5]["state_fault"] and "state_fault" not in infos[0]
    env.close()
