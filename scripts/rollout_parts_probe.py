"""Warm per-kernel times of the PPO rollout step at a given env count (each call captured 50x in a CUDA graph).
usage: rollout_parts_probe.py [n_envs] [preset]"""
import sys
sys.path.insert(0, '/root/repo')
import torch
from pyflyt_drone_b200 import _lib
from pyflyt_drone_b200.ppo import PPO, _p, _stream
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
preset = sys.argv[2] if len(sys.argv) > 2 else "waypoints_v3"
env = FixedwingVecEnv(n, preset=preset, seed=1)
m = PPO("MlpPolicy", env, n_steps=8, batch_size=n * 2, n_epochs=1, seed=1, use_cuda_graph=False)
m.collect_rollouts()
vn, b = m.vecnorm, m.buf
obs, rew, flags, term = env.step_tensor(m.act_env, want_terminal_obs=True)


def timed(fn, reps=50):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) * 1000 / (5 * reps)


parts = {
    "policy forward": lambda: m._forward(m._obs, 0),
    "env step": lambda: env.step_tensor(m.act_env, want_terminal_obs=True),
    "obs moments": lambda: m._update_obs_moments(obs),
    "reward normalise (2 kernels)": lambda: _lib.check(m.lib.ppo_reward_normalize(
        _p(rew), _p(flags), n, vn.gamma, vn.clip_reward, _p(vn.ret), _p(vn.ret_stats), _p(vn.ret_scratch), _p(vn.ret_accum),
        _p(b["rew"][0]), _p(b["done"][0]), _stream())),
    "time-limit bootstrap": lambda: _lib.check(m.lib.ppo_timeout_bootstrap_a(
        _p(m.policy.theta), m.d, m.a, _p(term), m._stats_ptr(), vn.clip_obs, _p(flags), n, m.gamma, _p(b["rew"][0]), _stream())),
}
tot = 0.0
for k, f in parts.items():
    t = timed(f); tot += t
    print(f"{k:32s} {t:7.2f} us")
print(f"{'sum':32s} {tot:7.2f} us   ({n} envs, {preset})")
env.close()
