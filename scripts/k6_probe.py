"""Warm-cache timings of the update kernels at several minibatch sizes (CUDA events around 200 back-to-back calls).
usage: k6_probe.py [preset]"""
import sys
sys.path.insert(0, '/root/repo')
import torch
from pyflyt_drone_b200 import _lib
from pyflyt_drone_b200.ppo import PPO, _p, _stream
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
preset = sys.argv[1] if len(sys.argv) > 1 else "waypoints_v3"
env = FixedwingVecEnv(4096, preset=preset, seed=1)
m = PPO("MlpPolicy", env, n_steps=128, batch_size=32768, n_epochs=1, seed=1)
m.collect_rollouts(); m.collect_rollouts(); m.collect_rollouts()
total = 4096 * 128
perm = torch.randperm(total, device="cuda")
stats = torch.zeros(8, device="cuda")


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1000 / reps


for bs in (128, 1024, 4096, 18944, 32768, 131072, 524288):
    idx = perm[:bs]
    g = torch.cuda.CUDAGraph()
    fn = lambda: m._minibatch_grad_kernel(idx, m._grad, stats)
    fn(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    t = timed(g.replay, 20) / 20
    print(f"batch {bs:7d}: grad call (adv stats + K6 + reduce) {t:7.1f} us in a graph; eager {timed(fn):7.1f} us")
theta = m.policy.theta.data
fa = lambda: _lib.check(m.lib.ppo_adam_step(_p(theta), _p(m._grad), _p(m._adam_m), _p(m._adam_v), m.policy.count, 0.0, 0.9, 0.999, 1e-5, 0.5,
                                            1.0, _p(m._adam_t), _p(m._grad_norm), _stream()))
print(f"adam eager {timed(fa):.1f} us")
env.close()
