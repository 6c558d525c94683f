import sys, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
N = 65536
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n): fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n
cfg = fw.waypoint_objlock()
env = FixedwingVecEnv(N, config=cfg, seed=1)
env.reset_tensor()
a0 = torch.zeros(N, 4, device='cuda')
for rep in range(6):
    ar = (torch.randn(N, 4, device='cuda')).clamp(-1, 1)
    t1_ = timeit(lambda: env.step_tensor(ar, want_terminal_obs=True), n=10)
    st = env.episode_stats()
    s = env.get_state()
    print(f"rep {rep}: step(gauss act, term obs) {t1_:.3f} ms episodes {st['episodes']:.0f} mean len {st['length_sum']/max(st['episodes'],1):.1f} "
          f"targets/ep {st['targets_reached_sum']/max(st['episodes'],1):.2f} coll {st['collisions']:.0f} oob {st['out_of_bounds']:.0f} "
          f"tidx>=8 {(s['target_idx']>=8).sum()} duck_phase {s['ol_i'][:,0].sum()} visible {s['ol_i'][:,4].sum()}")
t = timeit(lambda: env.step_random(1)); print("step_random", t)
env.close()
