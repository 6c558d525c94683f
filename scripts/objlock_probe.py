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
for label, kw in (("default", {}), ("res16", dict(cam_res=16)), ("res64", dict(cam_res=64)), ("obst5", dict(num_obstacles=5)),
                  ("obst10", dict(num_obstacles=10)), ("interval48", dict(cam_interval_substeps=48)),
                  ("far_obst", dict(obst_radius=0.01))):
    cfg = fw.waypoint_objlock(**kw)
    env = FixedwingVecEnv(N, config=cfg, seed=1)
    env.reset_tensor()
    t_rand = timeit(lambda: env.step_random(1))
    st = env.episode_stats()
    print(f"{label:16s} step_random {t_rand:.3f} ms  episodes {st['episodes']:.0f}")
    env.close()
