import sys, time, torch
sys.path.insert(0, "/root/repo")
import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
for preset in ("waypoint_objlock", "objlock_duck"):
    env = FixedwingVecEnv(65536, preset=preset, seed=1)
    env.step_random(300)
    torch.cuda.synchronize()
    s0 = env.spare_stats(); env.episode_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step_random(200); e1.record(); torch.cuda.synchronize()
    s1 = env.spare_stats(); ep = env.episode_stats()
    print(preset, "us/step", e0.elapsed_time(e1) / 200 * 1e3, "spare", s1["from_spare"] - s0["from_spare"], "inline", s1["inline"] - s0["inline"], "episodes", ep["episodes"])
    env.close()
