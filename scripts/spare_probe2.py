import sys, torch
sys.path.insert(0, "/root/repo")
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
env = FixedwingVecEnv(65536, preset=sys.argv[1] if len(sys.argv) > 1 else "waypoint_objlock", seed=1)
env.step_random(330)
torch.cuda.synchronize()
print(env.spare_stats())
