import sys
sys.path.insert(0, '/root/repo')
import torch
import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
env = FixedwingVecEnv(65536, config=fw.waypoint_objlock(), seed=1)
env.reset_tensor()
env.step_random(12)
torch.cuda.synchronize()
env.close()
