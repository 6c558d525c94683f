"""ncu driver: ObjLock random-action steps in steady state (episodes desynchronised, aircraft spread over the dome).
Usage under ncu: --launch-skip 150 --launch-count 1 (the 151-st launch of this process is a steady-state step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
preset = sys.argv[1] if len(sys.argv) > 1 else "waypoint_objlock"
env = FixedwingVecEnv(65536, config=fw.make_config(preset), seed=1)
env.reset_tensor()
env.step_random(160)
torch.cuda.synchronize()
env.close()
