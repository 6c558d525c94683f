"""Small PPO iteration for ncu captures of the tensor-core kernels (65,536 envs, 4 steps, 1 epoch)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from pyflyt_drone_b200.ppo import PPO
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
env = FixedwingVecEnv(65536, preset="waypoints_v3", seed=1)
m = PPO("MlpPolicy", env, n_steps=4, batch_size=65536, n_epochs=1, seed=1, use_cuda_graph=False)
m.learn(2 * 4 * 65536)
torch.cuda.synchronize()
env.close()
print("ok")
