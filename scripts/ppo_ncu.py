"""Small PPO run for ncu captures of the PPO kernels.  usage: ppo_ncu.py [n_steps] [batch_size] [preset]
default: 65,536 envs, n_steps 4, 65,536-row minibatches, 1 epoch, no CUDA graph (ncu sees individual launches)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from pyflyt_drone_b200.ppo import PPO
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
preset = sys.argv[3] if len(sys.argv) > 3 else "waypoints_v3"
env = FixedwingVecEnv(65536, preset=preset, seed=1)
m = PPO("MlpPolicy", env, n_steps=n_steps, batch_size=batch, n_epochs=1, seed=1, use_cuda_graph=False)
m.learn(2 * n_steps * 65536)
torch.cuda.synchronize()
env.close()
print("ok")
