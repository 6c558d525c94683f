"""Host-lane overhead breakdown (65,536 Waypoints envs)."""
import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
N = 65536
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (N, 4)).astype(np.float32) for _ in range(4)]
env = FixedwingVecEnv(N, preset="waypoints_v3", seed=1)
env.reset()
def timeit(f, n=200):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
k = [0]
def a():
    k[0] += 1; env.step_arrays(acts[k[0] % 4], want_terminal_obs=False)
def b():
    env.step_arrays(env.action_buffer, want_terminal_obs=False)
def c():
    k[0] += 1; np.copyto(env.action_buffer, acts[k[0] % 4])
def d():
    env.lib.fw_num_envs(env._h)
print(f"step_arrays(user array):     {timeit(a):7.1f} us")
print(f"step_arrays(action_buffer):  {timeit(b):7.1f} us")
print(f"np.copyto 1 MB:              {timeit(c):7.1f} us")
print(f"trivial ctypes call:         {timeit(d):7.1f} us")
env.close()
