"""e2e lane probe: zero-copy mapped host buffers (fw_step_host) vs explicit pinned-memory DMA copies around the
device-lane kernel, 65,536 Waypoints envs, host numpy actions in / obs+reward+flags out."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pyflyt_drone_b200.vec_env import FixedwingVecEnv

N = 65536
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (N, 4)).astype(np.float32) for _ in range(4)]

env = FixedwingVecEnv(N, preset="waypoints_v3", seed=1)
env.reset()
for s in range(5): env.step_arrays(acts[s % 4], want_terminal_obs=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(200): env.step_arrays(acts[s % 4], want_terminal_obs=False)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"zero-copy mapped buffers: {dt / 200 * 1e6:8.1f} us/step  {N * 200 / dt:.3e} env-steps/s")
env.close()

env = FixedwingVecEnv(N, preset="waypoints_v3", seed=1)
env.reset_tensor()
a_pin = torch.empty(N, 4, dtype=torch.float32).pin_memory()
o_pin = torch.empty(N, 28, dtype=torch.float32).pin_memory()
r_pin = torch.empty(N, dtype=torch.float32).pin_memory()
f_pin = torch.empty(N, dtype=torch.uint8).pin_memory()
a_dev = torch.empty(N, 4, device="cuda")
def step(a):
    a_pin.numpy()[...] = a
    a_dev.copy_(a_pin, non_blocking=True)
    out = env.step_tensor(a_dev)
    obs, rew, flags = out[0], out[1], out[2]
    o_pin.copy_(obs, non_blocking=True); r_pin.copy_(rew, non_blocking=True); f_pin.copy_(flags, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return o_pin.numpy(), r_pin.numpy(), f_pin.numpy()
for s in range(5): step(acts[s % 4])
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(200): step(acts[s % 4])
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pinned DMA copies:        {dt / 200 * 1e6:8.1f} us/step  {N * 200 / dt:.3e} env-steps/s")
# copy-only timings
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(200):
    o_pin.copy_(out_obs := env._t["obs"] if hasattr(env, "_t") and env._t else o_pin, non_blocking=True) if False else None
dev_obs = torch.empty(N, 28, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(200):
    o_pin.copy_(dev_obs, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"D2H 7.3 MB DMA alone:     {dt / 200 * 1e6:8.1f} us  {N * 28 * 4 / (dt / 200) * 1e-9:.1f} GB/s")
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(200):
    a_dev.copy_(a_pin, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D 1 MB DMA alone:       {dt / 200 * 1e6:8.1f} us  {N * 16 / (dt / 200) * 1e-9:.1f} GB/s")
env.close()
