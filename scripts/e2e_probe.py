"""Host-lane overhead breakdown (65,536 Waypoints envs); FWSIM_HOST_CHUNKS=<n> selects the chunk count."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
N = 65536
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (N, 4)).astype(np.float32) for _ in range(4)]
pinned = []
for a in acts:
    t = torch.empty((N, 4), dtype=torch.float32, pin_memory=True); t.numpy()[:] = a; pinned.append(t)
pin_np = [t.numpy() for t in pinned]
env = FixedwingVecEnv(N, preset="waypoints_v3", seed=1)
env.reset()
def timeit(f, n=300):
    for _ in range(10): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
k = [0]
def a():
    k[0] += 1; env.step_arrays(acts[k[0] % 4], want_terminal_obs=False)
def b():
    k[0] += 1; env.step_arrays(pin_np[k[0] % 4], want_terminal_obs=False)
def d():
    env.lib.fw_num_envs(env._h)
print(f"chunks={os.environ.get('FWSIM_HOST_CHUNKS', 'default')}: step_arrays(pageable) {timeit(a):7.1f} us, "
      f"step_arrays(pinned) {timeit(b):7.1f} us, trivial ctypes call {timeit(d):5.1f} us")
env.close()
# PCIe reference: plain pinned copies of the same sizes
dev = torch.empty(N * 28, dtype=torch.float32, device="cuda"); host = torch.empty(N * 28, dtype=torch.float32, pin_memory=True)
big_d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda"); big_h = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
def c1(): host.copy_(dev, non_blocking=True)
def c2(): big_h.copy_(big_d, non_blocking=True)
t1, t2 = timeit(c1, 100), timeit(c2, 30)
print(f"pinned D2H: 7.3 MB in {t1:.1f} us ({N * 28 * 4 / t1 / 1e3:.1f} GB/s), 64 MiB in {t2:.1f} us ({(64 << 20) / t2 / 1e3:.1f} GB/s)")
