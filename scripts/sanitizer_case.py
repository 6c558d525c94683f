"""The smallest run that launches every kernel family once -- what `compute-sanitizer --tool memcheck|racecheck` is pointed
at (one tool per gpurun call, see B200_PROFILING.md):

    compute-sanitizer --tool memcheck  python scripts/sanitizer_case.py > profiles/r2_sanitizer_memcheck.log
    compute-sanitizer --tool racecheck python scripts/sanitizer_case.py > profiles/r2_sanitizer_racecheck.log

K1 (Waypoints step, host and device lanes, odd batch so that the partial last warp / CTA paths run), K1p (packed pairs),
K3 (Waypoint-ObjLock and duck-only steps with their shared-memory obstacle tables, depth rows and vision history, plus
resets), the low-level head, and one PPO iteration at N = 1,024 (K4 tcgen05 forward, moments, reward normalisation,
bootstrap, GAE, permutation, K6 gradient, reduce, Adam)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyflyt_drone_b200 as fw  # noqa: E402
from pyflyt_drone_b200.ppo import PPO  # noqa: E402
from pyflyt_drone_b200.vec_env import FixedwingVecEnv  # noqa: E402

N = int(os.environ.get("SAN_N", "1024"))
rng = np.random.default_rng(0)
for preset, kw, n in (("waypoints_v3", {}, N + 37), ("waypoints_v3", {"packed_pairs": 1}, N + 37), ("physics_only", {}, N),
                      ("lowlevel", {}, N), ("waypoint_objlock", {}, N), ("objlock_duck", {}, N // 2)):
    env = FixedwingVecEnv(n, config=fw.make_config(preset, **kw), seed=5)
    if env.obs_dim:
        env.reset()
        for _ in range(3):
            env.step_arrays(rng.uniform(-1, 1, (n, env.act_dim)).astype(np.float32))
        obs = env.reset_tensor()
        for _ in range(2):
            env.step_tensor(torch.rand((n, env.act_dim), device="cuda") * 2 - 1, want_terminal_obs=True)
    env.step_random(3)
    torch.cuda.synchronize()
    print(preset, kw, "ok", env.launch_count, "launches, faults", env.fault_count(), flush=True)
    env.close()
env = FixedwingVecEnv(N, preset="waypoints_v3", seed=3)
model = PPO("MlpPolicy", env, n_steps=4, batch_size=N * 4 // 2, n_epochs=2, seed=3, use_cuda_graph=False)
model.learn(N * 4)
torch.cuda.synchronize()
assert torch.isfinite(model.policy.theta).all()
print("ppo ok", flush=True)
env.close()
