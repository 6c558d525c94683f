"""Where does the per-launch fixed cost of the env step go?  T(spl) for several steps-per-launch, with the state
rotating through DRAM (41 batches) and resident in L2 (1 batch)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.vec_env import FixedwingVecEnv

N = 65536
for replicas in (41, 1):
    envs = [FixedwingVecEnv(N, config=fw.physics_only(), seed=1, env_id0=r * N) for r in range(replicas)]
    for e in envs:
        e.reset_tensor()
    for spl in (1, 2, 4, 8):
        for use_graph in (True, False):
            n_launch = 410 if replicas > 1 else 400
            FixedwingVecEnv.rollout_random(envs, 41, spl, use_graph)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            FixedwingVecEnv.rollout_random(envs, n_launch, spl, use_graph)
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / n_launch
            print(f"replicas {replicas:2d} spl {spl} graph {int(use_graph)}: {us:7.2f} us/launch  {us / spl:6.2f} us/env-step  {N * spl / us * 1e-3:6.3f}e9 env-steps/s")
    for e in envs:
        e.close()
