"""BASELINE configs[4]: env-step scaling sweep, 2^10 .. 2^20 envs per GPU, physics-only and Waypoints kernels.
Prints one JSON line per (workload, N) with env-steps/s (whole job when launched under torchrun)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.vec_env import FixedwingVecEnv

rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
L2 = 126 * 2**20
for workload in ("physics_only", "waypoints_v3"):
    cfg = fw.make_config(workload)
    for lg in range(10, 21):
        N = 1 << lg
        per_env = 6 * 16 + 4 + cfg.num_targets * 12
        replicas = max(2, min(64, int(np.ceil(2 * L2 / (N * per_env)))))
        envs = [FixedwingVecEnv(N, config=cfg, device=local_rank, seed=1, env_id0=(rank * replicas + r) * N) for r in range(replicas)]
        for e in envs:
            e.reset_tensor() if cfg.task else None
        K = max(replicas * 4, min(2000, (1 << 27) // N))
        K = (K // replicas) * replicas
        FixedwingVecEnv.rollout_random(envs, replicas * 2, 1, True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        FixedwingVecEnv.rollout_random(envs, K, 1, True)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if rank == 0:
            print(json.dumps({"workload": workload, "envs_per_gpu": N, "n_gpus": world, "launches": K, "replicas": replicas,
                              "us_per_launch": ms * 1e3 / K, "env_steps_per_sec": N * K * world / (ms * 1e-3)}), flush=True)
        for e in envs:
            e.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
