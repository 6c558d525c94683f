"""Probe (GPU): how much of the ObjLock random-action step is the in-kernel reset (20 warm-up substeps per finished env,
executed by the whole warp)?  Times the step with the reference warm-up and with warmup_inner=0, and reports the
episode rate."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyflyt_drone_b200 as fw
from pyflyt_drone_b200.vec_env import FixedwingVecEnv

def run(name, preset, **kw):
    N, K = 65536, 200
    envs = [FixedwingVecEnv(N, config=fw.make_config(preset, **kw), seed=1, env_id0=r * N) for r in range(6)]
    FixedwingVecEnv.rollout_random(envs, 120, 1, use_graph=False)
    for e in envs: e.episode_stats()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    FixedwingVecEnv.rollout_random(envs, K, 1, use_graph=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    eps = sum(e.episode_stats()["episodes"] for e in envs)
    p = eps / (N * K)
    print(f"[{name}] {ms * 1e3:.1f} us/step, {N / ms / 1e3:.3e} env-steps/s, done per env-step {p:.4f}, "
          f"P(warp has a reset) {1 - (1 - p) ** 32:.2f}")
    for e in envs: e.close()

run("objlock", "waypoint_objlock")
run("objlock, no warm-up", "waypoint_objlock", warmup_inner=0)
run("duck", "objlock_duck")
run("duck, no warm-up", "objlock_duck", warmup_inner=0)
