import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import pyflyt_drone_b200 as p
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
from oracle import fw_oracle as fo
N=256
cfg = p.waypoints_v3(noise_ratio=0.0)
env = FixedwingVecEnv(N, config=cfg, seed=7)
orc = fo.OracleVecEnv(cfg.as_dict(), N, seed=7)
o_g = env.reset(); o_c = orc.reset()
print("reset obs maxabs diff", np.abs(o_g-o_c).max(), "rel", (np.abs(o_g-o_c)/(np.abs(o_c)+1e-3)).max())
rng = np.random.default_rng(0)
maxrel=0
for k in range(60):
    a = rng.uniform(-1,1,(N,4)).astype(np.float32)
    og, rg, dg, infos = env.step(a)
    oc, rc, fc, tc = orc.step(a.astype(np.float64))
    fg = env._h_flags
    d = np.abs(og-oc); rel = (d/(np.abs(oc)+1e-2)).max()
    if k%10==0 or not np.array_equal(fg.astype(np.int32), fc): print(k, "obs maxabs", d.max(), "rel", rel, "rew", np.abs(rg-rc).max(), "flags eq", np.array_equal(fg.astype(np.int32), fc), int(dg.sum()))
print("stats", env.episode_stats())
# throughput quick
for preset, nn in (("physics_only", 65536), ("waypoints_v3", 65536)):
    e2 = FixedwingVecEnv(nn, preset=preset, seed=1)
    e2.step_random(0, 20); torch.cuda.synchronize()
    t0=torch.cuda.Event(enable_timing=True); t1=torch.cuda.Event(enable_timing=True)
    t0.record(); e2.step_random(100, 200); t1.record(); torch.cuda.synchronize()
    ms=t0.elapsed_time(t1)/200
    print(preset, nn, "ms/step", ms, "env-steps/s %.3e"%(nn/ms*1e3), e2.episode_stats())
    e2.close()
