import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import pyflyt_drone_b200 as p
from pyflyt_drone_b200.vec_env import FixedwingVecEnv
from oracle import fw_oracle as fo
np.set_printoptions(precision=3, linewidth=200, suppress=False)
N=256
for fast in (0,1):
    cfg = p.waypoints_v3(noise_ratio=0.0, fast_trig=fast)
    env = FixedwingVecEnv(N, config=cfg, seed=7)
    orc = fo.OracleVecEnv(cfg.as_dict(), N, seed=7)
    o_g = env.reset(); o_c = orc.reset()
    print("fast", fast, "reset obs maxabs diff per col", np.abs(o_g-o_c).max(0))
    rng = np.random.default_rng(0)
    for k in range(31):
        a = rng.uniform(-1,1,(N,4)).astype(np.float32)
        og, rg, dg, infos = env.step(a)
        oc, rc, fc, tc = orc.step(a.astype(np.float64))
        fg = env._h_flags
        d = np.abs(og-oc)
        if k in (0,1,5,15,30) or not np.array_equal(fg.astype(np.int32), fc):
            print(k, "free-run maxabs per col", d.max(0), "rew", np.abs(rg-rc).max(), "flags eq", np.array_equal(fg.astype(np.int32), fc), int(dg.sum()))
    # single-step parity with injected state
    worst = 0
    for k in range(20):
        st = orc.get_state()
        env.set_state(st)
        a = rng.uniform(-1,1,(N,4)).astype(np.float32)
        og, rg, dg, infos = env.step(a)
        oc, rc, fc, tc = orc.step(a.astype(np.float64))
        fg = env._h_flags.astype(np.int32)
        done = (fc & 3) != 0
        rel = np.abs(og-oc)/np.maximum(np.abs(oc), 1e-2)
        rel[done] = 0   # reset obs compared separately
        worst = max(worst, rel.max())
        if not np.array_equal(fg, fc) or np.abs(rg-rc).max() > 1e-4: print("MISMATCH step", k, np.nonzero(fg!=fc), np.abs(rg-rc).max())
    print("fast", fast, "single-step worst rel err (floor 1e-2)", worst)
    env.close()
# throughput quick
for preset, nn in (("physics_only", 65536), ("waypoints_v3", 65536), ("physics_only", 1<<20)):
    for fast in (0,1):
        e2 = FixedwingVecEnv(nn, preset=preset, seed=1, fast_trig=fast)
        e2.step_random(0, 20); torch.cuda.synchronize()
        t0=torch.cuda.Event(enable_timing=True); t1=torch.cuda.Event(enable_timing=True)
        t0.record(); e2.step_random(100, 200); t1.record(); torch.cuda.synchronize()
        ms=t0.elapsed_time(t1)/200
        print(preset, nn, "fast", fast, "ms/step %.4f"%ms, "env-steps/s %.3e"%(nn/ms*1e3))
        e2.close()
