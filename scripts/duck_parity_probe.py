"""Debug probe (GPU): per-(env, step) single-step parity of the duck-only task against the oracle, listing outliers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyflyt_drone_b200 as fw
from oracle import fw_oracle as fo
from pyflyt_drone_b200.vec_env import FixedwingVecEnv

def run(name, **kw):
    cfg = fw.make_config("objlock_duck", noise_ratio=0.0, num_obstacles=8, **kw)
    N = 512
    env, orc = FixedwingVecEnv(N, config=cfg, seed=11), fo.OracleVecEnv(cfg.as_dict(), N, seed=11)
    env.reset(); orc.reset()
    rng = np.random.default_rng(5)
    out = []
    for k in range(60):
        st = orc.get_state()
        env.set_state(st)
        a = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
        og, rg, fg, tg = env.step_arrays(a)
        og = og.copy(); fg = fg.copy().astype(np.int32)
        oc, rc, fc, tc = orc.step(a.astype(np.float64))
        ok = fg == fc
        d = np.abs(og[:, 0:3] - oc[:, 0:3]).max(axis=1) / np.maximum(np.abs(oc[:, 0:3]).max(axis=1), 1.0)
        d[~ok] = 0
        for i in np.nonzero(d > 1e-4)[0]:
            R = st["quat"][i]
            spd = np.linalg.norm(st["vel"][i] - st["wind"][i][:3])
            out.append((k, int(i), float(d[i]), float(np.abs(st["omega"][i]).max()), float(spd), float(st["pos"][i][2]),
                        float(np.abs(oc[i, 0:3]).max())))
    print(f"[{name}] outliers {len(out)} / {N * 60}")
    for o in out[:12]:
        print("   step %d env %d relerr %.2e |omega_in| %.1f airspeed~ %.1f z %.1f |angvel_out| %.1f" % o)
    env.close()

run("default wind +-10")
run("wind off", wind={"enabled": False})
