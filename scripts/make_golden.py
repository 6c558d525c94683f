"""Generate tests/golden/oracle_traj_v1.npz -- REGRESSION fixtures written by this repo's own fp64 oracle.

These are NOT reference-generated vectors (PyFlyt/pybullet cannot be imported offline; the oracle stays "parity
unpinned").  They freeze today's oracle output so that (a) an accidental change of the oracle
shows up in the CPU suite and (b) the GPU suite has a box-independent target: per-step injected-state parity against
stored obs/reward/flags, with scripted (stored) actions.

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pyflyt_drone_b200 as fw  # noqa: E402
from oracle import fw_oracle  # noqa: E402

N, T, SEED = 24, 45, 11
STATE_KEYS = ("pos", "quat", "vel", "omega", "act", "targets", "target_idx", "step_count", "physics_steps", "episode",
              "new_dist", "wind")


def scripted_actions(n, t):
    rng = np.random.default_rng(1234)
    a = rng.uniform(-1, 1, (t, n, 4)).astype(np.float32)
    a[:, : n // 3, 3] = 1.0                      # a third of the fleet at full throttle
    a[:, n // 3: 2 * n // 3, 1] = -1.0           # a third diving: ground contact inside the horizon
    return a


def main():
    fw_oracle.build()
    out = {}
    for name, cfg in (("sparse_euler", fw.waypoints_v3(noise_ratio=0.0)),
                      ("dense_quat", fw.waypoints_v3(noise_ratio=0.0, sparse_reward=0, angle_repr=1, goal_reach=30.0))):
        orc = fw_oracle.OracleVecEnv(cfg.as_dict(), N, seed=SEED, env_id0=5)
        obs0 = orc.reset()
        acts = scripted_actions(N, T)
        obs = np.zeros((T,) + obs0.shape)
        rew = np.zeros((T, N))
        flg = np.zeros((T, N), np.uint8)
        pre = {k: [] for k in STATE_KEYS}
        for t in range(T):
            st = orc.get_state()
            for k in STATE_KEYS:
                pre[k].append(np.array(st[k]))
            o, r, f, _ = orc.step(acts[t].astype(np.float64))
            obs[t], rew[t], flg[t] = o, r, f
        out[f"{name}/obs0"] = obs0
        out[f"{name}/actions"] = acts
        out[f"{name}/obs"] = obs
        out[f"{name}/rew"] = rew
        out[f"{name}/flags"] = flg
        for k in STATE_KEYS:
            out[f"{name}/pre/{k}"] = np.stack(pre[k])
    path = os.path.join(ROOT, "tests", "golden", "oracle_traj_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.endswith("/obs")})
    print("flags seen:", {n: np.unique(out[f"{n}/flags"]).tolist() for n in ("sparse_euler", "dense_quat")})


if __name__ == "__main__":
    main()
