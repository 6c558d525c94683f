"""Generate tests/golden/oracle_traj_v1.npz -- REGRESSION fixtures written by this repo's own fp64 oracle.

These are NOT reference-generated vectors (PyFlyt/pybullet cannot be imported offline; the oracle stays "parity
unpinned").  They freeze today's oracle output so that (a) an accidental change of the oracle
shows up in the CPU suite and (b) the GPU suite has a box-independent target: per-step injected-state parity against
stored obs/reward/flags, with scripted (stored) actions.

    python scripts/make_golden.py            # both files
    python scripts/make_golden.py --heads    # only oracle_traj_v2_heads.npz (low-level and duck-only task heads)
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


HEAD_KEYS = {"lowlevel": STATE_KEYS,
             "objlock_duck": STATE_KEYS + ("duck", "obst", "ol_f", "ol_i", "vis_hist")}


def head_actions(name, n, t):
    rng = np.random.default_rng(4321)
    if name == "lowlevel":
        a = rng.uniform(-1, 1, (t, n, 6)).astype(np.float32)
        a[:, : n // 3, 2] = 1.0                  # a third with the elevator hard over: altitude-band terminations
        return a
    a = rng.uniform(-1, 1, (t, n, 4)).astype(np.float32)
    a[:, : n // 2, :3] *= 0.1                    # the approach half flies (nearly) straight at its duck
    a[:, : n // 2, 3] = 0.0
    return a


def main_heads():
    """tests/golden/oracle_traj_v2_heads.npz: the same kind of fixture for the two task heads of SURVEY 8 f3 --
    the low-level tracking env and the duck-only lock/strike env (half of the fleet starts low, 25-45 m short of its
    duck and heading at it, so that frames with the duck in view, locks and strikes fall inside the horizon)."""
    fw_oracle.build()
    out = {}
    for name in ("lowlevel", "objlock_duck"):
        cfg = fw.make_config(name, noise_ratio=0.0)
        orc = fw_oracle.OracleVecEnv(cfg.as_dict(), N, seed=SEED, env_id0=5)
        obs0 = orc.reset()
        if name == "objlock_duck":
            st = orc.get_state()
            rng = np.random.default_rng(77)
            h = N // 2
            ang, dist = rng.uniform(-np.pi, np.pi, h), rng.uniform(25, 45, h)
            st["pos"][:h, 0] = st["duck"][:h, 0] - dist * np.cos(ang)
            st["pos"][:h, 1] = st["duck"][:h, 1] - dist * np.sin(ang)
            st["pos"][:h, 2] = rng.uniform(4, 8, h)
            st["quat"][:h] = np.stack([0 * ang, 0 * ang, np.sin(ang / 2), np.cos(ang / 2)], 1)
            st["vel"][:h] = np.stack([20 * np.cos(ang), 20 * np.sin(ang), 0 * ang], 1)
            st["omega"][:h] = 0
            orc.set_state(st)
        else:
            # 0.375 s of flight cannot leave the 1..100 m altitude band from the 10 m start: put a few aircraft at its
            # edges and a few near the 2,000-step limit so that both flag kinds appear
            st = orc.get_state()
            st["pos"][:3, 2] = 1.15; st["vel"][:3, 2] = -3.0
            st["pos"][3:5, 2] = 99.9; st["vel"][3:5, 2] = 4.0
            st["step_count"][5:8] = [1960, 1975, 1990]
            orc.set_state(st)
        acts = head_actions(name, N, T)
        obs = np.zeros((T,) + obs0.shape)
        rew = np.zeros((T, N))
        flg = np.zeros((T, N), np.uint8)
        pre = {k: [] for k in HEAD_KEYS[name]}
        for t in range(T):
            st = orc.get_state()
            for k in HEAD_KEYS[name]:
                pre[k].append(np.array(st[k]))
            o, r, f, _ = orc.step(acts[t].astype(np.float64))
            obs[t], rew[t], flg[t] = o, r, f
        out[f"{name}/obs0"] = obs0
        out[f"{name}/actions"] = acts
        out[f"{name}/obs"] = obs
        out[f"{name}/rew"] = rew
        out[f"{name}/flags"] = flg
        for k in HEAD_KEYS[name]:
            out[f"{name}/pre/{k}"] = np.stack(pre[k])
    path = os.path.join(ROOT, "tests", "golden", "oracle_traj_v2_heads.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.endswith("/obs")})
    print("flags seen:", {n: np.unique(out[f"{n}/flags"]).tolist() for n in ("lowlevel", "objlock_duck")})


if __name__ == "__main__":
    if "--heads" in sys.argv:
        main_heads()
    else:
        main()
        main_heads()
