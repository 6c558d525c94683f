"""Replay of recorded PyFlyt trajectories (scripts/record_pyflyt_golden.py) through the oracle or the CUDA path.

Test infrastructure.  The recordings carry (a) what pybullet loaded from PyFlyt's fixedwing.urdf, from which the
aircraft configuration is rebuilt (the placeholder URDF of the package plays no part), and (b) two trajectories:
scenario A drives Aviary in flight mode -1 (six actuator channels, independent of every sign convention), scenario B is
the Fixedwing-Waypoints-v3 env through FlattenWaypointEnv with four-channel actions.  ``resolve_conventions`` searches the
[UP-RECALL] convention flags of SURVEY.md appendix B for the combination that reproduces the recording.
"""
from __future__ import annotations

import itertools
import json
import os

import numpy as np

import pyflyt_drone_b200 as fw
from pyflyt_drone_b200 import aircraft

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = ("pyflyt_body_v1.npz", "pyflyt_mode_m1_v1.npz", "pyflyt_waypoints_v1.npz")
# appendix B flags: every combination is tried, cheapest first (mode -1 only depends on the first two)
AERO_FLAGS = {"freestream_3d": (1, 0), "cd90_degrees": (1, 0)}
SIGN_FLAGS = {"ail": ((1.0, -1.0), (-1.0, 1.0)), "pitch_sign": (1.0, -1.0), "yaw_sign": (1.0, -1.0)}


def have_recording(directory: str = GOLDEN) -> bool:
    return all(os.path.exists(os.path.join(directory, f)) for f in FILES)


def load_recording(directory: str = GOLDEN) -> dict:
    body = json.loads(str(np.load(os.path.join(directory, FILES[0]))["body_json"]))
    m1 = dict(np.load(os.path.join(directory, FILES[1])))
    wp = dict(np.load(os.path.join(directory, FILES[2])))
    return {"body": body, "mode_m1": m1, "waypoints": wp}


def config_from_body(body: dict, preset: str, **over):
    rb = aircraft.body_from_bullet(body["links"], body.get("collision_points"))
    cfg = fw.make_config(preset, noise_ratio=0.0, **over).with_aircraft(body=rb)
    return cfg.replace(noise_ratio=0.0, **over)


def _state_at(rec: dict, k: int, act6: np.ndarray, physics_steps: int) -> dict:
    return {"pos": rec["sub_pos"][k][None], "quat": rec["sub_quat"][k][None], "vel": rec["sub_vel"][k][None],
            "omega": rec["sub_omega"][k][None], "act": act6[None], "physics_steps": np.array([physics_steps], np.int32),
            "step_count": np.array([0], np.int32), "episode": np.array([0], np.uint32)}


def replay_mode_m1(make_env, rec: dict, cfg, steps: int | None = None):
    """Scenario A through ``make_env(cfg)`` (an object with reset/set_state/get_state/step in the OracleVecEnv shape).
    Returns the worst relative state error after each control step: dict of per-field arrays."""
    env = make_env(cfg)
    env.reset()
    env.set_state(_state_at(rec, 0, np.asarray(rec["aux"][0], np.float64), 0))
    per = int(rec["physics_steps_per_control"])
    n = len(rec["cmd"]) if steps is None else min(steps, len(rec["cmd"]))
    err = {k: np.zeros(n) for k in ("pos", "quat", "vel", "omega", "aux")}
    for s in range(n):
        env.step(np.asarray(rec["cmd"][s], np.float64)[None])
        st = env.get_state()
        k = per * (s + 1)
        q = st["quat"][0] * np.sign(np.dot(st["quat"][0], rec["sub_quat"][k]) or 1.0)        # q and -q are the same attitude
        for name, got, ref in (("pos", st["pos"][0], rec["sub_pos"][k]), ("quat", q, rec["sub_quat"][k]),
                               ("vel", st["vel"][0], rec["sub_vel"][k]), ("omega", st["omega"][0], rec["sub_omega"][k]),
                               ("aux", st["act"][0], rec["aux"][s + 1])):
            err[name][s] = np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1.0)
    return err


def replay_waypoints(make_env, rec: dict, cfg, steps: int | None = None):
    """Scenario B: recorded waypoints injected, recorded actions applied; returns (obs error per step, reward error per
    step, flags equal per step)."""
    env = make_env(cfg)
    env.reset()
    T = cfg.num_targets
    warm = int(rec.get("warmup_physics_steps", 20))
    st = _state_at(rec, 0, np.zeros(6), warm)
    st["targets"] = np.asarray(rec["targets"], np.float64)[None, :T]
    st["target_idx"] = np.array([0], np.int32)
    st["new_dist"] = np.array([np.linalg.norm(rec["targets"][0] - rec["sub_pos"][0])])
    env.set_state(st)
    n = len(rec["rew"]) if steps is None else min(steps, len(rec["rew"]))
    oerr, rerr, fl = np.zeros(n), np.zeros(n), np.zeros(n, bool)
    for t in range(n):
        obs, rew, flags, term_obs = env.step(np.asarray(rec["actions"][t], np.float64)[None])
        done = bool(int(flags[0]) & 3)
        o = np.asarray(term_obs[0] if done else obs[0], np.float64)
        ref = np.asarray(rec["obs"][t + 1], np.float64)
        d = o - ref
        d[3:6] = (d[3:6] + np.pi) % (2 * np.pi) - np.pi                                      # euler angles wrap
        oerr[t] = np.abs(d).max() / max(np.abs(ref).max(), 1.0)
        rerr[t] = abs(float(rew[0]) - float(rec["rew"][t]))
        fl[t] = (bool(int(flags[0]) & 1) == bool(rec["term"][t])) and (bool(int(flags[0]) & 2) == bool(rec["trunc"][t]))
        if done:
            break
    return oerr, rerr, fl


def resolve_conventions(make_env, rec: dict, verbose: bool = True) -> dict:
    """Try every combination of the appendix-B convention flags; keep the one with the smallest error over the first
    control / agent steps.  Aero flags come from scenario A (mode -1 needs no signs), signs from scenario B.  A flag whose
    alternatives reproduce the recording equally well (e.g. ``cd90_degrees`` when no surface stalls inside the horizon) is
    listed under ``"ambiguous"`` and keeps the package default."""
    scores_a = {}
    for f3, cd in itertools.product(*AERO_FLAGS.values()):
        cfg = config_from_body(rec["body"], "lowlevel", freestream_3d=f3, cd90_degrees=cd)
        e = replay_mode_m1(make_env, rec["mode_m1"], cfg, steps=60)
        scores_a[(f3, cd)] = max(v.max() for v in e.values())
    (f3, cd), best_a = min(scores_a.items(), key=lambda kv: kv[1])
    flags = {"freestream_3d": f3, "cd90_degrees": cd}
    scores_b = {}
    for (al, ar), ps, ys in itertools.product(*SIGN_FLAGS.values()):
        cfg = config_from_body(rec["body"], "waypoints_v3", ail_left_sign=al, ail_right_sign=ar, pitch_sign=ps, yaw_sign=ys, **flags)
        oerr, _, _ = replay_waypoints(make_env, rec["waypoints"], cfg, steps=6)
        scores_b[(al, ar, ps, ys)] = oerr.max()
    (al, ar, ps, ys), best_b = min(scores_b.items(), key=lambda kv: kv[1])
    out = {**flags, "ail_left_sign": al, "ail_right_sign": ar, "pitch_sign": ps, "yaw_sign": ys}
    tie = lambda a, b: abs(a - b) <= 1e-12 + 1e-3 * min(a, b)      # noqa: E731
    amb = []
    if tie(scores_a[(1 - f3, cd)], best_a): amb.append("freestream_3d")
    if tie(scores_a[(f3, 1 - cd)], best_a): amb.append("cd90_degrees")
    if tie(scores_b[(-al, -ar, ps, ys)], best_b): amb.extend(["ail_left_sign", "ail_right_sign"])
    if tie(scores_b[(al, ar, -ps, ys)], best_b): amb.append("pitch_sign")
    if tie(scores_b[(al, ar, ps, -ys)], best_b): amb.append("yaw_sign")
    out["ambiguous"] = amb
    if verbose:
        print(f"\n[upstream parity] resolved conventions: {out} (scenario A error {best_a:.2e}, scenario B error {best_b:.2e})")
    return out
