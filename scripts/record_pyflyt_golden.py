#!/usr/bin/env python
"""Record golden trajectories from the REAL PyFlyt / pybullet fixed-wing -- the step that turns "parity unpinned" into pinned.

Run on any machine where ``import PyFlyt, pybullet, gymnasium`` works (they are not installable in the build image):

    python scripts/record_pyflyt_golden.py            # writes tests/golden/pyflyt_*.npz
    python -m pytest tests/test_upstream_parity.py    # oracle <= 1e-6, CUDA <= 1e-4, flags exact; prints the resolved conventions

Nothing here is imported by the package, the tests only read the .npz files it writes.  What is recorded, following the
reference's own call sites (/root/reference/envs/fixedwing_envs/fixedwing_base_env.py:230-257,263-290,314-348,
/root/reference/envs/fixedwing_envs/fixedwing_lowlevel_env.py:64-141, /root/reference/train/train_Fixedwing_Waypoints_v3.py:100-120):

``pyflyt_body_v1.npz``       what pybullet loaded from PyFlyt's fixedwing.urdf: per link mass, local inertia diagonal,
                             inertial frame, CoM relative to the base CoM, joint parent frames, collision shapes; PyFlyt's
                             drone attributes that fix conventions (control/physics ratios, starting velocity, camera).
``pyflyt_mode_m1_v1.npz``    scenario A: ``Aviary`` driven directly in flight mode -1 (six actuator channels, so no sign
                             convention is involved) for 120 control steps = 240 substeps, motor noise 0, fixed scripted
                             commands; raw base pose / velocities after EVERY ``stepSimulation`` and ``state(0)`` /
                             ``aux_state(0)`` after every ``Aviary.step``.
``pyflyt_waypoints_v1.npz``  scenario B: ``gymnasium.make("PyFlyt/Fixedwing-Waypoints-v3")`` + ``FlattenWaypointEnv`` configured
                             as the reference's ``make_env`` (euler, 8 targets, goal 4 m, sparse reward, dome 100, 120 s,
                             context 2), ``reset(seed=0)``, motor noise 0, scripted 4-channel actions for 30 agent steps
                             = 240 substeps (1 s): observation, reward, terminated, truncated, info per step, the waypoint
                             list, and the same raw per-substep states.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCHEMA = 1
# link order recalled from PyFlyt's fixedwing.urdf; the names actually found are stored beside them
LINK_NAMES = {-1: "base_link", 0: "motor_link", 1: "horizontal_tail_link", 2: "vertical_tail_link",
              3: "left_wing_flapped_link", 4: "right_wing_flapped_link", 5: "main_wing_link"}


def scripted_cmd6(n: int) -> np.ndarray:
    """Six-channel commands [left ail, right ail, h-tail, v-tail, main wing, thrust]: smooth, all channels exercised, a
    stall excursion on the tail in the second half."""
    t = np.arange(n) / 120.0
    c = np.zeros((n, 6))
    c[:, 0] = 0.6 * np.sin(2 * np.pi * 0.7 * t)
    c[:, 1] = -0.6 * np.sin(2 * np.pi * 0.7 * t + 0.4)
    c[:, 2] = 0.5 * np.sin(2 * np.pi * 0.5 * t) + np.where(t > 0.5, 0.5, 0.0)
    c[:, 3] = 0.4 * np.cos(2 * np.pi * 0.9 * t)
    c[:, 4] = 0.0
    c[:, 5] = 0.5 + 0.5 * np.sin(2 * np.pi * 0.3 * t) ** 2
    return np.clip(c, -1.0, 1.0)


def scripted_act4(n: int) -> np.ndarray:
    """Four-channel agent actions [roll, pitch, yaw, thrust] in [-1, 1], one per 30 Hz agent step."""
    t = np.arange(n) / 30.0
    a = np.zeros((n, 4))
    a[:, 0] = 0.8 * np.sin(2 * np.pi * 0.6 * t)
    a[:, 1] = 0.5 * np.cos(2 * np.pi * 0.4 * t) - 0.2
    a[:, 2] = 0.6 * np.sin(2 * np.pi * 0.8 * t + 1.0)
    a[:, 3] = 0.3 + 0.6 * np.cos(2 * np.pi * 0.25 * t)
    return np.clip(a, -1.0, 1.0)


class SubstepLog:
    """Wraps ``aviary.stepSimulation`` so that the raw multibody state is captured after every 240 Hz physics step."""

    def __init__(self, aviary, body_id: int):
        self.av, self.id = aviary, body_id
        self.pos, self.quat, self.vel, self.omega = [], [], [], []
        self._orig = aviary.stepSimulation
        aviary.stepSimulation = self._step
        self.snap()

    def snap(self) -> None:
        pos, orn = self.av.getBasePositionAndOrientation(self.id)      # base INERTIAL frame = point O of DESIGN.md
        lin, ang = self.av.getBaseVelocity(self.id)                     # world frame
        self.pos.append(pos); self.quat.append(orn); self.vel.append(lin); self.omega.append(ang)

    def _step(self, *a, **k):
        r = self._orig(*a, **k)
        self.snap()
        return r

    def arrays(self) -> dict:
        return {"sub_pos": np.array(self.pos), "sub_quat": np.array(self.quat), "sub_vel": np.array(self.vel),
                "sub_omega": np.array(self.omega)}


def silence_motor_noise(drone) -> None:
    m = drone.motors
    m.noise_ratio = np.zeros_like(np.asarray(m.noise_ratio, dtype=np.float64))


def dump_body(av, drone) -> dict:
    """Everything the composite rigid body of DESIGN.md section 2 needs, as pybullet sees the loaded URDF."""
    bid = drone.Id
    base_pos, base_orn = av.getBasePositionAndOrientation(bid)
    Rb = np.array(av.getMatrixFromQuaternion(base_orn)).reshape(3, 3)
    links = []
    n_joints = av.getNumJoints(bid)
    for li in range(-1, n_joints):
        dyn = av.getDynamicsInfo(bid, li)
        mass, inertia_diag, inertial_pos, inertial_orn = dyn[0], dyn[2], dyn[3], dyn[4]
        if li == -1:
            com_w, orn_w, up_name = np.array(base_pos), base_orn, "base"
            parent, pf_pos, pf_orn = -2, (0, 0, 0), (0, 0, 0, 1)
        else:
            st = av.getLinkState(bid, li, computeForwardKinematics=1)
            com_w, orn_w = np.array(st[0]), st[1]                         # link CoM (inertial frame) in the world
            ji = av.getJointInfo(bid, li)
            up_name, pf_pos, pf_orn, parent = ji[12].decode(), ji[14], ji[15], ji[16]
        q_rel = av.getDifferenceQuaternion(base_orn, orn_w) if hasattr(av, "getDifferenceQuaternion") else orn_w
        links.append({"index": li, "name": LINK_NAMES.get(li, f"link_{li}"), "upstream_name": up_name, "mass": float(mass),
                      "inertia_diag": [float(x) for x in inertia_diag],
                      "com": [float(x) for x in Rb.T @ (com_w - np.array(base_pos))],
                      "inertial_quat": [float(x) for x in q_rel],
                      "local_inertial_pos": [float(x) for x in inertial_pos],
                      "local_inertial_orn": [float(x) for x in inertial_orn],
                      "parent": int(parent), "parent_frame_pos": [float(x) for x in pf_pos],
                      "parent_frame_orn": [float(x) for x in pf_orn]})
    shapes = []
    for li in range(-1, n_joints):
        for sh in av.getCollisionShapeData(bid, li):
            shapes.append({"link": li, "geom_type": int(sh[2]), "dimensions": [float(x) for x in sh[3]],
                           "local_pos": [float(x) for x in sh[5]], "local_orn": [float(x) for x in sh[6]]})
    attrs = {}
    for k in ("physics_control_ratio", "physics_camera_ratio", "control_period", "physics_period", "starting_velocity"):
        if hasattr(drone, k):
            v = getattr(drone, k)
            attrs[k] = v.tolist() if hasattr(v, "tolist") else v
    for k in ("camera_FOV_degrees", "camera_angle_degrees", "camera_position_offset", "camera_resolution", "is_tracking_camera"):
        cam = getattr(drone, "camera", None)
        if cam is not None and hasattr(cam, k):
            v = getattr(cam, k)
            attrs["camera." + k] = v.tolist() if hasattr(v, "tolist") else v
    return {"links": links, "collision_shapes": shapes, "drone_attrs": attrs,
            "gravity": 9.81, "physics_hz": float(getattr(av, "physics_hz", 240))}


def collision_probe_points(body: dict) -> list:
    """Probe points of the ground-contact test: the corners of every collision box / the extreme points of the other
    shapes, relative to the base CoM in base axes (what params/fixedwing_placeholder.urdf lists by hand)."""
    import pybullet as pb
    by_index = {l["index"]: l for l in body["links"]}
    pts = []
    for sh in body["collision_shapes"]:
        link = by_index[sh["link"]]
        # local_pos is relative to the link's inertial frame
        Rl = np.array(pb.getMatrixFromQuaternion(link["inertial_quat"])).reshape(3, 3)
        Rs = np.array(pb.getMatrixFromQuaternion(sh["local_orn"])).reshape(3, 3)
        d = np.array(sh["dimensions"])
        half = d / 2 if sh["geom_type"] == pb.GEOM_BOX else np.array([d[0], d[0], d[0]])     # sphere / others: radius box
        for sx in (-1, 1):
            for sy in (-1, 1):
                for sz in (-1, 1):
                    local = Rs @ (half * np.array([sx, sy, sz])) + np.array(sh["local_pos"])
                    pts.append((Rl @ local + np.array(link["com"])).tolist())
    return pts


def record_mode_m1(n_control_steps: int = 120) -> tuple[dict, dict]:
    from PyFlyt.core import Aviary
    start_pos, start_orn = np.array([[0.0, 0.0, 10.0]]), np.array([[0.0, 0.0, 0.0]])
    av = Aviary(start_pos=start_pos, start_orn=start_orn, drone_type="fixedwing", render=False)
    av.set_mode(-1)
    drone = av.drones[0]
    silence_motor_noise(drone)
    body = dump_body(av, drone)
    body["collision_points"] = collision_probe_points(body)
    log = SubstepLog(av, drone.Id)
    cmd = scripted_cmd6(n_control_steps)
    state, aux = [np.array(av.state(0))], [np.array(av.aux_state(0))]
    for s in range(n_control_steps):
        av.set_setpoint(0, cmd[s])
        av.step()
        state.append(np.array(av.state(0))); aux.append(np.array(av.aux_state(0)))
    out = {"schema": SCHEMA, "cmd": cmd, "state": np.array(state), "aux": np.array(aux),
           "start_pos": start_pos[0], "physics_steps_per_control": len(log.pos) // n_control_steps, **log.arrays()}
    av.disconnect()
    return out, body


def record_waypoints(n_agent_steps: int = 30) -> dict:
    import gymnasium as gym
    import PyFlyt.gym_envs  # noqa: F401  (registers PyFlyt/Fixedwing-Waypoints-v3)
    from PyFlyt.gym_envs import FlattenWaypointEnv
    # train_Fixedwing_Waypoints_v3.py:27-55,100-117
    env = gym.make("PyFlyt/Fixedwing-Waypoints-v3", render_mode=None, num_targets=8, goal_reach_distance=4.0,
                   flight_dome_size=100.0, max_duration_seconds=120.0, angle_representation="euler", agent_hz=30,
                   sparse_reward=True)
    env = FlattenWaypointEnv(env, context_length=2)
    obs0, _ = env.reset(seed=0)
    base = env.unwrapped
    av, drone = base.env, base.env.drones[0]
    silence_motor_noise(drone)            # the 20 warm-up substeps ran at zero throttle: the noise term was 0 there anyway
    log = SubstepLog(av, drone.Id)
    targets = np.array(base.waypoints.targets)
    act = scripted_act4(n_agent_steps)
    obs, rew, term, trunc, reached, coll, oob = [np.array(obs0)], [], [], [], [], [], []
    for t in range(n_agent_steps):
        o, r, te, tr, info = env.step(act[t])
        obs.append(np.array(o)); rew.append(float(r)); term.append(bool(te)); trunc.append(bool(tr))
        reached.append(int(info.get("num_targets_reached", 0))); coll.append(bool(info.get("collision", False)))
        oob.append(bool(info.get("out_of_bounds", False)))
        if te or tr:
            break
    out = {"schema": SCHEMA, "actions": act[: len(rew)], "obs": np.array(obs), "rew": np.array(rew), "term": np.array(term),
           "trunc": np.array(trunc), "num_targets_reached": np.array(reached), "collision": np.array(coll),
           "out_of_bounds": np.array(oob), "targets": targets, "start_pos": np.array([0.0, 0.0, 10.0]),
           "warmup_physics_steps": int(av.physics_steps) - len(log.pos) + 1 if hasattr(av, "physics_steps") else 20,
           **log.arrays()}
    env.close()
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=GOLDEN)
    args = ap.parse_args()
    try:
        import PyFlyt, pybullet  # noqa: F401,E401
    except ImportError as e:
        print(f"record_pyflyt_golden: {e}.  This script needs the real PyFlyt + pybullet (+ gymnasium); run it on a machine "
              "that has them and commit the three tests/golden/pyflyt_*.npz files it writes.", file=sys.stderr)
        return 2
    os.makedirs(args.out, exist_ok=True)
    versions = {"PyFlyt": getattr(PyFlyt, "__version__", "?"), "numpy": np.__version__}
    m1, body = record_mode_m1()
    np.savez(os.path.join(args.out, "pyflyt_body_v1.npz"), schema=SCHEMA, body_json=json.dumps(body), versions=json.dumps(versions))
    np.savez(os.path.join(args.out, "pyflyt_mode_m1_v1.npz"), **m1)
    wp = record_waypoints()
    np.savez(os.path.join(args.out, "pyflyt_waypoints_v1.npz"), **wp)
    total = sum(l["mass"] for l in body["links"])
    print(f"recorded: body ({len(body['links'])} links, {total:.4f} kg), mode -1 ({len(m1['cmd'])} control steps, "
          f"{len(m1['sub_pos']) - 1} substeps), waypoints ({len(wp['rew'])} agent steps, {len(wp['sub_pos']) - 1} substeps) -> {args.out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
