"""ctypes front-end of the fp64 CPU oracle (oracle/fw_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see fw_oracle.h).  May be imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product package.
The oracle takes a plain dict (``EnvConfig.as_dict()``) so that it shares no code with the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfw_oracle.so")

NSURF, MAX_TARGETS, MAX_COL, MAX_OBST, MAX_OBS, MAX_HIST = 5, 16, 16, 32, 72, 4
_D, _I, _U = C.c_double, C.c_int32, C.c_uint32
_V3 = _D * 3


class OConfig(C.Structure):
    _fields_ = [
        ("cl_alpha_2d", _D * NSURF), ("chord", _D * NSURF), ("span", _D * NSURF), ("flap_to_chord", _D * NSURF),
        ("eta", _D * NSURF), ("alpha0_base_deg", _D * NSURF), ("stall_p_base_deg", _D * NSURF),
        ("stall_n_base_deg", _D * NSURF), ("cd0", _D * NSURF), ("defl_limit_deg", _D * NSURF), ("surf_tau", _D * NSURF),
        ("lift_unit", _V3 * NSURF), ("fwd_unit", _V3 * NSURF), ("r_surf", _V3 * NSURF),
        ("total_thrust", _D), ("thrust_coef", _D), ("torque_coef", _D), ("noise_ratio", _D), ("motor_tau", _D),
        ("r_motor", _V3), ("thrust_unit", _V3),
        ("mass", _D), ("com", _V3), ("inertia_o", _D * 9),
        ("n_col", _I), ("_pad0", _I),
        ("col_pts", _V3 * MAX_COL), ("contact_margin", _D),
        ("dt", _D), ("gravity", _D), ("rho", _D), ("max_coord_vel", _D),
        ("physics_per_control", _I), ("substeps_per_inner", _I), ("inner_per_step", _I), ("warmup_inner", _I),
        ("ail_left_sign", _D), ("ail_right_sign", _D), ("pitch_sign", _D), ("yaw_sign", _D),
        ("freestream_3d", _I), ("cd90_degrees", _I),
        ("task", _I), ("num_targets", _I), ("goal_reach", _D), ("sparse_reward", _I), ("angle_repr", _I),
        ("dome", _D), ("max_steps", _I), ("context_len", _I),
        ("start_pos", _V3), ("start_vel", _V3), ("spawn_size", _D), ("min_height", _D),
        ("early_return_on_crash", _I), ("complete_truncates", _I),
        ("wind_mode", _I), ("wind_randomize", _I), ("wind_rand_phase", _I), ("wind_start_substep", _I),
        ("wind_base", _V3), ("wind_base_lo", _V3), ("wind_base_hi", _V3),
        ("gust_amp", _V3), ("gust_amp_lo", _V3), ("gust_amp_hi", _V3), ("gust_freq", _D), ("gust_phase", _D),
        ("num_obstacles", _I), ("cam_interval_substeps", _I),
        ("obst_radius", _D), ("obst_h_lo", _D), ("obst_h_hi", _D), ("obst_safe", _D), ("obst_scale", _D),
        ("obst_max_pen", _D),
        ("lock_hold_steps", _I), ("switch_min_seen", _I),
        ("strike_dist", _D), ("strike_reward", _D), ("lock_step_reward", _D), ("approach_scale", _D),
        ("switch_min_area", _D), ("duck_radius", _D), ("cam_offset", _V3), ("cam_near", _D), ("cam_far", _D),
        ("cam_res", _I), ("cam_mode", _I),
        ("cam_tilt_deg", _D), ("duck_dist_scale", _D), ("lock_center_radius", _D), ("centering_scale", _D),
        ("visible_step_reward", _D), ("area_reward_scale", _D), ("lock_lost_penalty", _D), ("approach_clip", _D),
        ("vision_hist_len", _I), ("vision_use_deltas", _I), ("lock_decay_steps", _I), ("_pad1", _I),
    ]


class OEnv(C.Structure):
    _fields_ = [
        ("pos", _V3), ("quat", _D * 4), ("vel", _V3), ("omega", _V3), ("act", _D * NSURF), ("throttle", _D),
        ("surf_vel", _V3 * NSURF), ("setpoint", _D * 6), ("cmd", _D * 6), ("last_action", _D * 6), ("target_ref", _V3),
        ("targets", _V3 * MAX_TARGETS), ("n_remaining", _I), ("target_idx", _I), ("old_dist", _D), ("new_dist", _D),
        ("step_count", _I), ("physics_steps", _I), ("termination", _I), ("truncation", _I),
        ("info_collision", _I), ("info_oob", _I), ("info_complete", _I), ("info_strike", _I),
        ("num_targets_reached", _I), ("contact", _I), ("episode", _U), ("env_id", _U), ("reward", _D),
        ("wind_base", _V3), ("gust_amp", _V3), ("gust_phase", _D),
        ("duck_pos", _V3), ("obst", _V3 * MAX_OBST), ("n_obst", _I),
        ("duck_phase", _I), ("seen_consecutive", _I), ("lock_steps", _I), ("has_prev_dist", _I),
        ("post_waypoints", _I), ("steps_since_seen", _I), ("cam_valid", _I),
        ("prev_est_dist", _D), ("last_cx", _D), ("last_cy", _D), ("last_area", _D), ("last_depth", _D),
        ("vision", _D * 9),
        ("frame_visible", _I), ("_pad", _I),
        ("frame_cx", _D), ("frame_cy", _D), ("frame_area", _D), ("frame_depth", _D), ("frame_dl", _D),
        ("frame_dc", _D), ("frame_dr", _D),
        ("ep_return", _D), ("ep_length", _I), ("_pad2", _I),
        ("vis_hist", (_D * 9) * MAX_HIST), ("vis_deltas", _D * 4), ("target_vec", _V3), ("hist_filled", _I), ("_pad3", _I),
    ]


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fw_oracle.c")
    hdr = os.path.join(_HERE, "fw_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        r = subprocess.run(["make", "-C", _HERE, "-B"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.fwo_config_size.restype = C.c_int
        L.fwo_env_size.restype = C.c_int
        L.fwo_obs_dim.restype = C.c_int
        L.fwo_obs_dim.argtypes = [C.POINTER(OConfig)]
        L.fwo_act_dim.restype = C.c_int
        L.fwo_act_dim.argtypes = [C.POINTER(OConfig)]
        L.fwo_random_action6.argtypes = [C.c_uint64, _U, _U, _U, C.POINTER(_D)]
        L.fwo_u01.restype = _D
        L.fwo_u01.argtypes = [_U]
        L.fwo_philox.argtypes = [C.c_uint64, _U, _U, _U, _U, C.POINTER(_U)]
        L.fwo_normals4.argtypes = [C.c_uint64, _U, _U, _U, C.POINTER(_D)]
        L.fwo_random_action.argtypes = [C.c_uint64, _U, _U, _U, C.POINTER(_D)]
        L.fwo_aero_coeffs.argtypes = [C.POINTER(OConfig), C.c_int, _D, _D, C.POINTER(_D)]
        L.fwo_surface_force.argtypes = [C.POINTER(OConfig), C.c_int, _D, C.POINTER(_D), C.POINTER(_D), C.POINTER(_D)]
        L.fwo_substep.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_uint64]
        L.fwo_reset.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_uint64, _U, _U, C.c_void_p]
        L.fwo_step.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fwo_vec_reset.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_int, C.c_uint64, _U, C.c_void_p, C.c_int]
        L.fwo_vec_step.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_int, C.c_uint64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.fwo_vec_step_info.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_int, C.c_uint64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.fwo_rollout_random.restype = C.c_long
        L.fwo_rollout_random.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_int, C.c_uint64, C.c_int, C.c_int]
        L.fwo_render.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fwo_render.restype = None
        L.fwo_compute_obs.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_void_p, C.c_int]
        L.fwo_refresh_surface_vel.argtypes = [C.POINTER(OConfig), C.POINTER(OEnv), C.c_int]
        L.fwo_quat_to_euler.argtypes = [C.POINTER(_D), C.POINTER(_D)]
        L.fwo_euler_to_quat.argtypes = [C.POINTER(_D), C.POINTER(_D)]
        L.fwo_quat_to_mat.argtypes = [C.POINTER(_D), C.POINTER(_D)]
        if L.fwo_config_size() != C.sizeof(OConfig):
            raise RuntimeError(f"fwo_config mismatch: C {L.fwo_config_size()} vs ctypes {C.sizeof(OConfig)}")
        if L.fwo_env_size() != C.sizeof(OEnv):
            raise RuntimeError(f"fwo_env mismatch: C {L.fwo_env_size()} vs ctypes {C.sizeof(OEnv)}")
        _lib = L
    return _lib


def make_config(d: dict) -> OConfig:
    """Build the oracle's config struct from a plain dict (keys as in EnvConfig.as_dict())."""
    c = OConfig()
    for name, _ in OConfig._fields_:
        if name.startswith("_pad"):
            continue
        if name not in d:
            raise KeyError(f"oracle config needs key {name!r}")
        v = d[name]
        if name == "col_pts":
            for i, p in enumerate(v):
                for k in range(3):
                    c.col_pts[i][k] = float(p[k])
        elif isinstance(v, (list, tuple)):
            arr = getattr(c, name)
            if len(v) and isinstance(v[0], (list, tuple)):
                for i, row in enumerate(v):
                    for k, x in enumerate(row):
                        arr[i][k] = float(x)
            else:
                for i, x in enumerate(v):
                    arr[i] = float(x)
        else:
            setattr(c, name, v)
    return c


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleVecEnv:
    """N independent fp64 oracle envs with SubprocVecEnv auto-reset semantics."""

    def __init__(self, cfg: dict, n: int, seed: int = 0, env_id0: int = 0, nthreads: int = 1):
        self.L = lib()
        self.cfg = make_config(cfg)
        self.n, self.seed, self.env_id0, self.nthreads = int(n), int(seed), int(env_id0), int(nthreads)
        self.envs = (OEnv * self.n)()
        self.obs_dim = self.L.fwo_obs_dim(C.byref(self.cfg))
        self.act_dim = self.L.fwo_act_dim(C.byref(self.cfg))
        self.num_targets = int(cfg["num_targets"])

    def reset(self) -> np.ndarray:
        obs = np.zeros((self.n, max(self.obs_dim, 1)), dtype=np.float64)
        self.L.fwo_vec_reset(C.byref(self.cfg), self.envs, self.n, self.seed, self.env_id0, _ptr(obs), self.nthreads)
        return obs[:, : self.obs_dim]

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.n, self.act_dim)
        D = max(self.obs_dim, 1)
        obs = np.zeros((self.n, D)); term = np.zeros((self.n, D))
        rew = np.zeros(self.n); flags = np.zeros(self.n, dtype=np.int32)
        self.last_targets_reached = np.zeros(self.n, dtype=np.int32)      # info["num_targets_reached"], pre-reset
        self.L.fwo_vec_step_info(C.byref(self.cfg), self.envs, self.n, self.seed, _ptr(a), _ptr(obs), _ptr(rew), _ptr(flags),
                                 _ptr(term), _ptr(self.last_targets_reached), self.nthreads)
        return obs[:, : self.obs_dim], rew, flags, term[:, : self.obs_dim]

    def rollout_random(self, steps: int) -> int:
        return int(self.L.fwo_rollout_random(C.byref(self.cfg), self.envs, self.n, self.seed, int(steps), self.nthreads))

    def render(self, env_index: int = 0, width: int | None = None, height: int | None = None) -> dict:
        """Debug frame of one env (fwo_render): rgba uint8 [H,W,4], seg int32 [H,W], depth-buffer values float64 [H,W]."""
        W = int(width or self.cfg.cam_res); H = int(height or self.cfg.cam_res)
        rgba = np.zeros((H, W, 4), np.uint8); seg = np.zeros((H, W), np.int32); depth = np.zeros((H, W), np.float64)
        self.L.fwo_render(C.byref(self.cfg), C.byref(self.envs[int(env_index)]), W, H, _ptr(rgba), _ptr(seg), _ptr(depth))
        return dict(rgba=rgba, seg=seg, depth=depth)

    # ---- state exchange with the device path (same field meaning as FwStateHost) ----
    def get_state(self) -> dict:
        n, T = self.n, self.num_targets
        out = dict(pos=np.zeros((n, 3)), quat=np.zeros((n, 4)), vel=np.zeros((n, 3)), omega=np.zeros((n, 3)),
                   act=np.zeros((n, 6)), targets=np.zeros((n, max(T, 1), 3)), target_idx=np.zeros(n, np.int32),
                   step_count=np.zeros(n, np.int32), physics_steps=np.zeros(n, np.int32),
                   episode=np.zeros(n, np.uint32), new_dist=np.zeros(n), wind=np.zeros((n, 7)))
        for i in range(n):
            e = self.envs[i]
            out["pos"][i] = e.pos[:]; out["quat"][i] = e.quat[:]; out["vel"][i] = e.vel[:]; out["omega"][i] = e.omega[:]
            out["act"][i, :5] = e.act[:]; out["act"][i, 5] = e.throttle
            # original list: already-reached targets are unknown to the oracle -> leave zeros before target_idx
            for t in range(e.n_remaining):
                out["targets"][i, e.target_idx + t] = e.targets[t][:]
            if self.cfg.task == 3:                       # low-level task: the tracked reference rides in slot 0
                out["targets"][i, 0] = e.target_ref[:]
            out["target_idx"][i] = e.target_idx; out["step_count"][i] = e.step_count
            out["physics_steps"][i] = e.physics_steps; out["episode"][i] = e.episode
            out["new_dist"][i] = e.new_dist if np.isfinite(e.new_dist) else 0.0
            out["wind"][i, :3] = e.wind_base[:]; out["wind"][i, 3:6] = e.gust_amp[:]; out["wind"][i, 6] = e.gust_phase
        if self.cfg.task in (2, 4):
            out["duck"] = np.zeros((n, 3)); out["obst"] = np.zeros((n, MAX_OBST, 3))
            out["ol_f"] = np.zeros((n, 12)); out["ol_i"] = np.zeros((n, 9), np.int32)
            for i in range(n):
                e = self.envs[i]
                out["duck"][i] = e.duck_pos[:]
                for k in range(e.n_obst):
                    out["obst"][i, k] = e.obst[k][:]
                out["ol_f"][i] = [e.last_cx, e.last_cy, e.last_area, e.last_depth, e.frame_cx, e.frame_cy, e.frame_area,
                                  e.frame_depth, e.frame_dl, e.frame_dc, e.frame_dr, e.prev_est_dist]
                out["ol_i"][i] = [e.duck_phase, e.has_prev_dist, e.post_waypoints, e.cam_valid, e.frame_visible,
                                  e.seen_consecutive, e.lock_steps, e.steps_since_seen, e.n_obst]
        if self.cfg.task == 4:                           # duck-only task: history rows (newest first), deltas; slot 5 = filled
            out["vis_hist"] = np.zeros((n, MAX_HIST * 9 + 4))
            for i in range(n):
                e = self.envs[i]
                for h in range(MAX_HIST):
                    out["vis_hist"][i, h * 9:(h + 1) * 9] = e.vis_hist[h][:]
                out["vis_hist"][i, MAX_HIST * 9:] = e.vis_deltas[:]
                out["ol_i"][i, 5] = e.hist_filled
        return out

    def set_state(self, s: dict) -> None:
        """Overwrite dynamic state from arrays (any subset of get_state keys) and refresh the cached
        surface velocities the way Fixedwing.update_state would have left them."""
        for i in range(self.n):
            e = self.envs[i]
            if "pos" in s: e.pos[:] = [float(x) for x in s["pos"][i]]
            if "quat" in s: e.quat[:] = [float(x) for x in s["quat"][i]]
            if "vel" in s: e.vel[:] = [float(x) for x in s["vel"][i]]
            if "omega" in s: e.omega[:] = [float(x) for x in s["omega"][i]]
            if "act" in s:
                e.act[:] = [float(x) for x in s["act"][i][:5]]; e.throttle = float(s["act"][i][5])
            if "step_count" in s: e.step_count = int(s["step_count"][i])
            if "physics_steps" in s: e.physics_steps = int(s["physics_steps"][i])
            if "episode" in s: e.episode = int(s["episode"][i])
            if "new_dist" in s: e.new_dist = float(s["new_dist"][i])
            if "wind" in s:
                w = s["wind"][i]
                e.wind_base[:] = [float(x) for x in w[:3]]; e.gust_amp[:] = [float(x) for x in w[3:6]]
                e.gust_phase = float(w[6])
            if "targets" in s and self.cfg.task == 3:
                e.target_ref[:] = [float(x) for x in s["targets"][i][0]]
            elif "targets" in s:
                tidx = int(s["target_idx"][i]) if "target_idx" in s else e.target_idx
                T = self.num_targets
                e.target_idx = tidx; e.n_remaining = T - tidx
                for t in range(T - tidx):
                    e.targets[t][:] = [float(x) for x in s["targets"][i][tidx + t]]
            if "duck" in s: e.duck_pos[:] = [float(x) for x in s["duck"][i]]
            if "ol_f" in s:
                f = [float(x) for x in s["ol_f"][i]]
                (e.last_cx, e.last_cy, e.last_area, e.last_depth, e.frame_cx, e.frame_cy, e.frame_area, e.frame_depth,
                 e.frame_dl, e.frame_dc, e.frame_dr, e.prev_est_dist) = f
            if "ol_i" in s:
                v = [int(x) for x in s["ol_i"][i]]
                (e.duck_phase, e.has_prev_dist, e.post_waypoints, e.cam_valid, e.frame_visible, e.seen_consecutive,
                 e.lock_steps, e.steps_since_seen, e.n_obst) = v
                if self.cfg.task == 4:
                    e.hist_filled = v[5]
            if "vis_hist" in s and self.cfg.task == 4:
                vh = [float(x) for x in s["vis_hist"][i]]
                for h in range(MAX_HIST):
                    e.vis_hist[h][:] = vh[h * 9:(h + 1) * 9]
                e.vis_deltas[:] = vh[MAX_HIST * 9:MAX_HIST * 9 + 4]
            if "obst" in s:
                for k in range(MAX_OBST):
                    e.obst[k][:] = [float(x) for x in s["obst"][i][k]]
            self.L.fwo_refresh_surface_vel(C.byref(self.cfg), C.byref(e), e.physics_steps - 1)
