/*
 * fw_oracle.h -- fp64 CPU ORACLE for the fixed-wing env hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pyflyt_drone_b200/, include/)
 * may import, link or execute this.  Allowed callers: tests/, __graft_entry__.smoke(),
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference (WdBlink/pyflyt-drone) has no tests and no golden vectors,
 * and the arithmetic of this path lives in un-vendored, un-pinned third-party packages
 * (PyFlyt, pybullet/Bullet3) that are absent from /root/reference and not installable
 * offline.  This file restates (a) the in-tree env glue, citing /root/reference file:line,
 * and (b) the published upstream algorithms (Khan & Nahon 2015 lifting-surface model as
 * implemented by PyFlyt; Bullet btMultiBody semi-implicit Euler), marked [UP-RECALL].
 * It is pinned only by the closed-form KATs of SURVEY.md section 8(c).
 */
#ifndef FW_ORACLE_H
#define FW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FWO_NSURF 5        /* cmd order: left aileron, right aileron, h-tail, v-tail, main wing */
#define FWO_MAX_TARGETS 16
#define FWO_MAX_COL 16
#define FWO_MAX_OBST 32
#define FWO_MAX_OBS 72
#define FWO_MAX_HIST 4     /* duck-only task: frames of vision history in the observation */

/* flag bits returned by a step (same meaning as include/fwsim.h, restated independently) */
#define FWO_TERM      1
#define FWO_TRUNC     2
#define FWO_COLLISION 4
#define FWO_OOB       8
#define FWO_COMPLETE  16
#define FWO_STRIKE    32

/* task ids */
#define FWO_TASK_PHYSICS   0  /* dynamics + ground/dome termination only (BASELINE config 2) */
#define FWO_TASK_WAYPOINTS 1  /* PyFlyt/Fixedwing-Waypoints-v3 + FlattenWaypointEnv */
#define FWO_TASK_OBJLOCK   2  /* FixedwingWaypointObjLockEnv + FlattenWaypointEnv */
#define FWO_TASK_LOWLEVEL  3  /* FixedwingLowLevelEnv: mode -1 six-channel control, psi/h/V tracking (fixedwing_lowlevel_env.py) */
#define FWO_TASK_DUCK      4  /* FixedwingObjLockEnv + FlattenObjLockEnv: duck-only lock/strike, vision-history obs (fixedwing_objlock_env.py) */
#define FWO_MAX_ACT 6

typedef struct {
    /* ---- lifting surfaces: my_models/fixedwing/fixewing.yaml:8-71 ---- */
    double cl_alpha_2d[FWO_NSURF];
    double chord[FWO_NSURF];
    double span[FWO_NSURF];
    double flap_to_chord[FWO_NSURF];
    double eta[FWO_NSURF];
    double alpha0_base_deg[FWO_NSURF];
    double stall_p_base_deg[FWO_NSURF];
    double stall_n_base_deg[FWO_NSURF];
    double cd0[FWO_NSURF];
    double defl_limit_deg[FWO_NSURF];
    double surf_tau[FWO_NSURF];
    double lift_unit[FWO_NSURF][3];
    double fwd_unit[FWO_NSURF][3];
    double r_surf[FWO_NSURF][3];   /* link CoM relative to base-link CoM, body frame */
    /* ---- motor: fixewing.yaml:1-6 ---- */
    double total_thrust, thrust_coef, torque_coef, noise_ratio, motor_tau;
    double r_motor[3], thrust_unit[3];
    /* ---- composite rigid body expressed at the base-link CoM (point O) ---- */
    double mass;
    double com[3];          /* composite CoM relative to O, body frame */
    double inertia_o[9];    /* composite inertia about O, body frame, row-major */
    int32_t n_col;
    int32_t _pad0;
    double col_pts[FWO_MAX_COL][3];  /* collision probe points, body frame */
    double contact_margin;
    /* ---- simulator ---- */
    double dt;              /* 1/240 */
    double gravity;         /* 9.81 */
    double rho;             /* 1.225 */
    double max_coord_vel;   /* Bullet btMultiBody m_maxCoordinateVelocity = 100 */
    int32_t physics_per_control;   /* 2: control latches every 2nd physics step */
    int32_t substeps_per_inner;    /* 2: Aviary.step() = physics_hz / control_hz */
    int32_t inner_per_step;        /* 4: env_step_ratio = 120 / agent_hz */
    int32_t warmup_inner;          /* 10 Aviary.step() calls in end_reset */
    /* ---- [UP-RECALL] conventions that are not verifiable offline ---- */
    double ail_left_sign, ail_right_sign, pitch_sign, yaw_sign;
    int32_t freestream_3d;   /* 1: V = |v|_3 ; 0: V = hypot(lift_speed, fwd_speed) */
    int32_t cd90_degrees;    /* 1: Cd90 polynomial evaluated on the deflection in degrees */
    /* ---- env ---- */
    int32_t task;
    int32_t num_targets;
    double goal_reach;
    int32_t sparse_reward;
    int32_t angle_repr;      /* 0 euler, 1 quaternion */
    double dome;
    int32_t max_steps;
    int32_t context_len;
    double start_pos[3], start_vel[3];
    double spawn_size, min_height;
    int32_t early_return_on_crash;  /* fixedwing_waypoint_objlock_env.py:282-283 */
    int32_t complete_truncates;     /* upstream Waypoints env: truncation |= all_targets_reached */
    /* ---- wind: fixedwing_base_env.py:108-173, envs/utils.py:141-218 ---- */
    int32_t wind_mode;        /* 0 off, 1 constant, 2 gust_sine */
    int32_t wind_randomize;   /* randomize_on_reset */
    int32_t wind_rand_phase;  /* randomize_gust_phase */
    int32_t wind_start_substep; /* first state refresh that sees the wind (0 env hook, 20 wrapper) */
    double wind_base[3], wind_base_lo[3], wind_base_hi[3];
    double gust_amp[3], gust_amp_lo[3], gust_amp_hi[3];
    double gust_freq, gust_phase;
    /* ---- ObjLock task: fixedwing_waypoint_objlock_env.py:42-168 ---- */
    int32_t num_obstacles;
    int32_t cam_interval_substeps;   /* physics_control_ratio * capture_interval = 12 */
    double obst_radius, obst_h_lo, obst_h_hi, obst_safe, obst_scale, obst_max_pen;
    int32_t lock_hold_steps;
    int32_t switch_min_seen;
    double strike_dist, strike_reward, lock_step_reward, approach_scale, switch_min_area;
    double duck_radius;              /* analytic stand-in for the duck mesh extent */
    double cam_offset[3];            /* camera position offset, body frame (-3,0,1) */
    double cam_near, cam_far;
    int32_t cam_res;                 /* 128 */
    int32_t cam_mode;                /* 0: tracking chase camera, looks at the aircraft from cam_offset;
                                      * 1: fixed camera at cam_offset looking along the body x axis tilted by cam_tilt_deg
                                      *    (is_tracking_camera = False, fixedwing_objlock_env.py:184-231) */
    /* ---- duck-only task (FWO_TASK_DUCK): fixedwing_objlock_env.py:37-118 ---- */
    double cam_tilt_deg;             /* camera_angle_degrees; [UP-RECALL] rotation about body +y, positive = nose-down */
    double duck_dist_scale, lock_center_radius, centering_scale, visible_step_reward, area_reward_scale;
    double lock_lost_penalty, approach_clip;
    int32_t vision_hist_len;         /* duck_vision_history_len, 1..FWO_MAX_HIST */
    int32_t vision_use_deltas;       /* duck_vision_use_deltas */
    int32_t lock_decay_steps;        /* duck_lock_decay_steps (>= 1) */
    int32_t _pad1;
} fwo_config;

typedef struct {
    double pos[3];
    double quat[4];       /* x,y,z,w  body->world */
    double vel[3];        /* world */
    double omega[3];      /* world */
    double act[FWO_NSURF];
    double throttle;
    double surf_vel[FWO_NSURF][3];  /* cached local surface velocities (update_state) */
    double setpoint[FWO_MAX_ACT];   /* aviary setpoint (mode 0: 4 channels, thrust already remapped to [0,1]; mode -1: 6) */
    double cmd[6];
    double last_action[FWO_MAX_ACT];
    double target_ref[3];           /* low-level task: psi_ref, h_ref, V_ref */
    double targets[FWO_MAX_TARGETS][3];
    int32_t n_remaining;  /* len(waypoints.targets) */
    int32_t target_idx;   /* index of targets[0] in the original list */
    double old_dist, new_dist;
    int32_t step_count;
    int32_t physics_steps;
    int32_t termination, truncation;
    int32_t info_collision, info_oob, info_complete, info_strike;
    int32_t num_targets_reached;
    int32_t contact;       /* any(contact_array) for the current Aviary.step */
    uint32_t episode;
    uint32_t env_id;
    double reward;
    double wind_base[3], gust_amp[3], gust_phase;
    /* objlock */
    double duck_pos[3];
    double obst[FWO_MAX_OBST][3];   /* x, y, h */
    int32_t n_obst;
    int32_t duck_phase, seen_consecutive, lock_steps, has_prev_dist, post_waypoints;
    int32_t steps_since_seen;
    int32_t cam_valid;
    double prev_est_dist, last_cx, last_cy, last_area, last_depth;
    double vision[9];
    /* captured camera frame (latest) */
    int32_t frame_visible; int32_t _pad;
    double frame_cx, frame_cy, frame_area, frame_depth, frame_dl, frame_dc, frame_dr;
    double ep_return; int32_t ep_length; int32_t _pad2;
    /* duck-only task: _vision_history (row 0 = newest), the four delta features, _vision_history_filled */
    double vis_hist[FWO_MAX_HIST][9];
    double vis_deltas[4];
    double target_vec[3];            /* state["target_vector"]: duck position relative to the aircraft, body frame */
    int32_t hist_filled; int32_t _pad3;
} fwo_env;

/* size checks for the ctypes mirror */
int fwo_config_size(void);
int fwo_env_size(void);
int fwo_obs_dim(const fwo_config* c);

/* counter-based RNG shared by spec with the device path: Philox4x32-10 */
void fwo_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
double fwo_u01(uint32_t x);
void fwo_normals4(uint64_t seed, uint32_t env, uint32_t episode, uint32_t idx, double out[4]);
void fwo_random_action(uint64_t seed, uint32_t env, uint32_t episode, uint32_t step_count, double out[4]);
/* channels 4,5 of a six-channel action: a second Philox call at index step_count | 0x40000000 */
void fwo_random_action6(uint64_t seed, uint32_t env, uint32_t episode, uint32_t step_count, double out[6]);

/* lifting-surface model ([UP-RECALL] PyFlyt LiftingSurface): returns Cl, Cd, CM */
void fwo_aero_coeffs(const fwo_config* c, int s, double alpha, double actuation, double out[3]);
void fwo_surface_force(const fwo_config* c, int s, double actuation, const double vel[3],
                       double force[3], double torque[3]);

/* one 240 Hz physics substep (update_control, update_physics, stepSimulation, update_state) */
void fwo_substep(const fwo_config* c, fwo_env* e, uint64_t seed);

/* gymnasium-level API */
void fwo_reset(const fwo_config* c, fwo_env* e, uint64_t seed, uint32_t env_id, uint32_t episode, double* obs);
int fwo_act_dim(const fwo_config* c);   /* 4, or 6 for the low-level task */
void fwo_step(const fwo_config* c, fwo_env* e, uint64_t seed, const double* action,
              double* obs, double* reward, int32_t* flags);
/* debug / evaluation frame of one env (FixedwingBaseEnv.render): rgba [H,W,4], seg [H,W], depth-buffer values [H,W]; any
 * output may be NULL.  See fw_oracle.c for the scene and the class ids. */
void fwo_render(const fwo_config* c, const fwo_env* e, int W, int H, uint8_t* rgba, int32_t* seg, double* depth);
/* SubprocVecEnv worker semantics: step, and on done stash terminal obs then reset */
void fwo_vec_step(const fwo_config* c, fwo_env* envs, int n, uint64_t seed, const double* actions,
                  double* obs, double* rewards, int32_t* flags, double* term_obs, int nthreads);
/* fwo_vec_step plus info["num_targets_reached"] of every env, taken before the reset-on-done of a finished episode */
void fwo_vec_step_info(const fwo_config* c, fwo_env* envs, int n, uint64_t seed, const double* actions,
                       double* obs, double* rewards, int32_t* flags, double* term_obs, int32_t* targets_reached,
                       int nthreads);
void fwo_vec_reset(const fwo_config* c, fwo_env* envs, int n, uint64_t seed, uint32_t env_id0,
                   double* obs, int nthreads);
/* random-action rollout used as the CPU baseline: returns env-steps executed */
long fwo_rollout_random(const fwo_config* c, fwo_env* envs, int n, uint64_t seed, int steps, int nthreads);

void fwo_compute_obs(const fwo_config* c, fwo_env* e, double* obs, int update_dist);
void fwo_refresh_surface_vel(const fwo_config* c, fwo_env* e, int stamp);
void fwo_quat_to_euler(const double q[4], double rpy[3]);
void fwo_euler_to_quat(const double rpy[3], double q[4]);
void fwo_quat_to_mat(const double q[4], double R[9]);

#ifdef __cplusplus
}
#endif
#endif
