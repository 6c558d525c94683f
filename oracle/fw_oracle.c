/*
 * fw_oracle.c -- fp64 CPU ORACLE for the fixed-wing env hot path (see fw_oracle.h header note).
 *
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no golden vectors exist upstream).
 *
 * What each function restates:
 *   fwo_step / compute_base_term_trunc   /root/reference/envs/fixedwing_envs/fixedwing_base_env.py:296-348
 *   fwo_reset                            .../fixedwing_base_env.py:193-257
 *   fwo_compute_obs                      .../fixedwing_base_env.py:263-290,
 *                                        /root/reference/envs/fixedwing_waypoint_objlock_env.py:197-276,
 *                                        /root/reference/envs/flatten_waypoint_env.py:52-72
 *   waypoint reward                      /root/reference/envs/fixedwing_waypoint_objlock_env.py:278-343 (in-tree)
 *                                        and upstream FixedwingWaypointsEnv [UP-RECALL], call site
 *                                        /root/reference/train/train_Fixedwing_Waypoints_v3.py:100-110
 *   wind                                 .../fixedwing_base_env.py:108-173, /root/reference/envs/utils.py:141-218
 *   aircraft constants                   /root/reference/my_models/fixedwing/fixewing.yaml:1-71 (passed in via fwo_config)
 *   fwo_aero_coeffs / fwo_surface_force  [UP-RECALL] PyFlyt core/abstractions/lifting_surfaces.py (Khan & Nahon 2015)
 *   motor                                [UP-RECALL] PyFlyt core/abstractions/motors.py
 *   fwo_substep sequencing               [UP-RECALL] PyFlyt core/aviary.py Aviary.step, core/drones/fixedwing.py
 *   rigid body                           [UP-RECALL] Bullet3 btMultiBody.cpp (ABA for a base + fixed links ==
 *                                        one rigid body expressed at the base-link CoM; semi-implicit Euler;
 *                                        exponential-map quaternion update; +-100 coordinate-velocity clamp)
 *   euler/quaternion                     [UP-RECALL] pybullet getEulerFromQuaternion / getQuaternionFromEuler
 *   waypoint sampling                    [UP-RECALL] PyFlyt gym_envs/utils/waypoint_handler.py
 */
#include "fw_oracle.h"

#include <math.h>
#include <string.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define STREAM_TARGETS 1u
#define STREAM_NOISE   2u
#define STREAM_ACTION  3u
#define STREAM_WIND    4u
#define STREAM_DUCK    5u
#define STREAM_OBST    6u

int fwo_config_size(void) { return (int)sizeof(fwo_config); }
int fwo_env_size(void) { return (int)sizeof(fwo_env); }

int fwo_act_dim(const fwo_config* c) { return c->task == FWO_TASK_LOWLEVEL ? 6 : 4; }

int fwo_obs_dim(const fwo_config* c) {
    if (c->task == FWO_TASK_PHYSICS) return 0;
    if (c->task == FWO_TASK_LOWLEVEL) return 21;      /* fixedwing_lowlevel_env.py:64-68 */
    int att = (c->angle_repr == 0 ? 12 : 13) + 4 + 6;
    if (c->task == FWO_TASK_DUCK)                     /* flatten_objlock_env.py:20-31: attitude + target_vector + duck_vision */
        return att + 3 + 9 * c->vision_hist_len + (c->vision_use_deltas ? 4 : 0);
    return att + 3 * c->context_len;
}

/* ------------------------------------------------------------------ RNG */

void fwo_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 23-bit uniform in (0,1): (k + 0.5) / 2^23 needs 24 significant bits, exact in fp32 and fp64 */
double fwo_u01(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }

void fwo_normals4(uint64_t seed, uint32_t env, uint32_t episode, uint32_t idx, double out[4]) {
    uint32_t r[4];
    fwo_philox(seed, env, episode, idx, STREAM_NOISE, r);
    double u0 = fwo_u01(r[0]), u1 = fwo_u01(r[1]), u2 = fwo_u01(r[2]), u3 = fwo_u01(r[3]);
    double ra = sqrt(-2.0 * log(u0)), rb = sqrt(-2.0 * log(u2));
    out[0] = ra * cos(2.0 * M_PI * u1);
    out[1] = ra * sin(2.0 * M_PI * u1);
    out[2] = rb * cos(2.0 * M_PI * u3);
    out[3] = rb * sin(2.0 * M_PI * u3);
}

void fwo_random_action(uint64_t seed, uint32_t env, uint32_t episode, uint32_t step_count, double out[4]) {
    uint32_t r[4];
    fwo_philox(seed, env, episode, step_count, STREAM_ACTION, r);
    for (int i = 0; i < 4; ++i) out[i] = 2.0 * fwo_u01(r[i]) - 1.0;
}

void fwo_random_action6(uint64_t seed, uint32_t env, uint32_t episode, uint32_t step_count, double out[6]) {
    uint32_t r[4];
    fwo_random_action(seed, env, episode, step_count, out);
    fwo_philox(seed, env, episode, step_count | 0x40000000u, STREAM_ACTION, r);
    out[4] = 2.0 * fwo_u01(r[0]) - 1.0;
    out[5] = 2.0 * fwo_u01(r[1]) - 1.0;
}

/* ------------------------------------------------------------------ small vector helpers */

static inline void cross3(const double a[3], const double b[3], double o[3]) {
    double x = a[1] * b[2] - a[2] * b[1];
    double y = a[2] * b[0] - a[0] * b[2];
    double z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void mat_vec(const double R[9], const double v[3], double o[3]) {
    double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void matT_vec(const double R[9], const double v[3], double o[3]) {
    double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}

/* pybullet getMatrixFromQuaternion: row-major body->world */
void fwo_quat_to_mat(const double q[4], double R[9]) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double d = x * x + y * y + z * z + w * w;
    double s = 2.0 / d;
    double xs = x * s, ys = y * s, zs = z * s;
    double wx = w * xs, wy = w * ys, wz = w * zs;
    double xx = x * xs, xy = x * ys, xz = x * zs;
    double yy = y * ys, yz = y * zs, zz = z * zs;
    R[0] = 1.0 - (yy + zz); R[1] = xy - wz;          R[2] = xz + wy;
    R[3] = xy + wz;          R[4] = 1.0 - (xx + zz); R[5] = yz - wx;
    R[6] = xz - wy;          R[7] = yz + wx;          R[8] = 1.0 - (xx + yy);
}

/* pybullet getEulerFromQuaternion (btQuaternion::getEulerZYX with its +-0.99999 guard) */
void fwo_quat_to_euler(const double q[4], double rpy[3]) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double sqx = x * x, sqy = y * y, sqz = z * z, squ = w * w;
    double sarg = -2.0 * (x * z - w * y);
    if (sarg <= -0.99999) {
        rpy[0] = 0.0; rpy[1] = -0.5 * M_PI; rpy[2] = 2.0 * atan2(x, -y);
    } else if (sarg >= 0.99999) {
        rpy[0] = 0.0; rpy[1] = 0.5 * M_PI; rpy[2] = 2.0 * atan2(-x, y);
    } else {
        rpy[0] = atan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz);
        rpy[1] = asin(sarg);
        rpy[2] = atan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz);
    }
}

/* pybullet getQuaternionFromEuler (btQuaternion::setEulerZYX) */
void fwo_euler_to_quat(const double rpy[3], double q[4]) {
    double hr = 0.5 * rpy[0], hp = 0.5 * rpy[1], hy = 0.5 * rpy[2];
    double cr = cos(hr), sr = sin(hr), cp = cos(hp), sp = sin(hp), cy = cos(hy), sy = sin(hy);
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
    q[3] = cr * cp * cy + sr * sp * sy;
}

/* ------------------------------------------------------------------ lifting surfaces */

static double interp2(double x, double x0, double x1, double f0, double f1) {
    /* numpy.interp on a two-point table: clamps outside [x0, x1] */
    if (x <= x0) return f0;
    if (x >= x1) return f1;
    return (f1 - f0) / (x1 - x0) * (x - x0) + f0;
}

void fwo_aero_coeffs(const fwo_config* c, int s, double alpha, double actuation, double out[3]) {
    const double deg = M_PI / 180.0;
    double aspect = c->span[s] / c->chord[s];
    double cla = c->cl_alpha_2d[s] * (aspect / (aspect + ((2.0 * (aspect + 4.0)) / (aspect + 2.0))));
    double theta_f = acos(2.0 * c->flap_to_chord[s] - 1.0);
    double aero_tau = 1.0 - ((theta_f - sin(theta_f)) / M_PI);
    double alpha0_base = c->alpha0_base_deg[s] * deg;
    double stall_p_base = c->stall_p_base_deg[s] * deg;
    double stall_n_base = c->stall_n_base_deg[s] * deg;
    double cd0 = c->cd0[s];

    double defl_deg = actuation * c->defl_limit_deg[s];
    double defl_rad = defl_deg * deg;

    double delta_cl = cla * aero_tau * c->eta[s] * defl_rad;
    double delta_cl_max = c->flap_to_chord[s] * delta_cl;
    double cl_max_p = cla * (stall_p_base - alpha0_base) + delta_cl_max;
    double cl_max_n = cla * (stall_n_base - alpha0_base) + delta_cl_max;
    double alpha0 = alpha0_base - (delta_cl / cla);
    double stall_p = alpha0 + (cl_max_p / cla);
    double stall_n = alpha0 + (cl_max_n / cla);

    double Cl, Cd, CM;
    if (stall_n < alpha && alpha < stall_p) {
        Cl = cla * (alpha - alpha0);
        double alpha_i = Cl / (M_PI * aspect);
        double ae = alpha - alpha0 - alpha_i;
        double CT = cd0 * cos(ae);
        double CN = (Cl + (CT * sin(ae))) / cos(ae);
        Cd = (CN * sin(ae)) + (CT * cos(ae));
        CM = -CN * (0.25 - (0.175 * (1.0 - ((2.0 * ae) / M_PI))));
    } else {
        double alpha_i;
        if (alpha > 0.0) {
            double cl_stall = cla * (stall_p - alpha0);
            double ai_stall = cl_stall / (M_PI * aspect);
            alpha_i = interp2(alpha, stall_p, M_PI / 2.0, ai_stall, 0.0);
        } else {
            double cl_stall = cla * (stall_n - alpha0);
            double ai_stall = cl_stall / (M_PI * aspect);
            alpha_i = interp2(alpha, -M_PI / 2.0, stall_n, 0.0, ai_stall);
        }
        double ae = alpha - alpha0 - alpha_i;
        double d = c->cd90_degrees ? defl_deg : defl_rad;
        double cd90 = ((-4.26e-2) * (d * d)) + ((2.1e-1) * d) + 1.98;
        double sa = sin(ae), ca = cos(ae);
        double CN = cd90 * sa * (1.0 / (0.56 + 0.44 * fabs(sa)) - 0.41 * (1.0 - exp(-17.0 / aspect)));
        double CT = 0.5 * cd0 * ca;
        Cl = (CN * ca) - (CT * sa);
        Cd = (CN * sa) + (CT * ca);
        CM = -CN * (0.25 - (0.175 * (1.0 - ((2.0 * fabs(ae)) / M_PI))));
    }
    out[0] = Cl; out[1] = Cd; out[2] = CM;
}

void fwo_surface_force(const fwo_config* c, int s, double actuation, const double vel[3],
                       double force[3], double torque[3]) {
    const double* lu = c->lift_unit[s];
    const double* fu = c->fwd_unit[s];
    double lift_speed = dot3(vel, lu);
    double fwd_speed = dot3(vel, fu);
    double V2 = c->freestream_3d ? dot3(vel, vel) : (lift_speed * lift_speed + fwd_speed * fwd_speed);
    double alpha = atan2(-lift_speed, fwd_speed);
    double coef[3];
    fwo_aero_coeffs(c, s, alpha, actuation, coef);
    double area = c->chord[s] * c->span[s];
    double Q = 0.5 * c->rho * V2 * area;
    double lift = coef[0] * Q, drag = coef[1] * Q;
    double fn = lift * cos(alpha) + drag * sin(alpha);
    double fp = lift * sin(alpha) - drag * cos(alpha);
    double tu[3];
    cross3(lu, fu, tu);
    for (int k = 0; k < 3; ++k) {
        force[k] = lu[k] * fn + fu[k] * fp;
        torque[k] = Q * coef[2] * c->chord[s] * tu[k];
    }
}

/* ------------------------------------------------------------------ wind */

static void wind_at(const fwo_config* c, const fwo_env* e, double t, double w[3]) {
    if (c->wind_mode == 1) {
        w[0] = e->wind_base[0]; w[1] = e->wind_base[1]; w[2] = e->wind_base[2];
    } else if (c->wind_mode == 2) {
        double s = sin(2.0 * M_PI * c->gust_freq * t + e->gust_phase);
        for (int k = 0; k < 3; ++k) w[k] = e->wind_base[k] + e->gust_amp[k] * s;
    } else {
        w[0] = w[1] = w[2] = 0.0;
    }
}

/* Fixedwing.update_state -> LiftingSurfaces.state_update: local velocity of every surface link CoM,
 * wind subtracted in the world frame.  `stamp` = physics_steps at the time of the refresh, -1 = no wind. */
void fwo_refresh_surface_vel(const fwo_config* c, fwo_env* e, int stamp) {
    double R[9];
    fwo_quat_to_mat(e->quat, R);
    double w[3] = {0, 0, 0};
    if (c->wind_mode != 0 && stamp >= c->wind_start_substep) wind_at(c, e, (double)stamp * c->dt, w);
    for (int s = 0; s < FWO_NSURF; ++s) {
        double rw[3], wxr[3], vw[3];
        mat_vec(R, c->r_surf[s], rw);
        cross3(e->omega, rw, wxr);
        for (int k = 0; k < 3; ++k) vw[k] = e->vel[k] + wxr[k] - w[k];
        matT_vec(R, vw, e->surf_vel[s]);
    }
}

/* ------------------------------------------------------------------ rigid body */

static void solve6(double A[6][6], double b[6]) {
    /* Gaussian elimination with partial pivoting, in place */
    for (int i = 0; i < 6; ++i) {
        int p = i;
        for (int r = i + 1; r < 6; ++r) if (fabs(A[r][i]) > fabs(A[p][i])) p = r;
        if (p != i) {
            for (int k = 0; k < 6; ++k) { double t = A[i][k]; A[i][k] = A[p][k]; A[p][k] = t; }
            double t = b[i]; b[i] = b[p]; b[p] = t;
        }
        double inv = 1.0 / A[i][i];
        for (int r = i + 1; r < 6; ++r) {
            double f = A[r][i] * inv;
            if (f == 0.0) continue;
            for (int k = i; k < 6; ++k) A[r][k] -= f * A[i][k];
            b[r] -= f * b[i];
        }
    }
    for (int i = 5; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < 6; ++k) s -= A[i][k] * b[k];
        b[i] = s / A[i][i];
    }
}

static void map_setpoint(const fwo_config* c, fwo_env* e) {
    if (c->task == FWO_TASK_LOWLEVEL) {
        /* [UP-RECALL] Fixedwing.update_control mode -1: cmd = setpoint (6 channels, surface order of the YAML +
         * motor); docstring of fixedwing_lowlevel_env.py:13-14 agrees on the order */
        for (int k = 0; k < 6; ++k) e->cmd[k] = e->setpoint[k];
        return;
    }
    /* [UP-RECALL] Fixedwing.update_control mode 0: [roll,pitch,yaw,thrust] ->
     * [left ail, right ail, h-tail, v-tail, main wing, motor] */
    e->cmd[0] = c->ail_left_sign * e->setpoint[0];
    e->cmd[1] = c->ail_right_sign * e->setpoint[0];
    e->cmd[2] = c->pitch_sign * e->setpoint[1];
    e->cmd[3] = c->yaw_sign * e->setpoint[2];
    e->cmd[4] = 0.0;
    e->cmd[5] = e->setpoint[3];
}

static int ground_contact(const fwo_config* c, const fwo_env* e, const double R[9]) {
    for (int i = 0; i < c->n_col; ++i) {
        double z = e->pos[2] + R[6] * c->col_pts[i][0] + R[7] * c->col_pts[i][1] + R[8] * c->col_pts[i][2];
        if (z <= c->contact_margin) return 1;
    }
    return 0;
}

static int obstacle_contact(const fwo_config* c, const fwo_env* e, const double R[9]) {
    if (c->task != FWO_TASK_OBJLOCK && c->task != FWO_TASK_DUCK) return 0;
    for (int i = 0; i < c->n_col; ++i) {
        double p[3], pw[3];
        p[0] = c->col_pts[i][0]; p[1] = c->col_pts[i][1]; p[2] = c->col_pts[i][2];
        mat_vec(R, p, pw);
        for (int k = 0; k < 3; ++k) pw[k] += e->pos[k];
        for (int o = 0; o < e->n_obst; ++o) {
            double dx = pw[0] - e->obst[o][0], dy = pw[1] - e->obst[o][1];
            double rr = c->obst_radius + c->contact_margin;
            if (dx * dx + dy * dy <= rr * rr && pw[2] <= e->obst[o][2] + c->contact_margin) return 1;
        }
        double dd[3] = {pw[0] - e->duck_pos[0], pw[1] - e->duck_pos[1], pw[2] - (e->duck_pos[2] + c->duck_radius)};
        double rd = c->duck_radius + c->contact_margin;
        if (dot3(dd, dd) <= rd * rd) return 1;
    }
    return 0;
}

void fwo_substep(const fwo_config* c, fwo_env* e, uint64_t seed) {
    const double dt = c->dt;
    /* --- drone.update_control(physics_steps) --- */
    if (e->physics_steps % c->physics_per_control == 0) map_setpoint(c, e);

    /* --- drone.update_physics(): surfaces then motor; forces in LINK_FRAME at the link CoM --- */
    double F[3] = {0, 0, 0}, T[3] = {0, 0, 0};
    for (int s = 0; s < FWO_NSURF; ++s) {
        e->act[s] += (dt / c->surf_tau[s]) * (e->cmd[s] - e->act[s]);
        double f[3], tq[3], rxf[3];
        fwo_surface_force(c, s, e->act[s], e->surf_vel[s], f, tq);
        cross3(c->r_surf[s], f, rxf);
        for (int k = 0; k < 3; ++k) { F[k] += f[k]; T[k] += rxf[k] + tq[k]; }
    }
    {
        e->throttle += (dt / c->motor_tau) * (e->cmd[5] - e->throttle);
        if (c->noise_ratio > 0.0) {
            double n[4];
            fwo_normals4(seed, e->env_id, e->episode, (uint32_t)e->physics_steps >> 2, n);
            e->throttle += n[e->physics_steps & 3] * e->throttle * c->noise_ratio;
        }
        double max_rpm = sqrt(c->total_thrust / c->thrust_coef);
        double rpm = e->throttle * max_rpm;
        double thrust = rpm * rpm * c->thrust_coef;
        double torque = rpm * rpm * c->torque_coef;
        double f[3], rxf[3];
        for (int k = 0; k < 3; ++k) f[k] = thrust * c->thrust_unit[k];
        cross3(c->r_motor, f, rxf);
        for (int k = 0; k < 3; ++k) { F[k] += f[k]; T[k] += rxf[k] + torque * c->thrust_unit[k]; }
    }

    /* --- stepSimulation(): collision detection happens on the pose entering the step --- */
    double R[9];
    fwo_quat_to_mat(e->quat, R);
    if (ground_contact(c, e, R) || obstacle_contact(c, e, R)) e->contact = 1;

    /* gravity acts on every link: M g at the composite CoM */
    double gw[3] = {0.0, 0.0, -c->gravity}, gb[3], cxg[3];
    matT_vec(R, gw, gb);
    cross3(c->com, gb, cxg);
    for (int k = 0; k < 3; ++k) { F[k] += c->mass * gb[k]; T[k] += c->mass * cxg[k]; }

    double wb[3], vb[3];
    matT_vec(R, e->omega, wb);
    matT_vec(R, e->vel, vb);
    (void)vb;

    /* Newton-Euler about the body-fixed point O (base-link CoM):
     *   I_O a + M c x A            = T - w x (I_O w)
     *   -M c x a + M A             = F - M w x (w x c)
     * a = angular acceleration, A = classical linear acceleration of O, all in the body frame */
    double Iw[3], wxIw[3], wxc[3], wxwxc[3];
    mat_vec(c->inertia_o, wb, Iw);
    cross3(wb, Iw, wxIw);
    cross3(wb, c->com, wxc);
    cross3(wb, wxc, wxwxc);
    double A[6][6];
    double b[6];
    memset(A, 0, sizeof(A));
    const double M = c->mass, cx = c->com[0], cy = c->com[1], cz = c->com[2];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = c->inertia_o[3 * i + j];
    /* M [c]x */
    double Cx[3][3] = {{0, -cz, cy}, {cz, 0, -cx}, {-cy, cx, 0}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][3 + j] = M * Cx[i][j]; A[3 + i][j] = -M * Cx[i][j]; }
    for (int i = 0; i < 3; ++i) A[3 + i][3 + i] = M;
    for (int k = 0; k < 3; ++k) { b[k] = T[k] - wxIw[k]; b[3 + k] = F[k] - M * wxwxc[k]; }
    solve6(A, b);
    double aw[3], Aw[3];
    mat_vec(R, &b[0], aw);
    mat_vec(R, &b[3], Aw);

    /* semi-implicit Euler; btMultiBody::applyDeltaVeeMultiDof clamps every coordinate velocity */
    for (int k = 0; k < 3; ++k) {
        e->omega[k] += aw[k] * dt;
        e->vel[k] += Aw[k] * dt;
        if (e->omega[k] > c->max_coord_vel) e->omega[k] = c->max_coord_vel;
        if (e->omega[k] < -c->max_coord_vel) e->omega[k] = -c->max_coord_vel;
        if (e->vel[k] > c->max_coord_vel) e->vel[k] = c->max_coord_vel;
        if (e->vel[k] < -c->max_coord_vel) e->vel[k] = -c->max_coord_vel;
    }
    for (int k = 0; k < 3; ++k) e->pos[k] += e->vel[k] * dt;

    /* btMultiBody::stepPositionsMultiDof: exponential map with the pi/4 per-step limiter */
    {
        double wv[3] = {e->omega[0], e->omega[1], e->omega[2]};
        double ang = sqrt(dot3(wv, wv));
        if (ang * dt > 0.25 * M_PI) ang = 0.5 * (0.5 * M_PI) / dt;
        double k;
        if (ang < 0.001) k = 0.5 * dt - (dt * dt * dt) * 0.020833333333 * ang * ang;
        else k = sin(0.5 * ang * dt) / ang;
        double ax = wv[0] * k, ay = wv[1] * k, az = wv[2] * k, aw_ = cos(0.5 * ang * dt);
        double qx = e->quat[0], qy = e->quat[1], qz = e->quat[2], qw = e->quat[3];
        /* dq (x) q */
        double nx = aw_ * qx + qw * ax + (ay * qz - az * qy);
        double ny = aw_ * qy + qw * ay + (az * qx - ax * qz);
        double nz = aw_ * qz + qw * az + (ax * qy - ay * qx);
        double nw = aw_ * qw - (ax * qx + ay * qy + az * qz);
        double n = sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
        e->quat[0] = nx / n; e->quat[1] = ny / n; e->quat[2] = nz / n; e->quat[3] = nw / n;
    }

    /* --- drone.update_state(): cached surface velocities for the NEXT substep --- */
    fwo_refresh_surface_vel(c, e, e->physics_steps);
    e->physics_steps += 1;
}

/* ------------------------------------------------------------------ ObjLock analytic camera (stand-in) */

static double depth_buf(const fwo_config* c, double z) {
    if (z < c->cam_near) z = c->cam_near;
    if (z > c->cam_far) return 1.0;
    return c->cam_far * (z - c->cam_near) / ((c->cam_far - c->cam_near) * z);
}
static double buf_to_m(const fwo_config* c, double d) {
    double denom = c->cam_far - (c->cam_far - c->cam_near) * d;
    if (fabs(denom) < 1e-9) return c->cam_far;
    return c->cam_far * c->cam_near / denom;
}

/* nearest hit of ray o + t*d (t>0) against the finite vertical cylinder (x,y,h), radius r */
static double ray_cylinder(const double o[3], const double d[3], const double cyl[3], double r) {
    double ox = o[0] - cyl[0], oy = o[1] - cyl[1];
    double a = d[0] * d[0] + d[1] * d[1];
    double best = INFINITY;
    if (a > 1e-12) {
        double b = ox * d[0] + oy * d[1];
        double cc = ox * ox + oy * oy - r * r;
        double disc = b * b - a * cc;
        if (disc >= 0.0) {
            double sq = sqrt(disc);
            double t0 = (-b - sq) / a, t1 = (-b + sq) / a;
            double ts[2] = {t0, t1};
            for (int i = 0; i < 2; ++i) {
                double t = ts[i];
                if (t > 0.0) {
                    double z = o[2] + t * d[2];
                    if (z >= 0.0 && z <= cyl[2] && t < best) best = t;
                }
            }
        }
    }
    if (fabs(d[2]) > 1e-12) { /* top cap */
        double t = (cyl[2] - o[2]) / d[2];
        if (t > 0.0) {
            double x = ox + t * d[0], y = oy + t * d[1];
            if (x * x + y * y <= r * r && t < best) best = t;
        }
    }
    return best;
}

static double ray_sphere(const double o[3], const double d[3], const double ctr[3], double r) {
    double oc[3] = {o[0] - ctr[0], o[1] - ctr[1], o[2] - ctr[2]};
    double a = dot3(d, d), b = dot3(oc, d), cc = dot3(oc, oc) - r * r;
    double disc = b * b - a * cc;
    if (disc < 0.0) return INFINITY;
    double t = (-b - sqrt(disc)) / a;
    return t > 0.0 ? t : INFINITY;
}

/* Camera.capture_image stand-in: pin-hole chase camera (fov 90, square), duck = sphere, obstacles =
 * cylinders, ground plane z=0; only what _compute_vision_features consumes is produced
 * (/root/reference/envs/fixedwing_waypoint_objlock_env.py:575-693). */
static void capture_frame(const fwo_config* c, fwo_env* e) {
    double R[9];
    fwo_quat_to_mat(e->quat, R);
    double off_w[3], cam[3], f[3], up[3] = {R[2], R[5], R[8]}, r[3], u[3];
    mat_vec(R, c->cam_offset, off_w);
    for (int k = 0; k < 3; ++k) cam[k] = e->pos[k] + off_w[k];
    if (c->cam_mode == 0) {                       /* tracking camera: target = the aircraft, up = body z */
        double ol = sqrt(dot3(off_w, off_w));
        for (int k = 0; k < 3; ++k) f[k] = -off_w[k] / ol;
    } else {
        /* [UP-RECALL] PyFlyt Camera with is_tracking_camera = False (fixedwing_objlock_env.py:226-227): the camera sits at
         * the offset and looks along the body x axis rotated about body +y by camera_angle_degrees */
        double t = c->cam_tilt_deg * M_PI / 180.0;
        double fb[3] = {cos(t), 0.0, -sin(t)}, ub[3] = {sin(t), 0.0, cos(t)};
        mat_vec(R, fb, f);
        mat_vec(R, ub, up);
    }
    cross3(f, up, r);
    double rl = sqrt(dot3(r, r));
    for (int k = 0; k < 3; ++k) r[k] /= rl;
    cross3(r, f, u);

    double ctr[3] = {e->duck_pos[0], e->duck_pos[1], e->duck_pos[2] + c->duck_radius};
    double dv[3] = {ctr[0] - cam[0], ctr[1] - cam[1], ctr[2] - cam[2]};
    double zc = dot3(dv, f), xc = dot3(dv, r), yc = dot3(dv, u);
    double Rd = c->duck_radius;
    int vis = 0;
    if (zc - Rd > c->cam_near && fabs(xc) <= zc + Rd && fabs(yc) <= zc + Rd && zc - Rd < c->cam_far) {
        vis = 1;
        double dist = sqrt(dot3(dv, dv));
        double dir[3] = {dv[0] / dist, dv[1] / dist, dv[2] / dist};
        for (int o = 0; o < e->n_obst && vis; ++o) {
            double t = ray_cylinder(cam, dir, e->obst[o], c->obst_radius);
            if (t < dist - Rd) vis = 0;
        }
    }
    e->frame_visible = vis;
    if (vis) {
        double cx = 0.5 + 0.5 * xc / zc, cy = 0.5 - 0.5 * yc / zc;
        e->frame_cx = cx < 0 ? 0 : (cx > 1 ? 1 : cx);
        e->frame_cy = cy < 0 ? 0 : (cy > 1 ? 1 : cy);
        double a = M_PI * (Rd / zc) * (Rd / zc) * 0.25;
        e->frame_area = a > 1.0 ? 1.0 : a;
        e->frame_depth = zc - Rd;
    }
    /* mid-row obstacle bands: row h//2, columns [0,w/3) [w/3,2w/3) [2w/3,w); mean depth-BUFFER value
     * of non-duck pixels, then converted to metres (0 when the band has no non-duck pixel) */
    int w = c->cam_res, x1 = w / 3, x2 = 2 * w / 3, ymid = w / 2;
    double vrow = (2.0 * (ymid + 0.5) / w - 1.0);
    double sum[3] = {0, 0, 0};
    int cnt[3] = {0, 0, 0};
    for (int i = 0; i < w; ++i) {
        double xn = 2.0 * (i + 0.5) / w - 1.0;
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = f[k] + xn * r[k] - vrow * u[k];
        double best = INFINITY;
        int is_duck = 0;
        if (d[2] < -1e-12) { double t = -cam[2] / d[2]; if (t > 0 && t < best) best = t; }
        for (int o = 0; o < e->n_obst; ++o) {
            double t = ray_cylinder(cam, d, e->obst[o], c->obst_radius);
            if (t < best) best = t;
        }
        double td = ray_sphere(cam, d, ctr, Rd);
        if (td < best) { best = td; is_duck = 1; }
        if (is_duck) continue;
        int band = i < x1 ? 0 : (i < x2 ? 1 : 2);
        /* d has unit component along f, so t is the depth along the optical axis */
        sum[band] += isfinite(best) ? depth_buf(c, best) : 1.0;
        cnt[band] += 1;
    }
    double out[3];
    for (int b = 0; b < 3; ++b) {
        double m = cnt[b] ? sum[b] / cnt[b] : 0.0;
        out[b] = m > 0.0 ? buf_to_m(c, m) : 0.0;
    }
    e->frame_dl = out[0]; e->frame_dc = out[1]; e->frame_dr = out[2];
    e->cam_valid = 1;
}

/* Debug / evaluation frame of ONE env (FixedwingBaseEnv.render -> getCameraImage,
 * /root/reference/envs/fixedwing_envs/fixedwing_base_env.py:350-369; eval/eval_objlock.py:120-162 also keeps the
 * segmentation mask and the depth buffer): the scene of capture_frame() through the same pin-hole camera (vertical field of
 * view 90 degrees, horizontal scaled by W / H), every pixel ray-cast against the ground plane, the obstacle cylinders, the
 * duck sphere (camera tasks) and one sphere of radius goal_reach per waypoint not reached yet.
 *   seg:   -1 sky, 0 ground, 1 duck, 2 + k obstacle k, 64 + t waypoint t (index in the episode's original list)
 *   depth: OpenGL depth-buffer value of the hit (1.0 = far plane / sky), what pybullet's getCameraImage returns
 *   rgba:  flat colour per class, Lambert-shaded with a fixed light; alpha 255
 * PyBullet's rasteriser and meshes are not restated (nothing of the hot path reads these pixels); the frame is consistent
 * with the features the policy sees: its middle row at W = H = cam_res is the row capture_frame() integrates. */
void fwo_render(const fwo_config* c, const fwo_env* e, int W, int H, uint8_t* rgba, int32_t* seg, double* depth) {
    double R[9];
    fwo_quat_to_mat(e->quat, R);
    double off_w[3], cam[3], f[3], up[3] = {R[2], R[5], R[8]}, r[3], u[3];
    mat_vec(R, c->cam_offset, off_w);
    for (int k = 0; k < 3; ++k) cam[k] = e->pos[k] + off_w[k];
    if (c->cam_mode == 0) {
        double ol = sqrt(dot3(off_w, off_w));
        for (int k = 0; k < 3; ++k) f[k] = -off_w[k] / ol;
    } else {
        double t = c->cam_tilt_deg * M_PI / 180.0;
        double fb[3] = {cos(t), 0.0, -sin(t)}, ub[3] = {sin(t), 0.0, cos(t)};
        mat_vec(R, fb, f);
        mat_vec(R, ub, up);
    }
    cross3(f, up, r);
    double rl = sqrt(dot3(r, r));
    for (int k = 0; k < 3; ++k) r[k] /= rl;
    cross3(r, f, u);
    const int cam_task = c->task == FWO_TASK_OBJLOCK || c->task == FWO_TASK_DUCK;
    const int wp_task = c->task == FWO_TASK_WAYPOINTS || c->task == FWO_TASK_OBJLOCK;
    const double Rd = c->duck_radius;
    const double ctr[3] = {e->duck_pos[0], e->duck_pos[1], e->duck_pos[2] + Rd};
    const double L[3] = {0.30151134457776363, 0.20100756305184242, 0.9320390859672263};   /* (0.3, 0.2, 0.9273..) normalised */
    const double aspect = (double)W / (double)H;
    for (int y = 0; y < H; ++y) {
        const double yn = 2.0 * (y + 0.5) / H - 1.0;
        for (int x = 0; x < W; ++x) {
            const double xn = (2.0 * (x + 0.5) / W - 1.0) * aspect;
            double d[3];
            for (int k = 0; k < 3; ++k) d[k] = f[k] + xn * r[k] - yn * u[k];
            double best = INFINITY;
            int id = -1;
            if (d[2] < -1e-12) { double t = -cam[2] / d[2]; if (t > 0) { best = t; id = 0; } }
            for (int o = 0; o < e->n_obst; ++o) {
                double t = ray_cylinder(cam, d, e->obst[o], c->obst_radius);
                if (t < best) { best = t; id = 2 + o; }
            }
            if (cam_task) {
                double t = ray_sphere(cam, d, ctr, Rd);
                if (t < best) { best = t; id = 1; }
            }
            if (wp_task)
                for (int k = 0; k < e->n_remaining; ++k) {
                    double t = ray_sphere(cam, d, e->targets[k], c->goal_reach);
                    if (t < best) { best = t; id = 64 + e->target_idx + k; }
                }
            /* colour */
            double base[3] = {135, 206, 235}, shade = 1.0;
            if (id >= 0) {
                double hit[3] = {cam[0] + best * d[0], cam[1] + best * d[1], cam[2] + best * d[2]}, n[3] = {0, 0, 1};
                if (id == 0) {
                    /* 10 m checker inside the far plane, one tone beyond it */
                    int chk = ((long long)floor(hit[0] / 10.0) + (long long)floor(hit[1] / 10.0)) & 1;
                    int nearg = best <= c->cam_far;
                    base[0] = nearg ? (chk ? 96 : 80) : 88; base[1] = nearg ? (chk ? 160 : 140) : 150; base[2] = nearg ? (chk ? 96 : 80) : 88;
                } else if (id == 1) {
                    base[0] = 255; base[1] = 221; base[2] = 0;
                    for (int k = 0; k < 3; ++k) n[k] = (hit[k] - ctr[k]) / Rd;
                } else if (id < 64) {
                    const double* cy = e->obst[id - 2];
                    base[0] = 170; base[1] = 90; base[2] = 70;
                    double dx = hit[0] - cy[0], dy = hit[1] - cy[1], rr = c->obst_radius;
                    if (dx * dx + dy * dy >= rr * rr * (1.0 - 1e-3)) { n[0] = dx / rr; n[1] = dy / rr; n[2] = 0.0; }
                } else {
                    const double* tg = e->targets[id - 64 - e->target_idx];
                    int cur = id - 64 == e->target_idx;
                    base[0] = 60; base[1] = cur ? 220 : 120; base[2] = cur ? 60 : 255;
                    for (int k = 0; k < 3; ++k) n[k] = (hit[k] - tg[k]) / c->goal_reach;
                }
                double nl = dot3(n, L);
                shade = 0.55 + 0.45 * (nl > 0.0 ? nl : 0.0);
            }
            const size_t px = (size_t)y * W + x;
            if (rgba) {
                for (int k = 0; k < 3; ++k) rgba[4 * px + k] = (uint8_t)(base[k] * shade + 0.5);
                rgba[4 * px + 3] = 255;
            }
            if (seg) seg[px] = id;
            if (depth) depth[px] = isfinite(best) ? depth_buf(c, best) : 1.0;
        }
    }
}

/* _compute_vision_features + _build_vision_vector */
static void vision_features(const fwo_config* c, fwo_env* e) {
    (void)c;
    double visible = 0.0, dl = 0.0, dc = 0.0, dr = 0.0;
    if (e->cam_valid) {
        dl = e->frame_dl; dc = e->frame_dc; dr = e->frame_dr;
        if (!e->frame_visible) {
            e->steps_since_seen = e->steps_since_seen + 1 < 60 ? e->steps_since_seen + 1 : 60;
        } else {
            e->last_cx = e->frame_cx; e->last_cy = e->frame_cy; e->last_area = e->frame_area;
            e->last_depth = e->frame_depth;
            e->steps_since_seen = 0;
            visible = 1.0;
        }
    }
    /* the reference emits float32 */
    e->vision[0] = visible;
    e->vision[1] = (double)(float)e->last_cx;
    e->vision[2] = (double)(float)e->last_cy;
    e->vision[3] = (double)(float)e->last_area;
    e->vision[4] = (double)(float)e->last_depth;
    e->vision[5] = (double)(float)((double)e->steps_since_seen / 60.0);
    e->vision[6] = (double)(float)dl; e->vision[7] = (double)(float)dc; e->vision[8] = (double)(float)dr;
}

/* ------------------------------------------------------------------ env glue */

static void aviary_step(const fwo_config* c, fwo_env* e, uint64_t seed) {
    e->contact = 0;                                   /* contact_array &= False */
    for (int i = 0; i < c->substeps_per_inner; ++i) fwo_substep(c, e, seed);
    /* drone.update_last(): camera */
    if ((c->task == FWO_TASK_OBJLOCK || c->task == FWO_TASK_DUCK) && c->cam_interval_substeps > 0 &&
        e->physics_steps % c->cam_interval_substeps == 0)
        capture_frame(c, e);
}

/* FixedwingObjLockEnv._build_duck_vision_observation (fixedwing_objlock_env.py:421-459): push the newest nine
 * features into the history (row 0 = newest) and derive the four frame-to-frame deltas */
static void duck_push_history(const fwo_config* c, fwo_env* e) {
    const int H = c->vision_hist_len;
    double prev[9];
    for (int k = 0; k < 9; ++k) prev[k] = H >= 2 ? e->vis_hist[0][k] : 0.0;
    for (int h = H - 1; h >= 1; --h)
        for (int k = 0; k < 9; ++k) e->vis_hist[h][k] = e->vis_hist[h - 1][k];
    for (int k = 0; k < 9; ++k) e->vis_hist[0][k] = e->vision[k];
    e->hist_filled = e->hist_filled + 1 < H ? e->hist_filled + 1 : H;
    for (int k = 0; k < 4; ++k) e->vis_deltas[k] = 0.0;
    if (e->hist_filled >= 2 && e->vision[0] > 0.5 && prev[0] > 0.5)
        for (int k = 0; k < 4; ++k) e->vis_deltas[k] = (double)(float)(e->vision[1 + k] - prev[1 + k]);
}

/* FixedwingObjLockEnv.compute_state (fixedwing_objlock_env.py:253-287) + FlattenObjLockEnv._flatten_obs
 * (flatten_objlock_env.py:41-46).  update != 0: a live compute_state (advances the vision state and the history);
 * update == 0: re-emit the observation of the current state. */
static void duck_compute_obs(const fwo_config* c, fwo_env* e, double* obs, int update) {
    double R[9], rpy[3], q2[4], R2[9], wb[3], vb[3];
    fwo_quat_to_mat(e->quat, R);
    fwo_quat_to_euler(e->quat, rpy);
    fwo_euler_to_quat(rpy, q2);
    fwo_quat_to_mat(q2, R2);
    matT_vec(R, e->omega, wb);
    matT_vec(R, e->vel, vb);
    double d[3] = {e->duck_pos[0] - e->pos[0], e->duck_pos[1] - e->pos[1], e->duck_pos[2] - e->pos[2]};
    matT_vec(R2, d, e->target_vec);
    if (update) {
        vision_features(c, e);
        duck_push_history(c, e);
    }
    if (!obs) return;
    int k = 0;
    for (int i = 0; i < 3; ++i) obs[k++] = wb[i];
    if (c->angle_repr == 0) { for (int i = 0; i < 3; ++i) obs[k++] = rpy[i]; }
    else { for (int i = 0; i < 4; ++i) obs[k++] = q2[i]; }
    for (int i = 0; i < 3; ++i) obs[k++] = vb[i];
    for (int i = 0; i < 3; ++i) obs[k++] = e->pos[i];
    for (int i = 0; i < 4; ++i) obs[k++] = e->last_action[i];
    for (int i = 0; i < FWO_NSURF; ++i) obs[k++] = e->act[i];
    obs[k++] = e->throttle;
    for (int i = 0; i < 3; ++i) obs[k++] = e->target_vec[i];
    for (int h = 0; h < c->vision_hist_len; ++h)
        for (int i = 0; i < 9; ++i) obs[k++] = e->vis_hist[h][i];
    if (c->vision_use_deltas)
        for (int i = 0; i < 4; ++i) obs[k++] = e->vis_deltas[i];
}

void fwo_compute_obs(const fwo_config* c, fwo_env* e, double* obs, int update_dist) {
    if (c->task == FWO_TASK_DUCK) { duck_compute_obs(c, e, obs, update_dist); return; }
    double R[9], rpy[3], q2[4], R2[9];
    fwo_quat_to_mat(e->quat, R);
    fwo_quat_to_euler(e->quat, rpy);
    fwo_euler_to_quat(rpy, q2);               /* compute_attitude: quaternion = getQuaternionFromEuler(ang_pos) */
    fwo_quat_to_mat(q2, R2);
    double wb[3], vb[3];
    matT_vec(R, e->omega, wb);
    matT_vec(R, e->vel, vb);

    /* waypoints.distance_to_targets: (targets - pos) @ R ; old <- new ; new <- |row 0| */
    double rows[FWO_MAX_TARGETS + 1][3];
    int nrows = 0;
    if (c->task != FWO_TASK_PHYSICS) {
        for (int i = 0; i < e->n_remaining; ++i) {
            double d[3] = {e->targets[i][0] - e->pos[0], e->targets[i][1] - e->pos[1], e->targets[i][2] - e->pos[2]};
            matT_vec(R2, d, rows[nrows]);
            nrows++;
        }
        if (e->n_remaining > 0 && update_dist) {
            e->old_dist = e->new_dist;
            e->new_dist = sqrt(dot3(rows[0], rows[0]));
        }
        if (c->task == FWO_TASK_OBJLOCK) {
            double d[3] = {e->duck_pos[0] - e->pos[0], e->duck_pos[1] - e->pos[1], e->duck_pos[2] - e->pos[2]};
            matT_vec(R2, d, rows[nrows]);
            nrows++;
            if (update_dist) {
                vision_features(c, e);
                if (e->n_remaining == 0) {
                    e->post_waypoints = 1;
                    if (!e->duck_phase) {
                        int visible = (e->vision[0] > 0.5) && (e->last_area >= c->switch_min_area);
                        e->seen_consecutive = visible ? e->seen_consecutive + 1 : 0;
                        if (e->seen_consecutive >= c->switch_min_seen) e->duck_phase = 1;
                    }
                } else {
                    e->post_waypoints = 0;
                    e->duck_phase = 0;
                }
            }
        }
    }
    if (!obs || c->task == FWO_TASK_PHYSICS) return;
    int k = 0;
    for (int i = 0; i < 3; ++i) obs[k++] = wb[i];
    if (c->angle_repr == 0) { for (int i = 0; i < 3; ++i) obs[k++] = rpy[i]; }
    else { for (int i = 0; i < 4; ++i) obs[k++] = q2[i]; }
    for (int i = 0; i < 3; ++i) obs[k++] = vb[i];
    for (int i = 0; i < 3; ++i) obs[k++] = e->pos[i];
    for (int i = 0; i < 4; ++i) obs[k++] = e->last_action[i];
    for (int i = 0; i < FWO_NSURF; ++i) obs[k++] = e->act[i];
    obs[k++] = e->throttle;
    for (int r = 0; r < c->context_len; ++r)
        for (int i = 0; i < 3; ++i) obs[k++] = (r < nrows) ? rows[r][i] : 0.0;
}

static void advance_targets(fwo_env* e) {
    for (int i = 1; i < e->n_remaining; ++i)
        for (int k = 0; k < 3; ++k) e->targets[i - 1][k] = e->targets[i][k];
    e->n_remaining -= 1;
    e->target_idx += 1;
}

static void obstacle_penalty(const fwo_config* c, fwo_env* e, int is_duck_phase) {
    double dmin = INFINITY;
    for (int i = 6; i < 9; ++i) {
        double d = e->vision[i];
        if (d > 0.0 && isfinite(d) && d < dmin) dmin = d;
    }
    if (!isfinite(dmin)) return;
    double dsafe = c->obst_safe;
    if (dsafe <= 0.0 || dmin >= dsafe) return;
    double scale = c->obst_scale * (is_duck_phase ? 0.5 : 1.0);
    double pen = scale * (dsafe - dmin) / dsafe;
    if (pen > c->obst_max_pen) pen = c->obst_max_pen;
    e->reward -= pen;
}

/* FixedwingObjLockEnv.compute_term_trunc_reward after the base flags (fixedwing_objlock_env.py:293-372) */
static void duck_reward(const fwo_config* c, fwo_env* e) {
    if (e->info_collision || e->info_oob) return;                   /* :293-294 */
    obstacle_penalty(c, e, 1);                                      /* :376-407, always the halved scale */
    const double dist = sqrt(dot3(e->target_vec, e->target_vec));
    if (!c->sparse_reward) {
        e->reward += c->duck_dist_scale / fmax(dist, 2.0);
        /* the reward reads the newest history row, which is the float32 feature vector of this compute_state */
        const double* v = e->vis_hist[0];
        if (v[0] > 0.5) {
            const double cx = v[1], cy = v[2], area = v[3], est = v[4];
            e->reward += c->visible_step_reward;
            e->reward += c->area_reward_scale * fmax(0.0, area);
            const double dc = sqrt((cx - 0.5) * (cx - 0.5) + (cy - 0.5) * (cy - 0.5));
            const double rl = fmax(c->lock_center_radius, 1e-6);
            e->reward += c->centering_scale * fmax(0.0, (rl - dc) / rl);
            if (dc < rl) {
                e->lock_steps = e->lock_steps + 1 < c->lock_hold_steps ? e->lock_steps + 1 : c->lock_hold_steps;
                e->reward += c->lock_step_reward;
            } else {
                e->lock_steps = e->lock_steps - c->lock_decay_steps > 0 ? e->lock_steps - c->lock_decay_steps : 0;
            }
            if (e->has_prev_dist && est > 0.0 && isfinite(est)) {
                double diff = e->prev_est_dist - est;
                if (c->approach_clip > 0.0) diff = diff < -c->approach_clip ? -c->approach_clip : (diff > c->approach_clip ? c->approach_clip : diff);
                e->reward += diff * c->approach_scale;
            }
            if (est > 0.0 && isfinite(est)) { e->prev_est_dist = est; e->has_prev_dist = 1; }
            else { e->prev_est_dist = 0.0; e->has_prev_dist = 0; }
        } else {
            if (e->lock_steps > 0) e->reward -= c->lock_lost_penalty;
            e->lock_steps = e->lock_steps - c->lock_decay_steps > 0 ? e->lock_steps - c->lock_decay_steps : 0;
            e->prev_est_dist = 0.0; e->has_prev_dist = 0;
        }
    }
    if (e->lock_steps >= c->lock_hold_steps && dist <= c->strike_dist) {           /* :366-372 */
        e->termination = 1;
        e->reward += c->strike_reward;
        e->info_complete = 1;
        e->info_strike = 1;
    }
}

static void term_trunc_reward(const fwo_config* c, fwo_env* e) {
    /* compute_base_term_trunc_reward */
    if (e->step_count > c->max_steps) e->truncation = 1;
    if (e->contact) { e->reward = -100.0; e->info_collision = 1; e->termination = 1; }
    double p2 = dot3(e->pos, e->pos);
    if (sqrt(p2) > c->dome) { e->reward = -100.0; e->info_oob = 1; e->termination = 1; }
    if (c->task == FWO_TASK_PHYSICS) return;
    if (c->task == FWO_TASK_DUCK) { duck_reward(c, e); return; }
    if (c->early_return_on_crash && (e->info_collision || e->info_oob)) return;

    if (c->task == FWO_TASK_WAYPOINTS) {
        /* upstream always truncates on the last waypoint, so an empty target list is never scored there;
         * with complete_truncates == 0 (non-reference option) the waypoint terms simply stop */
        if (e->n_remaining == 0) return;
        if (!c->sparse_reward) {
            double prog = (isinf(e->old_dist) || isinf(e->new_dist)) ? 0.0 : e->old_dist - e->new_dist;
            e->reward += fmax(3.0 * prog, 0.0);
            e->reward += 1.0 / e->new_dist;
        }
        if (e->new_dist < c->goal_reach) {
            e->reward = 100.0;
            advance_targets(e);
            int all = e->n_remaining == 0;
            if (c->complete_truncates && all) e->truncation = 1;
            e->info_complete = all;
            e->num_targets_reached = c->num_targets - e->n_remaining;
        }
        return;
    }
    /* FWO_TASK_OBJLOCK: fixedwing_waypoint_objlock_env.py:285-343 */
    if (e->n_remaining > 0) {
        if (!c->sparse_reward) {
            double prog = (isinf(e->old_dist) || isinf(e->new_dist)) ? 0.0 : e->old_dist - e->new_dist;
            e->reward += fmax(3.0 * prog, 0.0);
            e->reward += 1.0 / e->new_dist;
        }
        if (e->new_dist < c->goal_reach) {
            e->reward = 100.0;
            advance_targets(e);
            e->num_targets_reached = c->num_targets - e->n_remaining;
            if (e->n_remaining == 0) { e->termination = 0; e->truncation = 0; }
        }
        obstacle_penalty(c, e, 0);
    } else {
        e->termination = 0;
        obstacle_penalty(c, e, 1);
        if (e->duck_phase) {
            if (!c->sparse_reward && e->last_depth > 0.0) e->reward += 1.0 / fmax(e->last_depth, 2.0);
            if (e->last_cx > 0.0) {
                double dx = e->last_cx - 0.5, dy = e->last_cy - 0.5;
                if (sqrt(dx * dx + dy * dy) < 0.35) { e->lock_steps += 1; e->reward += c->lock_step_reward; }
                else e->lock_steps = 0;
            } else e->lock_steps = 0;
            double est = e->last_depth;
            if (e->has_prev_dist && est > 0.0) {
                double diff = e->prev_est_dist - est;
                if (diff > 0.0) e->reward += diff * c->approach_scale;
            }
            e->prev_est_dist = est; e->has_prev_dist = 1;
            if (e->lock_steps >= c->lock_hold_steps && est > 0.0 && est <= c->strike_dist) {
                e->termination = 1;
                e->reward += c->strike_reward;
                e->info_complete = 1;
                e->info_strike = 1;
            }
        }
    }
}

static void sample_wind(const fwo_config* c, fwo_env* e, uint64_t seed) {
    for (int k = 0; k < 3; ++k) { e->wind_base[k] = c->wind_base[k]; e->gust_amp[k] = c->gust_amp[k]; }
    e->gust_phase = c->gust_phase;
    if (c->wind_mode == 0 || !c->wind_randomize) return;
    uint32_t r0[4], r1[4];
    fwo_philox(seed, e->env_id, e->episode, 0u, STREAM_WIND, r0);
    fwo_philox(seed, e->env_id, e->episode, 1u, STREAM_WIND, r1);
    for (int k = 0; k < 3; ++k)
        e->wind_base[k] = c->wind_base_lo[k] + fwo_u01(r0[k]) * (c->wind_base_hi[k] - c->wind_base_lo[k]);
    if (c->wind_mode == 2) {
        for (int k = 0; k < 3; ++k)
            e->gust_amp[k] = c->gust_amp_lo[k] + fwo_u01(r1[k]) * (c->gust_amp_hi[k] - c->gust_amp_lo[k]);
        if (c->wind_rand_phase) e->gust_phase = 2.0 * M_PI * fwo_u01(r0[3]);
    }
}

static void sample_targets(const fwo_config* c, fwo_env* e, uint64_t seed) {
    /* [UP-RECALL] WaypointHandler.reset: polar sampling */
    for (int i = 0; i < c->num_targets; ++i) {
        uint32_t r[4];
        fwo_philox(seed, e->env_id, e->episode, (uint32_t)i, STREAM_TARGETS, r);
        double theta = 2.0 * M_PI * fwo_u01(r[0]);
        double phi = 2.0 * M_PI * fwo_u01(r[1]);
        double dist = 1.0 + fwo_u01(r[2]) * (c->spawn_size * 0.9 - 1.0);
        double x = dist * sin(phi) * cos(theta);
        double y = dist * sin(phi) * sin(theta);
        double z = fabs(dist * cos(phi));
        e->targets[i][0] = x; e->targets[i][1] = y;
        e->targets[i][2] = z > c->min_height ? z : c->min_height;
    }
    e->n_remaining = c->num_targets;
    e->target_idx = 0;
    e->old_dist = INFINITY; e->new_dist = INFINITY;
}

static void spawn_duck_obstacles(const fwo_config* c, fwo_env* e, uint64_t seed) {
    /* _reset_duck_phase_state / _spawn_duck / _spawn_obstacles: objlock_env.py:382-519 */
    e->duck_phase = 0; e->seen_consecutive = 0; e->lock_steps = 0; e->has_prev_dist = 0; e->prev_est_dist = 0.0;
    e->last_cx = 0.5; e->last_cy = 0.5; e->last_area = 0.0; e->last_depth = 0.0;
    e->steps_since_seen = 60; e->post_waypoints = 0; e->cam_valid = 0; e->frame_visible = 0;
    memset(e->vision, 0, sizeof(e->vision));
    if (c->num_targets > 0) {
        e->duck_pos[0] = e->targets[c->num_targets - 1][0];
        e->duck_pos[1] = e->targets[c->num_targets - 1][1];
    } else { e->duck_pos[0] = 10.0; e->duck_pos[1] = 0.0; }
    e->duck_pos[2] = 0.05;
    e->n_obst = 0;
    for (int i = 0; i < c->num_obstacles; ++i) {
        uint32_t r[4];
        fwo_philox(seed, e->env_id, e->episode, (uint32_t)i, STREAM_OBST, r);
        double h = c->obst_h_lo + fwo_u01(r[0]) * (c->obst_h_hi - c->obst_h_lo);
        double x = -c->dome / 2 + fwo_u01(r[1]) * c->dome;
        double y = -c->dome / 2 + fwo_u01(r[2]) * c->dome;
        if (x * x + y * y < 100.0) continue;
        e->obst[e->n_obst][0] = x; e->obst[e->n_obst][1] = y; e->obst[e->n_obst][2] = h;
        e->n_obst++;
    }
}

/* FixedwingObjLockEnv._reset_duck_state / _spawn_duck / _spawn_obstacles (fixedwing_objlock_env.py:409-419,461-578) */
static void spawn_duck_only(const fwo_config* c, fwo_env* e, uint64_t seed) {
    e->duck_phase = 0; e->seen_consecutive = 0; e->lock_steps = 0; e->has_prev_dist = 0; e->prev_est_dist = 0.0;
    e->last_cx = 0.5; e->last_cy = 0.5; e->last_area = 0.0; e->last_depth = 0.0;
    e->steps_since_seen = 60; e->post_waypoints = 0; e->cam_valid = 0; e->frame_visible = 0;
    memset(e->vision, 0, sizeof(e->vision));
    memset(e->vis_hist, 0, sizeof(e->vis_hist));
    memset(e->vis_deltas, 0, sizeof(e->vis_deltas));
    e->hist_filled = 0;
    e->n_remaining = 0; e->target_idx = 0;
    e->old_dist = INFINITY; e->new_dist = INFINITY;
    uint32_t r[4];
    fwo_philox(seed, e->env_id, e->episode, 0u, STREAM_DUCK, r);       /* x, y ~ U(-dome/2, dome/2); yaw unused (sphere) */
    e->duck_pos[0] = -c->dome / 2 + fwo_u01(r[0]) * c->dome;
    e->duck_pos[1] = -c->dome / 2 + fwo_u01(r[1]) * c->dome;
    e->duck_pos[2] = 0.05;
    e->n_obst = 0;
    for (int i = 0; i < c->num_obstacles; ++i) {
        fwo_philox(seed, e->env_id, e->episode, (uint32_t)i, STREAM_OBST, r);
        double h = c->obst_h_lo + fwo_u01(r[0]) * (c->obst_h_hi - c->obst_h_lo);
        double x = -c->dome / 2 + fwo_u01(r[1]) * c->dome;
        double y = -c->dome / 2 + fwo_u01(r[2]) * c->dome;
        double dx = x - e->duck_pos[0], dy = y - e->duck_pos[1];
        if (sqrt(dx * dx + dy * dy) < 10.0) continue;                  /* :536-539 */
        if (x * x + y * y < 100.0) continue;                           /* :542-543 */
        e->obst[e->n_obst][0] = x; e->obst[e->n_obst][1] = y; e->obst[e->n_obst][2] = h;
        e->n_obst++;
    }
}

/* FixedwingLowLevelEnv._compute_obs (fixedwing_lowlevel_env.py:143-156): Aviary.state(0) flattened
 * [ang_vel_body, euler, lin_vel_body, lin_pos] + previous action (6) + target (3) */
static void lowlevel_obs(const fwo_config* c, fwo_env* e, double* obs) {
    (void)c;
    if (!obs) return;
    double R[9], rpy[3], wb[3], vb[3];
    fwo_quat_to_mat(e->quat, R);
    fwo_quat_to_euler(e->quat, rpy);
    matT_vec(R, e->omega, wb);
    matT_vec(R, e->vel, vb);
    int k = 0;
    for (int i = 0; i < 3; ++i) obs[k++] = wb[i];
    for (int i = 0; i < 3; ++i) obs[k++] = rpy[i];
    for (int i = 0; i < 3; ++i) obs[k++] = vb[i];
    for (int i = 0; i < 3; ++i) obs[k++] = e->pos[i];
    for (int i = 0; i < 6; ++i) obs[k++] = e->last_action[i];
    for (int i = 0; i < 3; ++i) obs[k++] = e->target_ref[i];
}

static double wrap_pi(double a) {            /* FixedwingLowLevelEnv._wrap_pi: Python's % is the floored modulo */
    double m = fmod(a + M_PI, 2.0 * M_PI);
    if (m < 0.0) m += 2.0 * M_PI;
    return m - M_PI;
}

/* FixedwingLowLevelEnv.step (fixedwing_lowlevel_env.py:97-141): one Aviary.step per env step */
static void lowlevel_step(const fwo_config* c, fwo_env* e, uint64_t seed, const double* action, double* obs) {
    e->step_count += 1;                                          /* self._episode_steps += 1 */
    for (int k = 0; k < 6; ++k) { e->last_action[k] = action[k]; e->setpoint[k] = action[k]; }
    aviary_step(c, e, seed);
    lowlevel_obs(c, e, obs);
    double R[9], rpy[3], vb[3];
    fwo_quat_to_mat(e->quat, R);
    fwo_quat_to_euler(e->quat, rpy);
    matT_vec(R, e->vel, vb);
    const double speed = sqrt(dot3(vb, vb)), alt = e->pos[2];
    const double psi_err = fabs(wrap_pi(e->target_ref[0] - rpy[2]));
    const double h_err = fabs(e->target_ref[1] - alt), v_err = fabs(e->target_ref[2] - speed);
    e->reward = -(1.0 * psi_err + 1.0 * h_err + 0.5 * v_err) + 0.1;
    e->termination = 0; e->truncation = 0; e->info_oob = 0;
    if (alt < 1.0 || alt > 100.0) { e->termination = 1; e->reward -= 100.0; e->info_oob = 1; }
    if (e->step_count >= c->max_steps) e->truncation = 1;        /* `>= 2000` */
}

void fwo_reset(const fwo_config* c, fwo_env* e, uint64_t seed, uint32_t env_id, uint32_t episode, double* obs) {
    memset(e, 0, sizeof(*e));
    e->env_id = env_id; e->episode = episode;
    /* begin_reset + Aviary(): fixedwing_base_env.py:200-237 */
    for (int k = 0; k < 3; ++k) { e->pos[k] = c->start_pos[k]; e->vel[k] = c->start_vel[k]; }
    e->quat[3] = 1.0;
    fwo_refresh_surface_vel(c, e, -1);                 /* Aviary.reset: drone.update_state() before any wind */
    sample_wind(c, e, seed);                           /* _maybe_apply_wind_field */
    if (c->task == FWO_TASK_LOWLEVEL) {
        /* fixedwing_lowlevel_env.py:87-91: psi_ref ~ U(-pi, pi), h_ref ~ U(5, 20), V_ref ~ U(10, 20) */
        uint32_t r[4];
        fwo_philox(seed, env_id, episode, 0u, STREAM_TARGETS, r);
        e->target_ref[0] = -M_PI + 2.0 * M_PI * fwo_u01(r[0]);
        e->target_ref[1] = 5.0 + 15.0 * fwo_u01(r[1]);
        e->target_ref[2] = 10.0 + 10.0 * fwo_u01(r[2]);
    } else if (c->task == FWO_TASK_DUCK) spawn_duck_only(c, e, seed);
    else if (c->task != FWO_TASK_PHYSICS) sample_targets(c, e, seed);
    if (c->task == FWO_TASK_OBJLOCK) spawn_duck_obstacles(c, e, seed);
    /* end_reset: set_mode(0) -> zero setpoint; 10 x Aviary.step; compute_state */
    for (int i = 0; i < c->warmup_inner; ++i) aviary_step(c, e, seed);
    if (c->task == FWO_TASK_LOWLEVEL) lowlevel_obs(c, e, obs);
    else fwo_compute_obs(c, e, obs, 1);
}

void fwo_step(const fwo_config* c, fwo_env* e, uint64_t seed, const double* action,
              double* obs, double* reward, int32_t* flags) {
    if (c->task == FWO_TASK_LOWLEVEL) {
        lowlevel_step(c, e, seed, action, obs);
        e->ep_return += e->reward;
        e->ep_length += 1;
        if (reward) *reward = e->reward;
        if (flags) *flags = (e->termination ? FWO_TERM : 0) | (e->truncation ? FWO_TRUNC : 0) | (e->info_oob ? FWO_OOB : 0);
        return;
    }
    e->reward = -0.1;
    for (int k = 0; k < 4; ++k) { e->last_action[k] = action[k]; e->setpoint[k] = action[k]; }
    e->setpoint[3] = action[3] / 2.0 + 0.5;
    for (int it = 0; it < c->inner_per_step; ++it) {
        if (e->termination || e->truncation) break;
        aviary_step(c, e, seed);
        fwo_compute_obs(c, e, obs, 1);
        term_trunc_reward(c, e);
    }
    e->step_count += 1;
    e->ep_return += e->reward;
    e->ep_length += 1;
    if (reward) *reward = e->reward;
    if (flags) {
        *flags = (e->termination ? FWO_TERM : 0) | (e->truncation ? FWO_TRUNC : 0) |
                 (e->info_collision ? FWO_COLLISION : 0) | (e->info_oob ? FWO_OOB : 0) |
                 (e->info_complete ? FWO_COMPLETE : 0) | (e->info_strike ? FWO_STRIKE : 0);
    }
}

/* ---- tiny pthread parallel-for (the image's alternate gcc has no libgomp) ---- */
#include <pthread.h>

typedef struct {
    const fwo_config* c; fwo_env* envs; int lo, hi; uint64_t seed; uint32_t env_id0;
    const double* actions; double* obs; double* rewards; int32_t* flags; double* term_obs;
    int32_t* targets_reached;   /* info["num_targets_reached"] of the step, before the worker's reset-on-done */
    int steps; int mode;
} fwo_job;

static void run_range(fwo_job* j) {
    const fwo_config* c = j->c;
    int D = fwo_obs_dim(c);
    for (int i = j->lo; i < j->hi; ++i) {
        fwo_env* e = &j->envs[i];
        if (j->mode == 0) {
            fwo_reset(c, e, j->seed, j->env_id0 + (uint32_t)i, 0u, j->obs ? j->obs + (size_t)i * D : 0);
        } else if (j->mode == 1) {
            double tmp[FWO_MAX_OBS];
            double* o = j->obs ? j->obs + (size_t)i * D : tmp;
            int32_t fl = 0;
            double r = 0;
            fwo_step(c, e, j->seed, j->actions + (size_t)i * fwo_act_dim(c), o, &r, &fl);
            if (j->rewards) j->rewards[i] = r;
            if (j->flags) j->flags[i] = fl;
            /* WaypointHandler.num_targets_reached = n - len(targets) (fixedwing_waypoint_objlock_env.py:296) */
            if (j->targets_reached) j->targets_reached[i] = c->num_targets > 0 ? c->num_targets - e->n_remaining : 0;
            if (fl & (FWO_TERM | FWO_TRUNC)) {
                /* SubprocVecEnv worker: info["terminal_observation"] = obs; obs = env.reset() */
                if (j->term_obs) memcpy(j->term_obs + (size_t)i * D, o, sizeof(double) * D);
                uint32_t ep = e->episode + 1, id = e->env_id;
                fwo_reset(c, e, j->seed, id, ep, o);
            }
        } else {
            double obs[FWO_MAX_OBS], a[FWO_MAX_ACT], r;
            int32_t fl;
            for (int s = 0; s < j->steps; ++s) {
                if (fwo_act_dim(c) == 6) fwo_random_action6(j->seed, e->env_id, e->episode, (uint32_t)e->step_count, a);
                else fwo_random_action(j->seed, e->env_id, e->episode, (uint32_t)e->step_count, a);
                fwo_step(c, e, j->seed, a, obs, &r, &fl);
                if (fl & (FWO_TERM | FWO_TRUNC)) {
                    uint32_t ep = e->episode + 1, id = e->env_id;
                    fwo_reset(c, e, j->seed, id, ep, obs);
                }
            }
        }
    }
}
static void* run_thread(void* p) { run_range((fwo_job*)p); return 0; }

static void parallel_run(fwo_job* proto, int n, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    if (nthreads == 1) { proto->lo = 0; proto->hi = n; run_range(proto); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    fwo_job* jobs = (fwo_job*)malloc(sizeof(fwo_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].lo = (int)((long)n * t / nthreads);
        jobs[t].hi = (int)((long)n * (t + 1) / nthreads);
        pthread_create(&th[t], 0, run_thread, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    free(th); free(jobs);
}

void fwo_vec_reset(const fwo_config* c, fwo_env* envs, int n, uint64_t seed, uint32_t env_id0,
                   double* obs, int nthreads) {
    fwo_job j; memset(&j, 0, sizeof(j));
    j.c = c; j.envs = envs; j.seed = seed; j.env_id0 = env_id0; j.obs = obs; j.mode = 0;
    parallel_run(&j, n, nthreads);
}

void fwo_vec_step(const fwo_config* c, fwo_env* envs, int n, uint64_t seed, const double* actions,
                  double* obs, double* rewards, int32_t* flags, double* term_obs, int nthreads) {
    fwo_job j; memset(&j, 0, sizeof(j));
    j.c = c; j.envs = envs; j.seed = seed; j.actions = actions; j.obs = obs; j.rewards = rewards;
    j.flags = flags; j.term_obs = term_obs; j.mode = 1;
    parallel_run(&j, n, nthreads);
}

void fwo_vec_step_info(const fwo_config* c, fwo_env* envs, int n, uint64_t seed, const double* actions,
                       double* obs, double* rewards, int32_t* flags, double* term_obs, int32_t* targets_reached,
                       int nthreads) {
    fwo_job j; memset(&j, 0, sizeof(j));
    j.c = c; j.envs = envs; j.seed = seed; j.actions = actions; j.obs = obs; j.rewards = rewards;
    j.flags = flags; j.term_obs = term_obs; j.targets_reached = targets_reached; j.mode = 1;
    parallel_run(&j, n, nthreads);
}

long fwo_rollout_random(const fwo_config* c, fwo_env* envs, int n, uint64_t seed, int steps, int nthreads) {
    fwo_job j; memset(&j, 0, sizeof(j));
    j.c = c; j.envs = envs; j.seed = seed; j.steps = steps; j.mode = 2;
    parallel_run(&j, n, nthreads);
    return (long)n * steps;
}
