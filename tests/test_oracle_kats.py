"""CPU tests pinning the fp64 oracle.

The reference has no tests or golden vectors for this path, so the oracle is pinned by
the closed forms and constant known-answer values of SURVEY.md section 8(c): Philox4x32-10 vectors from the
Random123 distribution, aero constants derived from /root/reference/my_models/fixedwing/fixewing.yaml,
main-wing force KATs, actuator/throttle closed forms, free fall under symplectic Euler, torque-free spin,
quaternion norm, wind field, waypoint geometry.  "parity unpinned" against PyFlyt itself.
"""
import ctypes as C
import math

import numpy as np
import pytest

import pyflyt_drone_b200 as fw


@pytest.fixture(scope="module")
def fo(oracle_mod):
    return oracle_mod


@pytest.fixture(scope="module")
def cfg():
    return fw.waypoints_v3()


def _coeffs(fo, oc, s, alpha, act=0.0):
    out = (C.c_double * 3)()
    fo.lib().fwo_aero_coeffs(C.byref(oc), s, alpha, act, out)
    return np.array(out[:])


def test_struct_sizes_match(fo):
    L = fo.lib()
    assert L.fwo_config_size() == C.sizeof(fo.OConfig)
    assert L.fwo_env_size() == C.sizeof(fo.OEnv)


def test_philox_random123_known_answers(fo):
    L = fo.lib()

    def ph(key, ctr):
        out = (C.c_uint32 * 4)()
        L.fwo_philox(key, *ctr, out)
        return list(out)

    assert ph(0, (0, 0, 0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert ph(0xFFFFFFFFFFFFFFFF, (0xFFFFFFFF,) * 4) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert ph((0x299F31D0 << 32) | 0xA4093822, (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_is_open_interval_and_fp32_exact(fo):
    L = fo.lib()
    for x in (0, 1, 255, 256, 0xFFFFFFFF, 0x80000000):
        u = L.fwo_u01(x)
        assert 0.0 < u < 1.0
        assert float(np.float32(u)) == u


def test_normals_have_unit_moments(fo):
    L = fo.lib()
    out = (C.c_double * 4)()
    xs = []
    for i in range(4000):
        L.fwo_normals4(123, 5, 0, i, out)
        xs += list(out)
    xs = np.array(xs)
    assert abs(xs.mean()) < 0.03 and abs(xs.std() - 1.0) < 0.03


def test_cl_alpha_3d_constants(fo, cfg):
    oc = fo.make_config(cfg.as_dict())
    # surfaces in cmd order: left ail, right ail, h-tail, v-tail, main; KATs of SURVEY 8(c)(1)
    expect = {4: 4.253108, 0: 1.449923, 1: 1.449923, 2: 3.324768, 3: 2.092726}
    for s, cla in expect.items():
        a0 = math.radians(cfg.alpha0_base_deg[s])
        a = a0 + math.radians(3.0)
        cl = _coeffs(fo, oc, s, a)[0]
        assert cl / math.radians(3.0) == pytest.approx(cla, rel=2e-6)


def test_main_wing_force_kats(fo, cfg):
    oc = fo.make_config(cfg.as_dict())
    kats = {0.0: (0.148461, 0.013868, -0.011590), 5.0: (0.519615, 0.057500, -0.044510),
            20.0: (0.611135, 0.206631, -0.071287)}
    for deg, (cl, cd, cm) in kats.items():
        got = _coeffs(fo, oc, 4, math.radians(deg))
        assert got == pytest.approx([cl, cd, cm], abs=1.5e-6)
    # dimensional forces at V = 20 m/s, alpha = 5 deg: F_normal 61.4636, F_parallel -1.4105, torque -1.5703
    a = math.radians(5.0)
    v = (C.c_double * 3)(20 * math.cos(a), 0.0, -20 * math.sin(a))
    f, t = (C.c_double * 3)(), (C.c_double * 3)()
    fo.lib().fwo_surface_force(C.byref(oc), 4, 0.0, v, f, t)
    assert f[2] == pytest.approx(61.4636, abs=1e-3) and f[0] == pytest.approx(-1.4105, abs=1e-3)
    assert t[1] == pytest.approx(-1.5703, abs=1e-3) and t[0] == 0.0 and t[2] == 0.0
    # alpha = 0: L 17.4591 N, D 1.6308 N
    v = (C.c_double * 3)(20.0, 0.0, 0.0)
    fo.lib().fwo_surface_force(C.byref(oc), 4, 0.0, v, f, t)
    assert f[2] == pytest.approx(17.4591, abs=1e-3) and -f[0] == pytest.approx(1.6308, abs=1e-3)
    assert t[1] == pytest.approx(-0.4089, abs=1e-3)


def test_stall_branches(fo, cfg):
    oc = fo.make_config(cfg.as_dict())
    # just inside / just outside the positive stall angle of the main wing: the coefficients jump by
    # construction of the model (attached-flow vs flat-plate branch), but both sides must be finite and
    # the induced angle must be continuous: alpha_i(stall-) == alpha_i(stall+) == Cl_stall / (pi AR)
    lo = _coeffs(fo, oc, 4, math.radians(13.999))
    hi = _coeffs(fo, oc, 4, math.radians(14.001))
    assert np.all(np.isfinite(lo)) and np.all(np.isfinite(hi))
    assert lo[0] == pytest.approx(4.253108 * math.radians(15.999), rel=1e-5)
    # flat plate at 90 deg: Cl ~ 0 (up to the CT term), Cd ~ Cd90 * (1/1.0 - k)
    c90 = _coeffs(fo, oc, 4, math.pi / 2 - 1e-9)
    assert abs(c90[0]) < 0.15 and c90[1] > 1.0
    neg = _coeffs(fo, oc, 4, math.radians(-30.0))
    assert neg[0] < 0 and neg[1] > 0


def test_vertical_tail_lifts_sideways(fo, cfg):
    oc = fo.make_config(cfg.as_dict())
    v = (C.c_double * 3)(20.0, -2.0, 0.0)   # sideslip: air comes from +y
    f, t = (C.c_double * 3)(), (C.c_double * 3)()
    fo.lib().fwo_surface_force(C.byref(oc), 3, 0.0, v, f, t)
    assert f[1] > 0 and f[2] == 0.0 and abs(t[2]) > 0 and t[1] == 0.0


def _one(fo, cfg, **kw):
    c = cfg.replace(**kw) if kw else cfg
    env = fo.OracleVecEnv(c.as_dict(), 1, seed=3)
    env.reset()
    return env


def test_actuator_and_throttle_closed_forms(fo, cfg):
    env = _one(fo, cfg, noise_ratio=0.0)
    e = env.envs[0]
    a0, t0 = list(e.act), e.throttle
    assert a0 == [0.0] * 5 and t0 == 0.0       # zero setpoint during the warm-up
    env.step(np.array([[1.0, -1.0, 0.5, 1.0]]))
    k = (1 / 240) / 0.05
    n = 8
    lag = 1 - (1 - k) ** n
    assert e.act[0] == pytest.approx(cfg.ail_left_sign * lag, rel=1e-12)
    assert e.act[1] == pytest.approx(cfg.ail_right_sign * lag, rel=1e-12)
    assert e.act[2] == pytest.approx(-1.0 * lag, rel=1e-12)
    assert e.act[3] == pytest.approx(0.5 * lag, rel=1e-12)
    assert e.act[4] == 0.0
    km = (1 / 240) / 0.01
    assert e.throttle == pytest.approx(1.0 * (1 - (1 - km) ** n), rel=1e-12)


def test_motor_constants(cfg):
    max_rpm = math.sqrt(cfg.total_thrust / cfg.thrust_coef)
    assert max_rpm == pytest.approx(238667.185, rel=1e-8)
    assert max_rpm ** 2 * cfg.torque_coef == pytest.approx(0.452278, rel=1e-5)
    assert (1 / 240) / cfg.motor_tau == pytest.approx(0.416667, rel=1e-5)


def test_free_fall_symplectic_euler(fo, cfg):
    # no air, no thrust: z_n = z_0 - g dt^2 n(n+1)/2 ; v_n = -g n dt ; no rotation although gravity is
    # applied about the base-link CoM and not the composite CoM (checks the 6x6 formulation)
    c = cfg.replace(rho=0.0, noise_ratio=0.0, start_pos=[0.0, 0.0, 50.0], warmup_inner=0)
    env = fo.OracleVecEnv(c.as_dict(), 1, seed=0)
    env.reset()
    e = env.envs[0]
    env.step(np.array([[0.0, 0.0, 0.0, -1.0]]))
    n, dt, g = 8, 1 / 240, 9.81
    assert e.physics_steps == n
    assert e.pos[2] == pytest.approx(50.0 - g * dt * dt * n * (n + 1) / 2, rel=1e-13)
    assert e.vel[2] == pytest.approx(-g * n * dt, rel=1e-13)
    assert e.pos[0] == pytest.approx(20.0 * n * dt, rel=1e-13)
    assert max(abs(x) for x in e.omega) < 1e-13
    assert list(e.quat)[:3] == pytest.approx([0, 0, 0], abs=1e-13)


def test_torque_free_spin_about_principal_axis_is_stationary(fo, cfg):
    c = cfg.replace(rho=0.0, gravity=0.0, noise_ratio=0.0, com=[0.0, 0.0, 0.0], warmup_inner=0,
                    inertia_o=[0.2, 0, 0, 0, 0.3, 0, 0, 0, 0.4], start_vel=[0.0, 0.0, 0.0], dome=1e9)
    env = fo.OracleVecEnv(c.as_dict(), 1, seed=0)
    env.reset()
    st = env.get_state()
    st["omega"][0] = [0.0, 0.0, 3.0]
    env.set_state(st)
    for _ in range(10):
        env.step(np.array([[0.0, 0.0, 0.0, -1.0]]))
    e = env.envs[0]
    assert list(e.omega) == pytest.approx([0.0, 0.0, 3.0], abs=1e-12)
    q = np.array(e.quat[:])
    assert np.linalg.norm(q) == pytest.approx(1.0, abs=1e-14)
    # yaw advanced by 3 rad/s * 80/240 s
    yaw = 2 * math.atan2(q[2], q[3])
    assert yaw == pytest.approx(3.0 * 80 / 240, abs=1e-12)


def test_intermediate_axis_spin_is_unstable_but_conserves_energy_roughly(fo, cfg):
    c = cfg.replace(rho=0.0, gravity=0.0, noise_ratio=0.0, com=[0.0, 0.0, 0.0], warmup_inner=0,
                    inertia_o=[0.2, 0, 0, 0, 0.3, 0, 0, 0, 0.4], start_vel=[0.0, 0.0, 0.0], dome=1e9)
    env = fo.OracleVecEnv(c.as_dict(), 1, seed=0)
    env.reset()
    st = env.get_state()
    st["omega"][0] = [0.01, 5.0, 0.01]
    env.set_state(st)
    I = np.diag([0.2, 0.3, 0.4])

    def energy():
        e = env.envs[0]
        R = np.zeros(9)
        q = (C.c_double * 4)(*e.quat)
        Rm = (C.c_double * 9)()
        fo.lib().fwo_quat_to_mat(q, Rm)
        R = np.array(Rm[:]).reshape(3, 3)
        wb = R.T @ np.array(e.omega[:])
        return 0.5 * wb @ I @ wb

    e0 = energy()
    for _ in range(60):
        env.step(np.array([[0.0, 0.0, 0.0, -1.0]]))
    assert energy() == pytest.approx(e0, rel=0.05)   # explicit Euler on the gyroscopic term drifts slowly


def test_quaternion_euler_round_trip_and_gimbal_guard(fo):
    L = fo.lib()
    rng = np.random.default_rng(0)
    for _ in range(200):
        rpy = rng.uniform([-3.1, -1.5, -3.1], [3.1, 1.5, 3.1])
        q = (C.c_double * 4)()
        back = (C.c_double * 3)()
        L.fwo_euler_to_quat((C.c_double * 3)(*rpy), q)
        L.fwo_quat_to_euler(q, back)
        assert np.array(back[:]) == pytest.approx(rpy, abs=1e-9)
        assert np.linalg.norm(q[:]) == pytest.approx(1.0, abs=1e-14)
    # pitch = +90 deg trips pybullet's 0.99999 guard: roll forced to 0
    q = (C.c_double * 4)()
    L.fwo_euler_to_quat((C.c_double * 3)(0.3, math.pi / 2, 0.2), q)
    back = (C.c_double * 3)()
    L.fwo_quat_to_euler(q, back)
    assert back[0] == 0.0 and back[1] == pytest.approx(math.pi / 2)


def test_wind_field_gust_sine(fo, cfg):
    wind = dict(enabled=True, mode="gust_sine", wind_enu_mps=[1.0, -2.0, 0.5], gust_amp_enu_mps=[0.5, 0.25, 0.1],
                gust_freq_hz=0.2, gust_phase_rad=0.3, randomize_on_reset=False)
    c = fw.waypoint_objlock(wind=wind).replace(task=1, noise_ratio=0.0)
    env = fo.OracleVecEnv(c.as_dict(), 1, seed=0)
    env.reset()
    e = env.envs[0]
    n = e.physics_steps - 1     # stamp of the last state refresh
    w = np.array([1.0, -2.0, 0.5]) + np.array([0.5, 0.25, 0.1]) * math.sin(2 * math.pi * 0.2 * n / 240 + 0.3)
    Rm = (C.c_double * 9)()
    fo.lib().fwo_quat_to_mat((C.c_double * 4)(*e.quat), Rm)
    R = np.array(Rm[:]).reshape(3, 3)
    r = np.array(c.r_surf[4])
    expect = R.T @ (np.array(e.vel[:]) + np.cross(np.array(e.omega[:]), R @ r) - w)
    assert np.array(e.surf_vel[4][:]) == pytest.approx(expect, abs=1e-12)


def test_wind_randomisation_ranges(fo):
    c = fw.waypoint_objlock().replace(task=1)
    env = fo.OracleVecEnv(c.as_dict(), 64, seed=11)
    env.reset()
    st = env.get_state()
    w = st["wind"]
    assert np.all(np.abs(w[:, 0]) <= 5) and np.all(np.abs(w[:, 1]) <= 5) and np.all(np.abs(w[:, 2]) <= 0.5)
    assert np.all(w[:, 3:5] >= 0) and np.all(w[:, 3:5] <= 3) and np.all(w[:, 5] <= 0.3)
    assert np.all(w[:, 6] >= 0) and np.all(w[:, 6] <= 2 * math.pi)
    assert w[:, 0].std() > 1.0


def test_waypoint_geometry(fo, cfg):
    env = fo.OracleVecEnv(cfg.as_dict(), 128, seed=5)
    env.reset()
    t = env.get_state()["targets"]
    d = np.linalg.norm(t, axis=-1)
    # z is floored at min_height, which can push the norm marginally above the sampled distance
    assert np.all(d <= 0.9 * cfg.dome + cfg.min_height) and np.all(t[..., 2] >= cfg.min_height)
    assert d.min() >= 1.0 - 1e-9
    # a target dead ahead has body-frame delta (+d, 0, 0)
    env1 = fo.OracleVecEnv(cfg.as_dict(), 1, seed=5)
    env1.reset()
    st = env1.get_state()
    e = env1.envs[0]
    Rm = (C.c_double * 9)()
    fo.lib().fwo_quat_to_mat((C.c_double * 4)(*e.quat), Rm)
    R = np.array(Rm[:]).reshape(3, 3)
    st["targets"][0, 0] = st["pos"][0] + 30.0 * R[:, 0]
    env1.set_state(st)
    obs = np.zeros(28)
    fo.lib().fwo_compute_obs(C.byref(env1.cfg), C.byref(e), obs.ctypes.data_as(C.c_void_p), 0)
    assert obs[22:25] == pytest.approx([30.0, 0.0, 0.0], abs=1e-9)
