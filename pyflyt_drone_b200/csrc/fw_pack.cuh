// fw_pack.cuh -- the 240 Hz substep of fw_device.cuh for TWO environments per thread on the packed fp32x2 path of
// sm_100a (FFMA2 / FMUL2 / FADD2: one issue slot, two IEEE fp32 results).
//
// Why: the scalar step kernel is ISSUE-bound, not pipe-bound (ncu, profiles/r2_k1_packed.md: 7,167 warp-instructions per
// env-step, 67 % of them FFMA/FMUL/FADD, issue slots 77 % busy with the FMA pipe at 48 %).  A packed instruction
// occupies the FP32 pipe for two cycles but takes ONE issue slot (scripts/f32x2_probe.cu: same 72-74 TFLOP/s as FFMA
// chains, and +20 % on an FFMA + ALU mix), so pairing two environments in the lanes of every FP32 instruction halves the
// issue cost of the arithmetic AND of everything that is per-thread rather than per-env (constant loads through LDCU,
// loop control): ~4,050 issue slots per env-step instead of 7,167.  The aircraft constants ride as scalar-broadcast
// operands (`UR.F32`), sign flips and |x| as operand modifiers, exactly as in the scalar code.
//
// The arithmetic is fw_substep<STD = true> (standard aircraft layout, verified by derive()) statement by statement, with
// the multiply-add contractions written out; MUFU ops, compares, selects and min/max stay one per environment.
#pragma once

#include "fw_device.cuh"

struct f2 {
    float2 v;
    __device__ __forceinline__ f2() {}
    __device__ __forceinline__ f2(float2 a) : v(a) {}
    __device__ __forceinline__ f2(float a, float b) : v(make_float2(a, b)) {}
    __device__ __forceinline__ explicit f2(float s) : v(make_float2(s, s)) {}
    __device__ __forceinline__ float operator[](int k) const { return k ? v.y : v.x; }
};
struct b2 { bool x, y; };

// Sign flip that ptxas folds into the consumer's operand modifier (`-R.F32x2`).  It must be the plain PTX `neg.f32`:
// under --ftz=true a C-level `-x` becomes `neg.ftz.f32`, which ptxas keeps as a separate FADD.FTZ per lane in front of
// a packed instruction (measured: 448 of 5,891 warp-instructions per 32 env-steps).
__device__ __forceinline__ float fw_neg(float x) { float y; asm("neg.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ f2 operator-(f2 a) { return f2(fw_neg(a.v.x), fw_neg(a.v.y)); }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return f2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return f2(__fadd2_rn(a.v, (-b).v)); }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return f2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ f2 operator+(f2 a, float s) { return a + f2(s); }
__device__ __forceinline__ f2 operator-(f2 a, float s) { return a + f2(-s); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c);
// s - a as a*(-1) + s: FADD2 has no negate modifier next to an immediate operand (two scalar FADDs otherwise)
__device__ __forceinline__ f2 operator-(float s, f2 a) { return fma2(a, f2(-1.0f), f2(s)); }
__device__ __forceinline__ f2 operator*(f2 a, float s) { return a * f2(s); }
__device__ __forceinline__ f2 operator*(float s, f2 a) { return a * f2(s); }
__device__ __forceinline__ f2& operator+=(f2& a, f2 b) { a = a + b; return a; }
__device__ __forceinline__ f2& operator-=(f2& a, f2 b) { a = a - b; return a; }
// a*b + c and friends, single rounding
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return f2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ f2 fma2(f2 a, float s, f2 c) { return fma2(a, f2(s), c); }
__device__ __forceinline__ f2 fma2(float s, f2 a, f2 c) { return fma2(a, f2(s), c); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, float c) { return fma2(a, b, f2(c)); }
__device__ __forceinline__ f2 fma2(f2 a, float s, float c) { return fma2(a, f2(s), f2(c)); }
__device__ __forceinline__ f2 abs2(f2 a) { return f2(fabsf(a.v.x), fabsf(a.v.y)); }
__device__ __forceinline__ f2 max2(f2 a, f2 b) { return f2(fmaxf(a.v.x, b.v.x), fmaxf(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 min2(f2 a, f2 b) { return f2(fminf(a.v.x, b.v.x), fminf(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 max2(f2 a, float s) { return f2(fmaxf(a.v.x, s), fmaxf(a.v.y, s)); }
__device__ __forceinline__ f2 clamp2(f2 a, float lo, float hi) { return f2(fminf(fmaxf(a.v.x, lo), hi), fminf(fmaxf(a.v.y, lo), hi)); }
__device__ __forceinline__ f2 sat2(f2 a) { return f2(__saturatef(a.v.x), __saturatef(a.v.y)); }
__device__ __forceinline__ f2 rsqrt2(f2 a) { return f2(rsqrtf(a.v.x), rsqrtf(a.v.y)); }
__device__ __forceinline__ float fw_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ f2 rcp2(f2 a) { return f2(fw_rcp(a.v.x), fw_rcp(a.v.y)); }
__device__ __forceinline__ f2 sin2(f2 a) { return f2(__sinf(a.v.x), __sinf(a.v.y)); }
__device__ __forceinline__ f2 cos2(f2 a) { return f2(__cosf(a.v.x), __cosf(a.v.y)); }
__device__ __forceinline__ b2 lt2(f2 a, f2 b) { return b2{a.v.x < b.v.x, a.v.y < b.v.y}; }
__device__ __forceinline__ b2 gt2(f2 a, f2 b) { return b2{a.v.x > b.v.x, a.v.y > b.v.y}; }
__device__ __forceinline__ b2 gt2(f2 a, float s) { return b2{a.v.x > s, a.v.y > s}; }
__device__ __forceinline__ b2 lt2(f2 a, float s) { return b2{a.v.x < s, a.v.y < s}; }
__device__ __forceinline__ b2 and2(b2 a, b2 b) { return b2{a.x && b.x, a.y && b.y}; }
__device__ __forceinline__ f2 sel2(b2 m, f2 a, f2 b) { return f2(m.x ? a.v.x : b.v.x, m.y ? a.v.y : b.v.y); }

// two environments of one thread: lane x = env 2t, lane y = env 2t + 1
struct EnvState2 {
    f2 px, py, pz;
    f2 qx, qy, qz, qw;
    f2 vx, vy, vz;
    f2 wx, wy, wz;
    f2 act[FWD_NSURF];
    f2 thr;
    f2 new_dist;
    int step_count[2], physics_steps[2], tidx[2];
    uint32_t episode[2];
};

__device__ __forceinline__ EnvState fw_lane(const EnvState2& s, int k) {
    EnvState e;
    e.px = s.px[k]; e.py = s.py[k]; e.pz = s.pz[k];
    e.qx = s.qx[k]; e.qy = s.qy[k]; e.qz = s.qz[k]; e.qw = s.qw[k];
    e.vx = s.vx[k]; e.vy = s.vy[k]; e.vz = s.vz[k];
    e.wx = s.wx[k]; e.wy = s.wy[k]; e.wz = s.wz[k];
#pragma unroll
    for (int i = 0; i < FWD_NSURF; ++i) e.act[i] = s.act[i][k];
    e.thr = s.thr[k]; e.new_dist = s.new_dist[k];
    e.step_count = s.step_count[k]; e.physics_steps = s.physics_steps[k]; e.tidx = s.tidx[k]; e.episode = s.episode[k];
    return e;
}

__device__ __forceinline__ void fw_set(float2& d, int k, float x) { if (k) d.y = x; else d.x = x; }

__device__ __forceinline__ void fw_set_lane(EnvState2& s, int k, const EnvState& e) {
    fw_set(s.px.v, k, e.px); fw_set(s.py.v, k, e.py); fw_set(s.pz.v, k, e.pz);
    fw_set(s.qx.v, k, e.qx); fw_set(s.qy.v, k, e.qy); fw_set(s.qz.v, k, e.qz); fw_set(s.qw.v, k, e.qw);
    fw_set(s.vx.v, k, e.vx); fw_set(s.vy.v, k, e.vy); fw_set(s.vz.v, k, e.vz);
    fw_set(s.wx.v, k, e.wx); fw_set(s.wy.v, k, e.wy); fw_set(s.wz.v, k, e.wz);
#pragma unroll
    for (int i = 0; i < FWD_NSURF; ++i) fw_set(s.act[i].v, k, e.act[i]);
    fw_set(s.thr.v, k, e.thr); fw_set(s.new_dist.v, k, e.new_dist);
    s.step_count[k] = e.step_count; s.physics_steps[k] = e.physics_steps; s.tidx[k] = e.tidx; s.episode[k] = e.episode;
}

// fw_atan2(-ny, x) for both lanes (the caller's y is -vl: taking vl avoids materialising the negation): the degree-17
// polynomial runs packed, range folding stays per lane
__device__ __forceinline__ f2 fw_atan2_neg_p(f2 ny, f2 x) {
    const f2 ax = abs2(x), ay = abs2(ny);
    const f2 mx = max2(ax, ay), mn = min2(ax, ay);
    const f2 a(mx.v.x > 0.0f ? mn.v.x * fw_rcp(mx.v.x) : 0.0f, mx.v.y > 0.0f ? mn.v.y * fw_rcp(mx.v.y) : 0.0f);
    const f2 q = a * a;
    f2 r = fma2(q, 0.00282363896258175373077393f, -0.0159569028764963150024414f);
    r = fma2(r, q, 0.0425049886107444763183594f);
    r = fma2(r, q, -0.0748900920152664184570312f);
    r = fma2(r, q, 0.106347933411598205566406f);
    r = fma2(r, q, -0.142027363181114196777344f);
    r = fma2(r, q, 0.199926957488059997558594f);
    r = fma2(r, q, -0.333331018686294555664062f);
    r = r * q;
    r = fma2(r, a, a);
    r = sel2(gt2(ay, ax), FWD_HALF_PI - r, r);
    r = sel2(lt2(x, 0.0f), FWD_PI - r, r);
    // r >= 0 here; the result carries the sign of y = -ny
    return f2(__int_as_float(__float_as_int(r.v.x) | (~__float_as_int(ny.v.x) & 0x80000000)),
              __int_as_float(__float_as_int(r.v.y) | (~__float_as_int(ny.v.y) & 0x80000000)));
}

// fw_surface<STD = true, LIFT_Y> for both lanes (same statements, same order)
template <bool LIFT_Y>
__device__ __forceinline__ void fw_surface_p(const FwDev& p, const SurfHot& sf, f2 act, f2 vx, f2 vy, f2 vz, f2& fn, f2& fp,
                                             f2& tq) {
    const f2 vl = LIFT_Y ? vy : vz;
    const f2 vf = vx;
    const f2 h2 = fma2(vl, vl, vf * vf);
    const f2 vo = LIFT_Y ? vz : vy;
    const f2 V2 = p.freestream_3d ? fma2(vo, vo, h2) : h2;
    const f2 alpha = fw_atan2_neg_p(vl, vf);
    const f2 inv_h = rsqrt2(max2(h2, 1e-30f));

    const f2 a0 = fma2(act, -sf.k_te, sf.a0_base);
    const f2 asp = fma2(act, -sf.k_shift, sf.asp_base);
    const f2 asn = fma2(act, -sf.k_shift, sf.asn_base);
    const b2 nostall = and2(lt2(asn, alpha), lt2(alpha, asp));
    const b2 pos = gt2(alpha, 0.0f);

    const f2 aa = alpha - a0;
    const f2 cl_lin = aa * sf.cla;
    const f2 ast = sel2(pos, asp, asn);
    const f2 num = sel2(pos, FWD_HALF_PI - alpha, alpha + FWD_HALF_PI);
    const f2 den = sel2(pos, FWD_HALF_PI - asp, asn + FWD_HALF_PI);
    const f2 fac = sat2(num * rcp2(den));
    const f2 ai_stall = ((ast - a0) * sf.cla_ipa) * fac;
    const f2 ai = sel2(nostall, aa * sf.cla_ipa, ai_stall);
    const f2 ae = aa - ai;
    const f2 s = sin2(ae), c = cos2(ae);

    const f2 CT_a = c * sf.cd0;
    const f2 CN_a = fma2(CT_a, s, cl_lin) * rcp2(c);
    const f2 d = act * sf.defl_sel;
    const f2 cd90 = fma2(fma2(d, -4.26e-2f, 2.1e-1f), d, 1.98f);
    const f2 CN_s = (cd90 * s) * (rcp2(fma2(abs2(s), 0.44f, 0.56f)) - sf.stall_k);
    const f2 CN = sel2(nostall, CN_a, CN_s);
    const f2 CT = sel2(nostall, CT_a, CT_a * 0.5f);
    const f2 Cl = fma2(CN, c, -(CT * s));
    const f2 Cd = fma2(CN, s, CT * c);
    const f2 aem = sel2(nostall, ae, abs2(ae));
    const f2 CMc = (-CN) * fma2(aem, sf.cm1c, sf.cm0c);

    const f2 Q = V2 * sf.qarea;
    const f2 Qi = Q * inv_h;
    fn = Qi * fma2(Cl, vf, -(Cd * vl));
    fp = (-Qi) * fma2(Cl, vl, Cd * vf);
    tq = Q * CMc;
}

// fw_substep<STD = true> for both lanes.  cmd[6]: latched actuator commands per lane; wn*: world-frame wind; nz: N(0,1)
// motor-noise draws; contact: per-lane ground-contact flag (accumulated).  p.quat_limiter == 0 is required (the host
// picks the scalar kernels otherwise; the limiter cannot fire at Bullet's default velocity clamp).
__device__ __forceinline__ void fw_substep_p(const FwDev& p, EnvState2& e, const f2 cmd[6], f2 wnx, f2 wny, f2 wnz, f2 nz,
                                             bool contact[2]) {
    const float dt = p.dt;
    // fw_quat_mat, s = 2
    f2 m[9];
    {
        const f2 xs = e.qx * 2.0f, ys = e.qy * 2.0f, zs = e.qz * 2.0f;
        const f2 wx = e.qw * xs, wy = e.qw * ys, wz = e.qw * zs;
        const f2 xx = e.qx * xs, xy = e.qx * ys, xz = e.qx * zs;
        const f2 yy = e.qy * ys, yz = e.qy * zs, zz = e.qz * zs;
        m[0] = 1.0f - (yy + zz); m[1] = xy - wz;          m[2] = xz + wy;
        m[3] = xy + wz;          m[4] = 1.0f - (xx + zz); m[5] = yz - wx;
        m[6] = xz - wy;          m[7] = yz + wx;          m[8] = 1.0f - (xx + yy);
    }
    const f2 rvx = e.vx - wnx, rvy = e.vy - wny, rvz = e.vz - wnz;
    const f2 vbx = fma2(m[6], rvz, fma2(m[3], rvy, m[0] * rvx));
    const f2 vby = fma2(m[7], rvz, fma2(m[4], rvy, m[1] * rvx));
    const f2 vbz = fma2(m[8], rvz, fma2(m[5], rvy, m[2] * rvx));
    const f2 wbx = fma2(m[6], e.wz, fma2(m[3], e.wy, m[0] * e.wx));
    const f2 wby = fma2(m[7], e.wz, fma2(m[4], e.wy, m[1] * e.wx));
    const f2 wbz = fma2(m[8], e.wz, fma2(m[5], e.wy, m[2] * e.wx));

    f2 Fx(0.f), Fy(0.f), Fz(0.f), Tx(0.f), Ty(0.f), Tz(0.f);
#pragma unroll
    for (int s = 0; s < FWD_NSURF; ++s) {
        const SurfHot sf = fw_load_hot(p.hot[s]);
        e.act[s] = fma2(cmd[s] - e.act[s], sf.k_act, e.act[s]);
        const f2 sx = fma2(wby, sf.r[2], fma2(wbz, -sf.r[1], vbx));
        const f2 sy = fma2(wbz, sf.r[0], fma2(wbx, -sf.r[2], vby));
        const f2 sz = fma2(wbx, sf.r[1], fma2(wby, -sf.r[0], vbz));
        f2 fn, fp, tq;
        if (s != 3) {
            fw_surface_p<false>(p, sf, e.act[s], sx, sy, sz, fn, fp, tq);
            Fx += fp; Fz += fn;
            Tx = fma2(fn, sf.ra[0], Tx);
            Ty = Ty + (fma2(fp, sf.rb[1], fn * sf.ra[1]) + tq);
            Tz = fma2(fp, sf.rb[2], Tz);
        } else {
            fw_surface_p<true>(p, sf, e.act[s], sx, sy, sz, fn, fp, tq);
            Fx += fp; Fy += fn;
            Tx = fma2(fn, sf.ra[0], Tx);
            Ty = fma2(fp, sf.rb[1], Ty);
            Tz = Tz + (fma2(fp, sf.rb[2], fn * sf.ra[2]) - tq);
        }
    }
    const float4 q0 = p.rbk[0], q1 = p.rbk[1], q2 = p.rbk[2], q3 = p.rbk[3], q4 = p.rbk[4], q5 = p.rbk[5], q6 = p.rbk[6],
                 q7 = p.rbk[7], q8 = p.rbk[8], q9 = p.rbk[9];
    const float k_motor_k = q0.x, k_noise = q0.y, k_thrust_max = q0.z, k_torque_max = q0.w;
    const float k_rm1 = q1.x, k_rm2 = q1.y, k_gravity = q1.z, k_mass = q1.w;
    const float c0 = q2.x, c1 = q2.y, c2 = q2.z, mv = q2.w;
    {   // motor
        e.thr = fma2(cmd[5] - e.thr, k_motor_k, e.thr);
        e.thr = fma2(nz * e.thr, k_noise, e.thr);
        const f2 t2 = e.thr * e.thr;
        const f2 thrust = t2 * k_thrust_max;
        Fx += thrust;
        Tx = fma2(t2, k_torque_max, Tx);
        Ty = fma2(thrust, k_rm2, Ty);
        Tz = fma2(thrust, -k_rm1, Tz);
    }
    // ground contact on the pose entering the step (rare: per lane, scalar)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (e.pz[k] <= p.col_radius + p.contact_margin) {
            for (int i = 0; i < p.n_col; ++i) {
                const float z = e.pz[k] + m[6][k] * p.col[i][0] + m[7][k] * p.col[i][1] + m[8][k] * p.col[i][2];
                contact[k] = contact[k] || (z <= p.contact_margin);
            }
        }
    }
    // gravity at the composite CoM; g_body = R^T (0,0,-g)
    const f2 gx = m[6] * -k_gravity, gy = m[7] * -k_gravity, gz = m[8] * -k_gravity;
    Fx = fma2(gx, k_mass, Fx); Fy = fma2(gy, k_mass, Fy); Fz = fma2(gz, k_mass, Fz);
    Tx = fma2(fma2(gz, c1, gy * -c2), k_mass, Tx);
    Ty = fma2(fma2(gx, c2, gz * -c0), k_mass, Ty);
    Tz = fma2(fma2(gy, c0, gx * -c1), k_mass, Tz);
    // bias terms: w x (I w) and M w x (w x c)
    const f2 Iwx = fma2(wbz, q3.z, fma2(wby, q3.y, wbx * q3.x));
    const f2 Iwy = fma2(wbz, q4.y, fma2(wby, q4.x, wbx * q3.w));
    const f2 Iwz = fma2(wbz, q5.x, fma2(wby, q4.w, wbx * q4.z));
    const f2 b0 = Tx - fma2(wby, Iwz, -(wbz * Iwy));
    const f2 b1 = Ty - fma2(wbz, Iwx, -(wbx * Iwz));
    const f2 b2_ = Tz - fma2(wbx, Iwy, -(wby * Iwx));
    const f2 cx = fma2(wby, c2, wbz * -c1);
    const f2 cy = fma2(wbz, c0, wbx * -c2);
    const f2 cz = fma2(wbx, c1, wby * -c0);
    const f2 b3 = fma2(fma2(wby, cz, -(wbz * cy)), -k_mass, Fx);
    const f2 b4 = fma2(fma2(wbz, cx, -(wbx * cz)), -k_mass, Fy);
    const f2 b5 = fma2(fma2(wbx, cy, -(wby * cx)), -k_mass, Fz);
    // lateral block {wx, wz, vy} <- (b0, b2, b4); longitudinal block {wy, vx, vz} <- (b1, b3, b5)
    const f2 a0 = fma2(b4, q5.w, fma2(b2_, q5.z, b0 * q5.y));
    const f2 a2 = fma2(b4, q6.z, fma2(b2_, q6.y, b0 * q6.x));
    const f2 a4 = fma2(b4, q7.y, fma2(b2_, q7.x, b0 * q6.w));
    const f2 a1 = fma2(b5, q8.x, fma2(b3, q7.w, b1 * q7.z));
    const f2 a3 = fma2(b5, q8.w, fma2(b3, q8.z, b1 * q8.y));
    const f2 a5 = fma2(b5, q9.z, fma2(b3, q9.y, b1 * q9.x));
    // to the world frame, semi-implicit Euler, Bullet's per-coordinate velocity clamp
    const f2 awx = fma2(m[2], a2, fma2(m[1], a1, m[0] * a0));
    const f2 awy = fma2(m[5], a2, fma2(m[4], a1, m[3] * a0));
    const f2 awz = fma2(m[8], a2, fma2(m[7], a1, m[6] * a0));
    const f2 Awx = fma2(m[2], a5, fma2(m[1], a4, m[0] * a3));
    const f2 Awy = fma2(m[5], a5, fma2(m[4], a4, m[3] * a3));
    const f2 Awz = fma2(m[8], a5, fma2(m[7], a4, m[6] * a3));
    e.wx = clamp2(fma2(awx, dt, e.wx), -mv, mv);
    e.wy = clamp2(fma2(awy, dt, e.wy), -mv, mv);
    e.wz = clamp2(fma2(awz, dt, e.wz), -mv, mv);
    e.vx = clamp2(fma2(Awx, dt, e.vx), -mv, mv);
    e.vy = clamp2(fma2(Awy, dt, e.vy), -mv, mv);
    e.vz = clamp2(fma2(Awz, dt, e.vz), -mv, mv);
    e.px = fma2(e.vx, dt, e.px); e.py = fma2(e.vy, dt, e.py); e.pz = fma2(e.vz, dt, e.pz);
    {   // exponential-map quaternion update, half angle by Taylor series (see fw_substep)
        const f2 ang2 = fma2(e.wz, e.wz, fma2(e.wy, e.wy, e.wx * e.wx));
        const f2 h2 = ang2 * (0.25f * dt * dt);
        const f2 sinc = fma2(h2, fma2(h2, fma2(h2, -1.0f / 5040.0f, 1.0f / 120.0f), -1.0f / 6.0f), 1.0f);
        const f2 cw = fma2(h2, fma2(h2, fma2(h2, fma2(h2, 1.0f / 40320.0f, -1.0f / 720.0f), 1.0f / 24.0f), -0.5f), 1.0f);
        const f2 k = sinc * (0.5f * dt);
        const f2 ax = e.wx * k, ay = e.wy * k, az = e.wz * k;
        const f2 nx = fma2(cw, e.qx, e.qw * ax) + fma2(ay, e.qz, -(az * e.qy));
        const f2 ny = fma2(cw, e.qy, e.qw * ay) + fma2(az, e.qx, -(ax * e.qz));
        const f2 nzq = fma2(cw, e.qz, e.qw * az) + fma2(ax, e.qy, -(ay * e.qx));
        const f2 nw = fma2(cw, e.qw, -fma2(az, e.qz, fma2(ay, e.qy, ax * e.qx)));
        const f2 inv = rsqrt2(fma2(nw, nw, fma2(nzq, nzq, fma2(ny, ny, nx * nx))));
        e.qx = nx * inv; e.qy = ny * inv; e.qz = nzq * inv; e.qw = nw * inv;
    }
    e.physics_steps[0] += 1; e.physics_steps[1] += 1;
}

__device__ __forceinline__ void fw_load2(const FwPlanes& pl, int i0, int n_end, EnvState2& s) {
    // lane y of the last thread of an odd batch has no env: it mirrors lane x and is never stored
    const int i1 = (i0 + 1 < n_end) ? i0 + 1 : i0;
    EnvState a, b;
    fw_load(pl, i0, a);
    fw_load(pl, i1, b);
    fw_set_lane(s, 0, a);
    fw_set_lane(s, 1, b);
}
