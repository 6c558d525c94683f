// fw_objlock.cuh -- device side of the Waypoint + ObjLock task (BASELINE config 4, "ground-truth target
// observation, no YOLO"): duck/obstacle spawn, analytic pin-hole chase camera, the nine vision features, the
// duck-phase state machine and the obstacle-avoidance / lock / approach / strike rewards.
//
// Restates /root/reference/envs/fixedwing_waypoint_objlock_env.py:
//   reset/_spawn_duck/_spawn_obstacles  :170-195,382-519     compute_state            :197-276
//   compute_term_trunc_reward           :278-343              _apply_obstacle_avoidance :347-380
//   _compute_vision_features & helpers  :575-693              camera interval           :563-573
// The reference rasterises a 128x128 segmentation/depth image with PyBullet (TinyRenderer) every 12 substeps and
// reduces it to nine numbers; here the same nine numbers come from an analytic camera: duck = sphere,
// obstacles = finite vertical cylinders, ground = plane z = 0, one ray per pixel column of the middle row.
// That is a documented stand-in for the mesh renderer, restated identically in the fp64 oracle.
#pragma once

#include "fw_device.cuh"

#define OL_MAX_OBST 32

// per-env ObjLock state carried in registers during a step
struct OlState {
    float dkx, dky, dkz;                             // duck position (world)
    float last_cx, last_cy, last_area, last_depth;   // _last_* of the reference
    float f_cx, f_cy, f_area, f_depth;               // latest captured frame: duck silhouette
    float f_dl, f_dc, f_dr;                          // latest captured frame: obstacle bands (metres)
    float prev_est;                                  // _prev_est_dist_m
    int duck_phase, has_prev, post_wp, cam_valid, f_visible;
    int seen, lock, since;                           // _seen_consecutive, _lock_steps, _steps_since_seen
    int n_obst;
    float vis_dl, vis_dc, vis_dr;                    // duck_vision[6:9] of the current compute_state
    float vis_flag;                                  // duck_vision[0]
};

__device__ __forceinline__ void ol_load(const FwPlanes& pl, int i, OlState& o) {
    float4 d = pl.dk[i], a = pl.v0[i], b = pl.v1[i], c = pl.v2[i];
    int4 g = pl.v3[i];
    o.dkx = d.x; o.dky = d.y; o.dkz = d.z;
    o.last_cx = a.x; o.last_cy = a.y; o.last_area = a.z; o.last_depth = a.w;
    o.f_cx = b.x; o.f_cy = b.y; o.f_area = b.z; o.f_depth = b.w;
    o.f_dl = c.x; o.f_dc = c.y; o.f_dr = c.z; o.prev_est = c.w;
    o.duck_phase = g.x & 1; o.has_prev = (g.x >> 1) & 1; o.post_wp = (g.x >> 2) & 1; o.cam_valid = (g.x >> 3) & 1;
    o.f_visible = (g.x >> 4) & 1; o.n_obst = (g.x >> 8) & 0xff;
    o.seen = g.y; o.lock = g.z; o.since = g.w;
    o.vis_dl = o.vis_dc = o.vis_dr = 0.0f; o.vis_flag = 0.0f;
}

__device__ __forceinline__ void ol_store(const FwPlanes& pl, int i, const OlState& o) {
    pl.dk[i] = make_float4(o.dkx, o.dky, o.dkz, 0.0f);
    pl.v0[i] = make_float4(o.last_cx, o.last_cy, o.last_area, o.last_depth);
    pl.v1[i] = make_float4(o.f_cx, o.f_cy, o.f_area, o.f_depth);
    pl.v2[i] = make_float4(o.f_dl, o.f_dc, o.f_dr, o.prev_est);
    int bits = (o.duck_phase & 1) | ((o.has_prev & 1) << 1) | ((o.post_wp & 1) << 2) | ((o.cam_valid & 1) << 3) |
               ((o.f_visible & 1) << 4) | ((o.n_obst & 0xff) << 8);
    pl.v3[i] = make_int4(bits, o.seen, o.lock, o.since);
}

// obstacle table of this thread in shared memory: so[(k*3 + c) * stride + tid]; c = x, y, height
#define OL_S(so, k, c, tid, stride) (so)[((k) * 3 + (c)) * (stride) + (tid)]

__device__ __forceinline__ void ol_stage_obstacles(const FwDev& p, const FwPlanes& pl, int i, int n_obst, float* so, int tid,
                                                   int stride) {
    for (int k = 0; k < n_obst; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) OL_S(so, k, c, tid, stride) = pl.obst[(size_t)(k * 3 + c) * p.n + i];
}

// nearest positive hit of ray o + t d with the finite vertical cylinder (cx, cy, height h, radius r); inf if none
__device__ __forceinline__ float ol_ray_cylinder(float ox, float oy, float oz, float dx, float dy, float dz, float cx,
                                                 float cy, float h, float r) {
    const float INF = __int_as_float(0x7f800000);
    float px = ox - cx, py = oy - cy;
    float a = dx * dx + dy * dy;
    float best = INF;
    if (a > 1e-12f) {
        float b = px * dx + py * dy;
        float cc = px * px + py * py - r * r;
        float disc = b * b - a * cc;
        if (disc >= 0.0f) {
            float sq = sqrtf(disc);
            float inv = 1.0f / a;
            float t0 = (-b - sq) * inv, t1 = (-b + sq) * inv;
            if (t0 > 0.0f) { float z = oz + t0 * dz; if (z >= 0.0f && z <= h && t0 < best) best = t0; }
            if (t1 > 0.0f) { float z = oz + t1 * dz; if (z >= 0.0f && z <= h && t1 < best) best = t1; }
        }
    }
    if (fabsf(dz) > 1e-12f) {     // top cap
        float t = (h - oz) / dz;
        if (t > 0.0f) {
            float x = px + t * dx, y = py + t * dy;
            if (x * x + y * y <= r * r && t < best) best = t;
        }
    }
    return best;
}

__device__ __forceinline__ float ol_ray_sphere(float ox, float oy, float oz, float dx, float dy, float dz, float sx,
                                               float sy, float sz, float r) {
    const float INF = __int_as_float(0x7f800000);
    float px = ox - sx, py = oy - sy, pz = oz - sz;
    float a = dx * dx + dy * dy + dz * dz, b = px * dx + py * dy + pz * dz, cc = px * px + py * py + pz * pz - r * r;
    float disc = b * b - a * cc;
    if (disc < 0.0f) return INF;
    float t = (-b - sqrtf(disc)) / a;
    return t > 0.0f ? t : INF;
}

// The reference averages OpenGL depth-BUFFER values d = far (z - near) / ((far - near) z) over a band and maps the
// mean back to metres with z = far near / (far - (far - near) d).  Since d is affine in 1/z, that round trip is
// exactly the HARMONIC mean of the clamped depths, 1 / mean(1 / clamp(z, near, far)) (a miss counts as z = far).
// The harmonic form is what the kernel evaluates: the literal form loses all fp32 precision near d = 1.
__device__ __forceinline__ float ol_inv_depth(const FwDev& p, float z) {
    return 1.0f / fminf(fmaxf(z, p.cam_near), p.cam_far);
}
__device__ __forceinline__ float ol_band_metres(const FwDev& p, float sum_inv, int cnt) {
    if (cnt == 0) return 0.0f;
    float mean_inv = sum_inv / (float)cnt;
    if (mean_inv * p.cam_near >= 1.0f) return 0.0f;       // every pixel on the near plane: mean buffer value 0 -> 0.0
    return 1.0f / mean_inv;
}

// Conservative pixel-column interval [lo, hi] of the middle row whose rays can pass within `rad` (horizontally) of
// the vertical axis through (px, py) relative to the camera.  Horizontal ray directions form the family A + xn B
// (xn = normalised column coordinate in [-1, 1]); |cross(A + xn B, P)| <= rad |A + xn B| is a quadratic in xn.
// Columns outside the interval cannot hit the cylinder / sphere, so skipping them leaves the frame unchanged.
__device__ __forceinline__ void ol_col_interval(float Ax, float Ay, float Bx, float By, float AA, float AB, float BB,
                                                float px, float py, float rad, int w, int& lo, int& hi) {
    const float rr = rad * rad * 1.01f + 1e-4f;
    const float cA = Ax * py - Ay * px, cB = Bx * py - By * px;
    const float qa = cB * cB - rr * BB, qb = cA * cB - rr * AB, qc = cA * cA - rr * AA;
    lo = 0; hi = w - 1;
    if (qa > 1e-7f) {
        const float disc = qb * qb - qa * qc;
        if (disc < 0.0f) { lo = 1; hi = 0; return; }
        const float sq = sqrtf(disc), inv = 1.0f / qa;
        const float x0 = (-qb - sq) * inv, x1 = (-qb + sq) * inv;
        const float c0 = (x0 + 1.0f) * 0.5f * (float)w - 0.5f, c1 = (x1 + 1.0f) * 0.5f * (float)w - 0.5f;
        // column j can be hit only if c0 <= j <= c1; floor / ceil leave up to one column of slack on each side on top of the
        // 1 % inflation of the radius, far more than the ~1e-4 columns of fp32 error in c0 and c1
        lo = max(0, (int)floorf(fmaxf(c0, -4.0f)));
        hi = min(w - 1, (int)ceilf(fminf(c1, (float)w + 4.0f)));
    }
    // The quadratic bounds LINES through the camera, so an axis behind the camera passes it as well.  A ray t > 0 along
    // d = A + xn B reaches the disk only if P.d + rad |d| > 0; P.d is linear and |d|^2 convex in xn, so both are largest
    // at an end of the interval: when the test fails at both ends (and the camera is outside the disk) no column of the
    // interval can hit the cylinder or its cap, and skipping it leaves the frame unchanged.
    if (lo <= hi && px * px + py * py > rr) {
        const float xl = (2.0f * (float)lo + 1.0f) / (float)w - 1.0f, xh = (2.0f * (float)hi + 1.0f) / (float)w - 1.0f;
        const float PA = px * Ax + py * Ay, PB = px * Bx + py * By;
        const float pd = fmaxf(fmaf(xl, PB, PA), fmaf(xh, PB, PA));
        const float dd = fmaxf(fmaf(xl, fmaf(xl, BB, 2.0f * AB), AA), fmaf(xh, fmaf(xh, BB, 2.0f * AB), AA));
        if (pd <= 0.0f && pd * pd >= rr * dd) { lo = 1; hi = 0; }
    }
}

// Camera.capture_image stand-in + the image reductions of _compute_vision_features, WARP-COOPERATIVE.
// A thread-per-env ray caster does cam_res x n_obst cylinder tests per frame and, because every aircraft sees a
// different obstacle layout, SIMT execution cannot skip any of them.  Here the 32 lanes of the warp serve one
// capturing env at a time: lanes = obstacles (each rasterises only the few pixel columns its cylinder can cover,
// into a shared depth row with atomicMin), then lanes = pixel columns (duck mask, inverse depth, band sums by
// shuffle).  Must be called by all 32 lanes; `need` marks the lanes whose env captures a frame now.
__device__ __forceinline__ void ol_capture_warp(const FwDev& p, bool need, const EnvState& e, OlState& o, const float* so,
                                                int tid, int stride, float* depth_row) {
    const unsigned FULL = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);
    const int lane = tid & 31;
    unsigned pending = __ballot_sync(FULL, need);
    if (pending == 0u) return;
    // every capturing lane prepares its own camera (position, ray family A + xn r, duck sphere)
    float cx = 0.f, cy = 0.f, cz = 0.f, Ax = 0.f, Ay = 0.f, Az = 0.f, rx = 0.f, ry = 0.f, rz = 0.f;
    float sx = 0.f, sy = 0.f, sz = 0.f, zc = 1.f, xc = 0.f, yc = 0.f, ddx = 0.f, ddy = 0.f, ddz = 0.f, dist = 0.f;
    int cand = 0;
    const float Rd = p.duck_radius;
    const int w = p.cam_res, x1 = w / 3, x2 = 2 * w / 3, ymid = w / 2;
    const float vrow = 2.0f * ((float)ymid + 0.5f) / (float)w - 1.0f;
    // normalised coordinate of pixel column c: xn = 2 (c + 0.5) / w - 1 = c * xs + x0
    const float xs = 2.0f / (float)w, x0 = 0.5f * xs - 1.0f;
    // the depth row is all-INF whenever a frame starts: cleared here once per call, and every frame's band pass puts INF
    // back into the entries it has read
    for (int col = lane; col < w; col += 32) depth_row[col] = INF;
    __syncwarp();
    if (need) {
        Mat3 R = fw_quat_mat(e.qx, e.qy, e.qz, e.qw);
        const float* m = R.m;
        // camera position = pos + R * offset ; forward = -R*offset/|offset| ; up hint = body z
        float ofx = m[0] * p.cam_offset[0] + m[1] * p.cam_offset[1] + m[2] * p.cam_offset[2];
        float ofy = m[3] * p.cam_offset[0] + m[4] * p.cam_offset[1] + m[5] * p.cam_offset[2];
        float ofz = m[6] * p.cam_offset[0] + m[7] * p.cam_offset[1] + m[8] * p.cam_offset[2];
        cx = e.px + ofx; cy = e.py + ofy; cz = e.pz + ofz;
        float fx, fy, fz, ux, uy, uz;
        if (p.cam_mode == 0) {       // tracking camera: looks at the aircraft, up hint = body z
            float il = rsqrtf(ofx * ofx + ofy * ofy + ofz * ofz);
            fx = -ofx * il; fy = -ofy * il; fz = -ofz * il;
            ux = m[2]; uy = m[5]; uz = m[8];
        } else {                     // fixed camera (is_tracking_camera = False): tilted body axes, folded on the host
            fx = m[0] * p.cam_fb[0] + m[1] * p.cam_fb[1] + m[2] * p.cam_fb[2];
            fy = m[3] * p.cam_fb[0] + m[4] * p.cam_fb[1] + m[5] * p.cam_fb[2];
            fz = m[6] * p.cam_fb[0] + m[7] * p.cam_fb[1] + m[8] * p.cam_fb[2];
            ux = m[0] * p.cam_ub[0] + m[1] * p.cam_ub[1] + m[2] * p.cam_ub[2];
            uy = m[3] * p.cam_ub[0] + m[4] * p.cam_ub[1] + m[5] * p.cam_ub[2];
            uz = m[6] * p.cam_ub[0] + m[7] * p.cam_ub[1] + m[8] * p.cam_ub[2];
        }
        rx = fy * uz - fz * uy; ry = fz * ux - fx * uz; rz = fx * uy - fy * ux;
        float rl = rsqrtf(rx * rx + ry * ry + rz * rz);
        rx *= rl; ry *= rl; rz *= rl;
        ux = ry * fz - rz * fy; uy = rz * fx - rx * fz; uz = rx * fy - ry * fx;
        Ax = fx - vrow * ux; Ay = fy - vrow * uy; Az = fz - vrow * uz;
        // duck silhouette: sphere of radius duck_radius resting on the ground
        sx = o.dkx; sy = o.dky; sz = o.dkz + Rd;
        float dvx = sx - cx, dvy = sy - cy, dvz = sz - cz;
        zc = dvx * fx + dvy * fy + dvz * fz; xc = dvx * rx + dvy * ry + dvz * rz; yc = dvx * ux + dvy * uy + dvz * uz;
        cand = (zc - Rd > p.cam_near && fabsf(xc) <= zc + Rd && fabsf(yc) <= zc + Rd && zc - Rd < p.cam_far) ? 1 : 0;
        dist = sqrtf(dvx * dvx + dvy * dvy + dvz * dvz);
        float id = 1.0f / dist;
        ddx = dvx * id; ddy = dvy * id; ddz = dvz * id;
    }
    while (pending) {
        const int s = __ffs(pending) - 1;
        pending &= pending - 1u;
        const int src = (tid & ~31) | s;                      // shared-memory column of the capturing env
        const float bcx = __shfl_sync(FULL, cx, s), bcy = __shfl_sync(FULL, cy, s), bcz = __shfl_sync(FULL, cz, s);
        const float bAx = __shfl_sync(FULL, Ax, s), bAy = __shfl_sync(FULL, Ay, s), bAz = __shfl_sync(FULL, Az, s);
        const float brx = __shfl_sync(FULL, rx, s), bry = __shfl_sync(FULL, ry, s), brz = __shfl_sync(FULL, rz, s);
        const float bsx = __shfl_sync(FULL, sx, s), bsy = __shfl_sync(FULL, sy, s), bsz = __shfl_sync(FULL, sz, s);
        const float bdx = __shfl_sync(FULL, ddx, s), bdy = __shfl_sync(FULL, ddy, s), bdz = __shfl_sync(FULL, ddz, s);
        const float bdist = __shfl_sync(FULL, dist, s);
        const int bcand = __shfl_sync(FULL, cand, s), bn = __shfl_sync(FULL, o.n_obst, s);
        // phase A: lane k owns cylinder k for the set-up (column interval, duck-occlusion ray); the (cylinder,
        // column) pairs to rasterise are then dealt out evenly over the 32 lanes through a warp prefix sum, so a
        // wide nearby cylinder does not serialise the warp behind one lane
        bool occl = false;
        const float AA = bAx * bAx + bAy * bAy, AB = bAx * brx + bAy * bry, BB = brx * brx + bry * bry;
        float ox = 0.f, oy = 0.f, oh = 0.f;
        int lo = 1, hi = 0;
        if (lane < bn) {
            ox = OL_S(so, lane, 0, src, stride); oy = OL_S(so, lane, 1, src, stride); oh = OL_S(so, lane, 2, src, stride);
            ol_col_interval(bAx, bAy, brx, bry, AA, AB, BB, ox - bcx, oy - bcy, p.obst_radius, w, lo, hi);
            if (bcand) {
                float t = ol_ray_cylinder(bcx, bcy, bcz, bdx, bdy, bdz, ox, oy, oh, p.obst_radius);
                occl = t < bdist - Rd;
            }
        }
        const int wdt = max(hi - lo + 1, 0);
        int scan = wdt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(FULL, scan, off);
            if (lane >= off) scan += t;
        }
        const int total = __shfl_sync(FULL, scan, 31);
        const int start = scan - wdt;
        for (int base = 0; base < total; base += 32) {
            const int pidx = base + lane;
            int k = 0;                                   // owner = number of lanes whose inclusive scan <= pidx
#pragma unroll
            for (int stp = 16; stp > 0; stp >>= 1) {
                const int candk = k + stp;
                const int sv = __shfl_sync(FULL, scan, (candk - 1) & 31);
                if (candk <= 32 && sv <= pidx) k = candk;
            }
            const int kk = k & 31;
            const int klo = __shfl_sync(FULL, lo, kk), kst = __shfl_sync(FULL, start, kk);
            const float kx = __shfl_sync(FULL, ox, kk), ky = __shfl_sync(FULL, oy, kk), kh = __shfl_sync(FULL, oh, kk);
            if (pidx < total) {
                const int col = klo + (pidx - kst);
                const float xn = fmaf((float)col, xs, x0);
                float t = ol_ray_cylinder(bcx, bcy, bcz, bAx + xn * brx, bAy + xn * bry, bAz + xn * brz, kx, ky, kh, p.obst_radius);
                if (t < INF) atomicMin(reinterpret_cast<int*>(&depth_row[col]), __float_as_int(t));
            }
        }
        const unsigned hidden = __ballot_sync(FULL, occl);
        __syncwarp();
        // phase B: the three harmonic band means of the middle row; duck pixels drop out.  The sphere shows in the row only
        // if its centre lies within a radius of the plane of the row's rays (normal (0, 1, vrow) in camera coordinates) and
        // not wholly behind the camera; only then is its column interval worked out.
        int dlo = 1, dhi = 0;
        bool duck_row;
        {
            const float byc = __shfl_sync(FULL, yc, s), bzc = __shfl_sync(FULL, zc, s);
            duck_row = fabsf(fmaf(vrow, bzc, byc)) <= (Rd * 1.001f + 1e-4f) * sqrtf(fmaf(vrow, vrow, 1.0f)) && bzc + Rd > 0.0f;
        }
        if (duck_row) {
            ol_col_interval(bAx, bAy, brx, bry, AA, AB, BB, bsx - bcx, bsy - bcy, Rd, w, dlo, dhi);
            duck_row = dlo <= dhi;
        }
        // The inverse depth of the GROUND along the middle row is linear in the column index, iv_j = clamp(alpha + beta j,
        // 1/far, 1/near) -- a ray that misses the ground has alpha + beta j <= 0 and clamps to 1/far like a miss -- so the
        // ground-only sum of a band is two clamped runs plus one arithmetic series: lanes 0..2 evaluate the three bands in
        // closed form.  Only columns a cylinder was rasterised into, or that may show the duck, are then visited to CORRECT
        // those sums (32-column groups without such a column are skipped; frames without any -- most frames of the
        // obstacle-free duck-only configuration -- need no column pass at all).
        const float inv_far = 1.0f / p.cam_far, inv_near = 1.0f / p.cam_near;
        float alpha = 0.0f, beta = 0.0f;                         // camera at or below the ground: every ray misses
        if (bcz > 0.0f) {
            const float icz = 1.0f / bcz;
            alpha = -fmaf(x0, brz, bAz) * icz; beta = -(xs * brz) * icz;
        }
        float bsum;
        int bcnt;
        {
            const int j0 = lane == 0 ? 0 : (lane == 1 ? x1 : x2), j1 = lane == 0 ? x1 : (lane == 1 ? x2 : w);
            const float n = (float)(j1 - j0);
            if (fabsf(beta) * (float)w < 1e-12f) {
                bsum = n * fminf(fmaxf(alpha, inv_far), inv_near);
            } else {
                const float ib = 1.0f / beta;
                const float ja = (inv_far - alpha) * ib, jb = (inv_near - alpha) * ib;      // where the two clamps engage
                const float jA = fminf(fmaxf(fminf(ja, jb), -1.0f), (float)w + 1.0f);
                const float jB = fminf(fmaxf(fmaxf(ja, jb), -1.0f), (float)w + 1.0f);
                const int cA = min(max((int)ceilf(jA), j0), j1);                            // below: [j0, cA)
                const int cB = min(max((int)floorf(jB) + 1, cA), j1);                       // linear: [cA, cB), above: [cB, j1)
                const float below = beta > 0.0f ? inv_far : inv_near, above = beta > 0.0f ? inv_near : inv_far;
                const float nm = (float)(cB - cA);
                bsum = below * (float)(cA - j0) + above * (float)(j1 - cB) + fmaf(beta, 0.5f * (float)(cA + cB - 1) * nm, alpha * nm);
            }
            bcnt = j1 - j0;
        }
        if (total != 0 || duck_row) {
            float cor0 = 0.f, cor1 = 0.f, cor2 = 0.f;
            int dec3 = 0;                                    // duck pixels leaving the three bands, 10 bits each
            for (int base = 0; base < w; base += 32) {
                const int col = base + lane;
                const bool valid = col < w;
                const float tc = valid ? depth_row[col] : INF;
                const bool touched = tc < INF;
                const bool ind = duck_row && valid && col >= dlo && col <= dhi;
                if (!__any_sync(FULL, touched || ind)) continue;
                if (touched) depth_row[col] = INF;           // ready for the next frame
                if (touched || ind) {
                    const float gl = fmaf(beta, (float)col, alpha);              // ground, unclamped (<= 0: no hit)
                    const float ground = fminf(fmaxf(gl, inv_far), inv_near);
                    const float itc = touched ? 1.0f / tc : 0.0f;
                    bool duck_px = false;
                    if (ind) {
                        const float xn = fmaf((float)col, xs, x0);
                        const float td = ol_ray_sphere(bcx, bcy, bcz, bAx + xn * brx, bAy + xn * bry, bAz + xn * brz, bsx, bsy, bsz, Rd);
                        duck_px = td < INF && 1.0f / td > fmaxf(itc, gl);        // nearer than the cylinder and the ground
                    }
                    // a duck pixel leaves its band; otherwise the nearer of cylinder and ground replaces the ground
                    const float corr = duck_px ? -ground : fmaxf(fminf(fmaxf(itc, inv_far), inv_near), ground) - ground;
                    const int dec = duck_px ? 1 : 0;
                    if (col < x1) { cor0 += corr; dec3 += dec; }
                    else if (col < x2) { cor1 += corr; dec3 += dec << 10; }
                    else { cor2 += corr; dec3 += dec << 20; }
                }
            }
            // NOTE: fp32 addition is not associative; the oracle sums columns left to right in fp64, so the order
            // here only moves the last bits of a quantity that is compared with a 2e-4 relative tolerance
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                cor0 += __shfl_xor_sync(FULL, cor0, off); cor1 += __shfl_xor_sync(FULL, cor1, off);
                cor2 += __shfl_xor_sync(FULL, cor2, off); dec3 += __shfl_xor_sync(FULL, dec3, off);
            }
            bsum += lane == 0 ? cor0 : (lane == 1 ? cor1 : cor2);
            bcnt -= (dec3 >> (10 * min(lane, 2))) & 1023;
        }
        // mean inverse depth -> metres for the three bands at once (lane b evaluates band b)
        const float bmet = ol_band_metres(p, bsum, bcnt);
        const float met_l = __shfl_sync(FULL, bmet, 0), met_c = __shfl_sync(FULL, bmet, 1), met_r = __shfl_sync(FULL, bmet, 2);
        if (lane == s) {
            const int vis = (cand && hidden == 0u) ? 1 : 0;
            o.f_visible = vis;
            if (vis) {
                float ccx = 0.5f + 0.5f * xc / zc, ccy = 0.5f - 0.5f * yc / zc;
                o.f_cx = fminf(fmaxf(ccx, 0.0f), 1.0f);
                o.f_cy = fminf(fmaxf(ccy, 0.0f), 1.0f);
                float a = FWD_PI * (Rd / zc) * (Rd / zc) * 0.25f;
                o.f_area = fminf(a, 1.0f);
                o.f_depth = zc - Rd;
            }
            o.f_dl = met_l; o.f_dc = met_c; o.f_dr = met_r;
            o.cam_valid = 1;
        }
        __syncwarp();
    }
}

// _compute_vision_features + the phase switching of compute_state (fixedwing_waypoint_objlock_env.py:248-274)
__device__ __forceinline__ void ol_vision_and_phase(const FwDev& p, OlState& o, bool all_reached) {
    float visible = 0.0f, dl = 0.0f, dc = 0.0f, dr = 0.0f;
    if (o.cam_valid) {
        dl = o.f_dl; dc = o.f_dc; dr = o.f_dr;
        if (!o.f_visible) {
            o.since = min(o.since + 1, 60);
        } else {
            o.last_cx = o.f_cx; o.last_cy = o.f_cy; o.last_area = o.f_area; o.last_depth = o.f_depth;
            o.since = 0;
            visible = 1.0f;
        }
    }
    o.vis_flag = visible; o.vis_dl = dl; o.vis_dc = dc; o.vis_dr = dr;
    if (all_reached) {
        o.post_wp = 1;
        if (!o.duck_phase) {
            const bool vis = visible > 0.5f && o.last_area >= p.switch_min_area;
            o.seen = vis ? o.seen + 1 : 0;
            if (o.seen >= p.switch_min_seen) o.duck_phase = 1;
        }
    } else {
        o.post_wp = 0;
        o.duck_phase = 0;
    }
}

__device__ __forceinline__ float ol_obstacle_penalty(const FwDev& p, const OlState& o, bool duck_phase) {
    const float INF = __int_as_float(0x7f800000);
    float dmin = INF;
    if (o.vis_dl > 0.0f && o.vis_dl < dmin) dmin = o.vis_dl;
    if (o.vis_dc > 0.0f && o.vis_dc < dmin) dmin = o.vis_dc;
    if (o.vis_dr > 0.0f && o.vis_dr < dmin) dmin = o.vis_dr;
    if (!(dmin < INF)) return 0.0f;
    if (p.obst_safe <= 0.0f || dmin >= p.obst_safe) return 0.0f;
    float scale = p.obst_scale * (duck_phase ? 0.5f : 1.0f);
    return fminf(scale * (p.obst_safe - dmin) / p.obst_safe, p.obst_max_pen);
}

// Bodies the aircraft could touch during the coming agent step: bit k = cylinder k, `duck_near` = the duck.  The velocity
// clamp bounds the travel of one agent step (sqrt(3) max_vel dt per substep), so a body farther than that plus its
// own reach cannot be in contact on any of the step's substeps; the per-substep test then only visits the set bits
// (usually none) instead of scanning the whole obstacle table eight times per step.
__device__ __forceinline__ uint32_t ol_contact_candidates(const FwDev& p, const EnvState& e, const OlState& o, const float* so,
                                                          int tid, int stride, bool& duck_near) {
    const float reach = p.col_radius + p.contact_margin;
    const float travel = 1.7320508f * p.max_vel * p.dt * (float)(p.substeps_per_inner * p.inner_per_step) * 1.001f + 1e-3f;
    uint32_t mask = 0u;
    const float rr = p.obst_radius + reach + travel;
    for (int k = 0; k < o.n_obst; ++k) {
        float ddx = e.px - OL_S(so, k, 0, tid, stride), ddy = e.py - OL_S(so, k, 1, tid, stride);
        if (ddx * ddx + ddy * ddy <= rr * rr && e.pz <= OL_S(so, k, 2, tid, stride) + reach + travel) mask |= 1u << k;
    }
    {
        float ddx = e.px - o.dkx, ddy = e.py - o.dky, ddz = e.pz - (o.dkz + p.duck_radius);
        float rd = p.duck_radius + reach + travel;
        duck_near = ddx * ddx + ddy * ddy + ddz * ddz <= rd * rd;
    }
    return mask;
}

// contact with obstacles / duck on the pose entering the step (collision probe points vs cylinders and sphere);
// `cand` = ol_contact_candidates of the agent step this substep belongs to
__device__ __forceinline__ bool ol_contact(const FwDev& p, const EnvState& e, const OlState& o, const float* so, int tid,
                                           int stride, uint32_t cand, bool duck_near) {
    if (cand == 0u && !duck_near) return false;
    bool hit = false;
    const float reach = p.col_radius + p.contact_margin;
    // cheap reject first: most aircraft are nowhere near an obstacle, so the rotation matrix is rarely needed
    bool near_any = false;
    for (uint32_t mk = cand; mk; mk &= mk - 1u) {
        const int k = __ffs(mk) - 1;
        float ddx = e.px - OL_S(so, k, 0, tid, stride), ddy = e.py - OL_S(so, k, 1, tid, stride);
        float rr = p.obst_radius + reach;
        near_any = near_any || (ddx * ddx + ddy * ddy <= rr * rr && e.pz <= OL_S(so, k, 2, tid, stride) + reach);
    }
    if (duck_near) {
        float ddx = e.px - o.dkx, ddy = e.py - o.dky, ddz = e.pz - (o.dkz + p.duck_radius);
        float rr = p.duck_radius + reach;
        near_any = near_any || (ddx * ddx + ddy * ddy + ddz * ddz <= rr * rr);
    }
    if (!near_any) return false;
    Mat3 R = fw_quat_mat(e.qx, e.qy, e.qz, e.qw);
    const float* m = R.m;
    for (uint32_t mk = cand; mk; mk &= mk - 1u) {
        const int k = __ffs(mk) - 1;
        float ox = OL_S(so, k, 0, tid, stride), oy = OL_S(so, k, 1, tid, stride), oh = OL_S(so, k, 2, tid, stride);
        float ddx = e.px - ox, ddy = e.py - oy;
        float rr = p.obst_radius + reach;
        if (ddx * ddx + ddy * ddy > rr * rr || e.pz > oh + reach) continue;     // conservative prune
        const float r2 = (p.obst_radius + p.contact_margin) * (p.obst_radius + p.contact_margin);
        for (int c = 0; c < p.n_col; ++c) {
            float wx = e.px + m[0] * p.col[c][0] + m[1] * p.col[c][1] + m[2] * p.col[c][2];
            float wy = e.py + m[3] * p.col[c][0] + m[4] * p.col[c][1] + m[5] * p.col[c][2];
            float wz = e.pz + m[6] * p.col[c][0] + m[7] * p.col[c][1] + m[8] * p.col[c][2];
            float qx = wx - ox, qy = wy - oy;
            hit = hit || (qx * qx + qy * qy <= r2 && wz <= oh + p.contact_margin);
        }
    }
    if (duck_near) {
        float sx = o.dkx, sy = o.dky, sz = o.dkz + p.duck_radius;
        float ddx = e.px - sx, ddy = e.py - sy, ddz = e.pz - sz;
        float rr = p.duck_radius + reach;
        if (ddx * ddx + ddy * ddy + ddz * ddz <= rr * rr) {
            const float r2 = (p.duck_radius + p.contact_margin) * (p.duck_radius + p.contact_margin);
            for (int c = 0; c < p.n_col; ++c) {
                float wx = e.px + m[0] * p.col[c][0] + m[1] * p.col[c][1] + m[2] * p.col[c][2] - sx;
                float wy = e.py + m[3] * p.col[c][0] + m[4] * p.col[c][1] + m[5] * p.col[c][2] - sy;
                float wz = e.pz + m[6] * p.col[c][0] + m[7] * p.col[c][1] + m[8] * p.col[c][2] - sz;
                hit = hit || (wx * wx + wy * wy + wz * wz <= r2);
            }
        }
    }
    return hit;
}

// _reset_duck_phase_state + _spawn_duck + _spawn_obstacles (targets must already be sampled)
__device__ __forceinline__ void ol_reset(const FwDev& p, const FwPlanes& pl, OlState& o, int i, uint32_t gid, uint32_t episode,
                                         float* so, int tid, int stride) {
    o.duck_phase = 0; o.seen = 0; o.lock = 0; o.has_prev = 0; o.prev_est = 0.0f;
    o.last_cx = 0.5f; o.last_cy = 0.5f; o.last_area = 0.0f; o.last_depth = 0.0f;
    o.since = 60; o.post_wp = 0; o.cam_valid = 0; o.f_visible = 0;
    o.f_cx = o.f_cy = o.f_area = o.f_depth = 0.0f; o.f_dl = o.f_dc = o.f_dr = 0.0f;
    o.vis_dl = o.vis_dc = o.vis_dr = 0.0f; o.vis_flag = 0.0f;
    if (p.num_targets > 0) {
        const int t = p.num_targets - 1;
        o.dkx = pl.targets[(size_t)(t * 3 + 0) * p.n + i];
        o.dky = pl.targets[(size_t)(t * 3 + 1) * p.n + i];
    } else { o.dkx = 10.0f; o.dky = 0.0f; }
    o.dkz = 0.05f;
    int n = 0;
    for (int k = 0; k < p.num_obstacles; ++k) {
        uint4 r = fw_philox(p.seed_lo, p.seed_hi, gid, episode, (uint32_t)k, FWD_STREAM_OBST);
        float h = p.obst_h_lo + fw_u01(r.x) * (p.obst_h_hi - p.obst_h_lo);
        float x = -0.5f * p.dome + fw_u01(r.y) * p.dome;
        float y = -0.5f * p.dome + fw_u01(r.z) * p.dome;
        if (x * x + y * y < 100.0f) continue;
        OL_S(so, n, 0, tid, stride) = x; OL_S(so, n, 1, tid, stride) = y; OL_S(so, n, 2, tid, stride) = h;
        pl.obst[(size_t)(n * 3 + 0) * p.n + i] = x;
        pl.obst[(size_t)(n * 3 + 1) * p.n + i] = y;
        pl.obst[(size_t)(n * 3 + 2) * p.n + i] = h;
        ++n;
    }
    o.n_obst = n;
}

// ================================================================== duck-only task (FixedwingObjLockEnv, task 4)
// Restates /root/reference/envs/fixedwing_objlock_env.py: reset/_spawn_duck/_spawn_obstacles :177-251,461-578,
// compute_state :253-287, _build_duck_vision_observation :421-459, compute_term_trunc_reward :289-372 and
// FlattenObjLockEnv._flatten_obs (/root/reference/envs/flatten_objlock_env.py:41-46).
// The vision history (hist_len x 9 floats + 4 deltas per env) lives in shared memory during a step:
// hs[slot * stride + tid], slot = h * 9 + c for history row h (0 = newest), then the deltas.  `o.seen` carries
// _vision_history_filled in this task (the waypoint task's consecutive-seen counter has no use here).
#define OL_H(hs, slot, tid, stride) (hs)[(slot) * (stride) + (tid)]

__device__ __forceinline__ void ol_hist_load(const FwDev& p, const FwPlanes& pl, int i, float* hs, int tid, int stride) {
    for (int k = 0; k < p.hist_slots; ++k) OL_H(hs, k, tid, stride) = pl.hist[(size_t)k * p.n + i];
}
__device__ __forceinline__ void ol_hist_store(const FwDev& p, const FwPlanes& pl, int i, const float* hs, int tid, int stride) {
    for (int k = 0; k < p.hist_slots; ++k) pl.hist[(size_t)k * p.n + i] = OL_H(hs, k, tid, stride);
}

// _build_duck_vision_observation: newest feature vector in, oldest out, deltas against the previous newest row
__device__ __forceinline__ void ol_hist_push(const FwDev& p, OlState& o, float* hs, int tid, int stride) {
    const int H = p.hist_len;
    float base[9];
    base[0] = o.vis_flag; base[1] = o.last_cx; base[2] = o.last_cy; base[3] = o.last_area; base[4] = o.last_depth;
    base[5] = (float)o.since * (1.0f / 60.0f);
    base[6] = o.vis_dl; base[7] = o.vis_dc; base[8] = o.vis_dr;
    float prev[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) prev[c] = H >= 2 ? OL_H(hs, c, tid, stride) : 0.0f;
    for (int h = H - 1; h >= 1; --h)
#pragma unroll
        for (int c = 0; c < 9; ++c) OL_H(hs, h * 9 + c, tid, stride) = OL_H(hs, (h - 1) * 9 + c, tid, stride);
#pragma unroll
    for (int c = 0; c < 9; ++c) OL_H(hs, c, tid, stride) = base[c];
    o.seen = min(o.seen + 1, H);
    const bool both = o.seen >= 2 && base[0] > 0.5f && prev[0] > 0.5f;
#pragma unroll
    for (int c = 0; c < 4; ++c) OL_H(hs, 9 * H + c, tid, stride) = both ? base[1 + c] - prev[1 + c] : 0.0f;
}

// compute_term_trunc_reward past the early return on crash (:296-372).  `dist` = |target_vector|.
__device__ __forceinline__ void ol_duck_reward(const FwDev& p, const EnvState& e, OlState& o, float& reward, bool& term,
                                               bool& complete, bool& strike) {
    reward -= ol_obstacle_penalty(p, o, true);                      // always the halved scale (:403)
    const float ddx = o.dkx - e.px, ddy = o.dky - e.py, ddz = o.dkz - e.pz;
    const float dist = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
    if (!p.sparse_reward) {
        reward += p.duck_dist_scale / fmaxf(dist, 2.0f);
        if (o.vis_flag > 0.5f) {
            const float est = o.last_depth;
            reward += p.visible_step_reward;
            reward += p.area_reward_scale * fmaxf(0.0f, o.last_area);
            const float cxd = o.last_cx - 0.5f, cyd = o.last_cy - 0.5f;
            const float dc = sqrtf(cxd * cxd + cyd * cyd);
            const float rl = fmaxf(p.lock_center_radius, 1e-6f);
            reward += p.centering_scale * fmaxf(0.0f, (rl - dc) / rl);
            if (dc < rl) { o.lock = min(o.lock + 1, p.lock_hold); reward += p.lock_step_reward; }
            else o.lock = max(o.lock - p.lock_decay, 0);
            if (o.has_prev && est > 0.0f) {
                float diff = o.prev_est - est;
                if (p.approach_clip > 0.0f) diff = fminf(fmaxf(diff, -p.approach_clip), p.approach_clip);
                reward += diff * p.approach_scale;
            }
            if (est > 0.0f) { o.prev_est = est; o.has_prev = 1; } else { o.prev_est = 0.0f; o.has_prev = 0; }
        } else {
            if (o.lock > 0) reward -= p.lock_lost_penalty;
            o.lock = max(o.lock - p.lock_decay, 0);
            o.prev_est = 0.0f; o.has_prev = 0;
        }
    }
    if (o.lock >= p.lock_hold && dist <= p.strike_dist) {
        term = true; reward += p.strike_reward; complete = true; strike = true;
    }
}

// the part of the flattened observation after the attitude: target_vector (duck relative to the aircraft, body
// frame), the history rows and the deltas
__device__ __forceinline__ void ol_write_obs_duck_tail(const FwDev& p, const EnvState& e, const OlState& o, const float* hs,
                                                       int tid, int stride, float* tail) {
    Mat3 R = fw_quat_mat(e.qx, e.qy, e.qz, e.qw);
    const float* m = R.m;
    const float dx = o.dkx - e.px, dy = o.dky - e.py, dz = o.dkz - e.pz;
    tail[0] = m[0] * dx + m[3] * dy + m[6] * dz;
    tail[1] = m[1] * dx + m[4] * dy + m[7] * dz;
    tail[2] = m[2] * dx + m[5] * dy + m[8] * dz;
    const int nv = 9 * p.hist_len + (p.use_deltas ? 4 : 0);
    for (int k = 0; k < nv; ++k) tail[3 + k] = OL_H(hs, k, tid, stride);
}

// _reset_duck_state + _spawn_duck + _spawn_obstacles of the duck-only env
__device__ __forceinline__ void ol_reset_duck(const FwDev& p, const FwPlanes& pl, OlState& o, int i, uint32_t gid,
                                              uint32_t episode, float* so, float* hs, int tid, int stride) {
    o.duck_phase = 0; o.seen = 0; o.lock = 0; o.has_prev = 0; o.prev_est = 0.0f;
    o.last_cx = 0.5f; o.last_cy = 0.5f; o.last_area = 0.0f; o.last_depth = 0.0f;
    o.since = 60; o.post_wp = 0; o.cam_valid = 0; o.f_visible = 0;
    o.f_cx = o.f_cy = o.f_area = o.f_depth = 0.0f; o.f_dl = o.f_dc = o.f_dr = 0.0f;
    o.vis_dl = o.vis_dc = o.vis_dr = 0.0f; o.vis_flag = 0.0f;
    for (int k = 0; k < p.hist_slots; ++k) OL_H(hs, k, tid, stride) = 0.0f;
    uint4 rd = fw_philox(p.seed_lo, p.seed_hi, gid, episode, 0u, FWD_STREAM_DUCK);
    o.dkx = -0.5f * p.dome + fw_u01(rd.x) * p.dome;
    o.dky = -0.5f * p.dome + fw_u01(rd.y) * p.dome;
    o.dkz = 0.05f;
    int n = 0;
    for (int k = 0; k < p.num_obstacles; ++k) {
        uint4 r = fw_philox(p.seed_lo, p.seed_hi, gid, episode, (uint32_t)k, FWD_STREAM_OBST);
        float h = p.obst_h_lo + fw_u01(r.x) * (p.obst_h_hi - p.obst_h_lo);
        float x = -0.5f * p.dome + fw_u01(r.y) * p.dome;
        float y = -0.5f * p.dome + fw_u01(r.z) * p.dome;
        const float qx = x - o.dkx, qy = y - o.dky;
        if (qx * qx + qy * qy < 100.0f) continue;           // within 10 m of the duck
        if (x * x + y * y < 100.0f) continue;               // within 10 m of the start position
        OL_S(so, n, 0, tid, stride) = x; OL_S(so, n, 1, tid, stride) = y; OL_S(so, n, 2, tid, stride) = h;
        pl.obst[(size_t)(n * 3 + 0) * p.n + i] = x;
        pl.obst[(size_t)(n * 3 + 1) * p.n + i] = y;
        pl.obst[(size_t)(n * 3 + 2) * p.n + i] = h;
        ++n;
    }
    o.n_obst = n;
}
