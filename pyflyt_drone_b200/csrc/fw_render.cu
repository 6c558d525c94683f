// fw_render.cu -- debug / evaluation frame of ONE env: what FixedwingBaseEnv.render() returns through pybullet's
// getCameraImage (/root/reference/envs/fixedwing_envs/fixedwing_base_env.py:350-369; eval/eval_objlock.py:120-162 also
// keeps the segmentation mask and the depth buffer).  SURVEY.md section 8 row f4: visual debugging, not a hot path -- one
// thread per pixel, ray-cast against the scene of the analytic camera (fw_objlock.cuh): ground plane z = 0, obstacle
// cylinders, the duck sphere (camera tasks) and one sphere of radius goal_reach per waypoint not reached yet.
//   seg:   -1 sky, 0 ground, 1 duck, 2 + k obstacle k, 64 + t waypoint t
//   depth: OpenGL depth-buffer value of the hit (1.0 = far plane / sky), as getCameraImage returns it
//   rgba:  flat colour per class, Lambert-shaded with a fixed light; alpha 255
// PyBullet's rasteriser and meshes are not restated (nothing of the hot path reads these pixels); the camera pose, the
// ray family and the primitives are the ones the policy's vision features are computed from, so at W = H = cam_res the
// middle row of this frame is the row ol_capture_warp integrates.
#include "fw_device.cuh"
#include "fw_objlock.cuh"
#include "fw_kernels.h"

__global__ void __launch_bounds__(256)
fw_render_kernel(FwDev p, FwPlanes pl, int i, int W, int H, uint8_t* __restrict__ rgba, int32_t* __restrict__ seg,
                 float* __restrict__ depth) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= W * H) return;
    const int y = px / W, x = px - y * W;
    const float INF = __int_as_float(0x7f800000);
    EnvState e;
    fw_load(pl, i, e);
    const Mat3 R = fw_quat_mat(e.qx, e.qy, e.qz, e.qw);
    const float* m = R.m;
    const float ofx = m[0] * p.cam_offset[0] + m[1] * p.cam_offset[1] + m[2] * p.cam_offset[2];
    const float ofy = m[3] * p.cam_offset[0] + m[4] * p.cam_offset[1] + m[5] * p.cam_offset[2];
    const float ofz = m[6] * p.cam_offset[0] + m[7] * p.cam_offset[1] + m[8] * p.cam_offset[2];
    const float cx = e.px + ofx, cy = e.py + ofy, cz = e.pz + ofz;
    float fx, fy, fz, ux, uy, uz;
    if (p.cam_mode == 0) {
        const float il = rsqrtf(ofx * ofx + ofy * ofy + ofz * ofz);
        fx = -ofx * il; fy = -ofy * il; fz = -ofz * il;
        ux = m[2]; uy = m[5]; uz = m[8];
    } else {
        fx = m[0] * p.cam_fb[0] + m[1] * p.cam_fb[1] + m[2] * p.cam_fb[2];
        fy = m[3] * p.cam_fb[0] + m[4] * p.cam_fb[1] + m[5] * p.cam_fb[2];
        fz = m[6] * p.cam_fb[0] + m[7] * p.cam_fb[1] + m[8] * p.cam_fb[2];
        ux = m[0] * p.cam_ub[0] + m[1] * p.cam_ub[1] + m[2] * p.cam_ub[2];
        uy = m[3] * p.cam_ub[0] + m[4] * p.cam_ub[1] + m[5] * p.cam_ub[2];
        uz = m[6] * p.cam_ub[0] + m[7] * p.cam_ub[1] + m[8] * p.cam_ub[2];
    }
    float rx = fy * uz - fz * uy, ry = fz * ux - fx * uz, rz = fx * uy - fy * ux;
    const float rl = rsqrtf(rx * rx + ry * ry + rz * rz);
    rx *= rl; ry *= rl; rz *= rl;
    ux = ry * fz - rz * fy; uy = rz * fx - rx * fz; uz = rx * fy - ry * fx;
    const float xn = (2.0f * ((float)x + 0.5f) / (float)W - 1.0f) * ((float)W / (float)H);
    const float yn = 2.0f * ((float)y + 0.5f) / (float)H - 1.0f;
    const float dx = fx + xn * rx - yn * ux, dy = fy + xn * ry - yn * uy, dz = fz + xn * rz - yn * uz;

    const bool cam_task = p.task == 2 || p.task == 4, wp_task = p.task == 1 || p.task == 2;
    float best = INF;
    int id = -1;
    if (dz < -1e-12f) { const float t = -cz / dz; if (t > 0.0f) { best = t; id = 0; } }
    float hx = 0.f, hy = 0.f, hr = 1.f;            // centre (and radius) of the primitive that is hit, for the normal
    float hz = 0.f;
    if (cam_task) {
        const int n_obst = (pl.v3[i].x >> 8) & 0xff;
        for (int k = 0; k < n_obst; ++k) {
            const float ox = pl.obst[(size_t)(k * 3 + 0) * p.n + i], oy = pl.obst[(size_t)(k * 3 + 1) * p.n + i];
            const float oh = pl.obst[(size_t)(k * 3 + 2) * p.n + i];
            const float t = ol_ray_cylinder(cx, cy, cz, dx, dy, dz, ox, oy, oh, p.obst_radius);
            if (t < best) { best = t; id = 2 + k; hx = ox; hy = oy; }
        }
        const float4 dk = pl.dk[i];
        const float t = ol_ray_sphere(cx, cy, cz, dx, dy, dz, dk.x, dk.y, dk.z + p.duck_radius, p.duck_radius);
        if (t < best) { best = t; id = 1; hx = dk.x; hy = dk.y; hz = dk.z + p.duck_radius; hr = p.duck_radius; }
    }
    if (wp_task)
        for (int t = e.tidx; t < p.num_targets; ++t) {
            const float tx = pl.targets[(size_t)(t * 3 + 0) * p.n + i], ty = pl.targets[(size_t)(t * 3 + 1) * p.n + i];
            const float tz = pl.targets[(size_t)(t * 3 + 2) * p.n + i];
            const float tt = ol_ray_sphere(cx, cy, cz, dx, dy, dz, tx, ty, tz, p.goal_reach);
            if (tt < best) { best = tt; id = 64 + t; hx = tx; hy = ty; hz = tz; hr = p.goal_reach; }
        }
    // colour
    float b0 = 135.f, b1 = 206.f, b2 = 235.f, shade = 1.0f;
    if (id >= 0) {
        const float px_ = cx + best * dx, py_ = cy + best * dy, pz_ = cz + best * dz;
        float n0 = 0.f, n1 = 0.f, n2 = 1.f;
        if (id == 0) {
            // 10 m checker inside the far plane, one tone beyond it (towards the horizon the cells shrink below a pixel)
            const int chk = ((int)floorf(px_ / 10.0f) + (int)floorf(py_ / 10.0f)) & 1;
            const bool nearg = best <= p.cam_far;
            b0 = nearg ? (chk ? 96.f : 80.f) : 88.f; b1 = nearg ? (chk ? 160.f : 140.f) : 150.f; b2 = nearg ? (chk ? 96.f : 80.f) : 88.f;
        } else if (id == 1) {
            b0 = 255.f; b1 = 221.f; b2 = 0.f;
            n0 = (px_ - hx) / hr; n1 = (py_ - hy) / hr; n2 = (pz_ - hz) / hr;
        } else if (id < 64) {
            b0 = 170.f; b1 = 90.f; b2 = 70.f;
            const float ex = px_ - hx, ey = py_ - hy, rr = p.obst_radius;
            if (ex * ex + ey * ey >= rr * rr * (1.0f - 1e-3f)) { n0 = ex / rr; n1 = ey / rr; n2 = 0.f; }
        } else {
            const bool cur = id - 64 == e.tidx;
            b0 = 60.f; b1 = cur ? 220.f : 120.f; b2 = cur ? 60.f : 255.f;
            n0 = (px_ - hx) / hr; n1 = (py_ - hy) / hr; n2 = (pz_ - hz) / hr;
        }
        const float nl = n0 * 0.30151134457776363f + n1 * 0.20100756305184242f + n2 * 0.9320390859672263f;
        shade = 0.55f + 0.45f * fmaxf(nl, 0.0f);
    }
    if (rgba) {
        rgba[4 * (size_t)px + 0] = (uint8_t)(b0 * shade + 0.5f);
        rgba[4 * (size_t)px + 1] = (uint8_t)(b1 * shade + 0.5f);
        rgba[4 * (size_t)px + 2] = (uint8_t)(b2 * shade + 0.5f);
        rgba[4 * (size_t)px + 3] = 255;
    }
    if (seg) seg[px] = id;
    if (depth) {
        float d = 1.0f;
        if (best < INF) {
            const float z = fmaxf(best, p.cam_near);
            d = z > p.cam_far ? 1.0f : p.cam_far * (z - p.cam_near) / ((p.cam_far - p.cam_near) * z);
        }
        depth[px] = d;
    }
}

cudaError_t fwk_render(const FwDev& p, const FwPlanes& pl, int env, int W, int H, uint8_t* rgba, int32_t* seg, float* depth,
                       cudaStream_t st) {
    const int n = W * H;
    fw_render_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, pl, env, W, H, rgba, seg, depth);
    return cudaGetLastError();
}
