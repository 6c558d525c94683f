// f32x2_probe.cu -- does the packed fp32x2 path of sm_100a (FFMA2 / FMUL2 / FADD2) relieve an ISSUE-bound kernel?
// (a) scalar FFMA chains, (b) FFMA2 chains, (c) FFMA + one ALU-pipe op per FMA, (d) FFMA2 + the same ALU ops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/f32x2_probe scripts/f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float2 x[8];
    int m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i); m[i] = threadIdx.x + i; }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0 || MODE == 2) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
                else x[i] = __ffma2_rn(x[i], a2, b2);
                if (MODE >= 2) { m[i] = (m[i] ^ (m[i] >> 3)) + it; m[(i + 1) & 7] = max(m[(i + 1) & 7], m[i]); }
            }
        }
    }
    float s = 0; int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i].x + x[i].y; t += m[i]; }
    if (s == 123.456f || t == 12345) out[0] = s + t;
}

template <int MODE>
void run(const char* name, int sms) {
    float* d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(d, 100, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fma = (double)blocks * 256 * iters * 4 * 8 * 2;   // scalar FMAs
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  (%.1f G fma-lane-ops/s)\n", name, best, fma * 2 / best / 1e9, fma / best / 1e6);
    cudaFree(d);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("scalar FFMA x16 chains", sms);
    run<1>("FFMA2 x8 chains", sms);
    run<2>("scalar FFMA + ALU ops", sms);
    run<3>("FFMA2 + ALU ops", sms);
    return 0;
}
