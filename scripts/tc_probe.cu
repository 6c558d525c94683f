// Probe: tcgen05.mma kind::tf32 with MN-major operands from the K-major-written [128 x C] buffers.
// D[m][n] = sum_s A[s][m] * B[s][n]  (A buffer: 128 samples x 128, B buffer: 128 samples x NB)
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t off(int r, int c, int K) { return (r >> 3) * (K * 32) + (c >> 2) * 128 + (r & 7) * 16 + (c & 3) * 4; }
__device__ __forceinline__ uint64_t desc(uint32_t a, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__global__ void probe(const float* A, const float* B, float* D, int NB, int variant) {
    extern __shared__ __align__(1024) char smem[];
    __shared__ __align__(8) unsigned long long barmem;
    __shared__ uint32_t tptr;
    char* sa = smem; char* sbuf = smem + 128 * 128 * 4;
    int tid = threadIdx.x, warp = tid >> 5;
    const bool amn = variant & 1, bmn = variant & 2;
    // A logical [M=128][K=128]: A[s][m] given as samples x m.  MN-major: store sample rows (K-major layout of the [s][m] matrix);
    // K-major: store element (m, s) at off(m, s, 128)
    for (int c = 0; c < 128; ++c) {
        if (amn) *(float*)(sa + off(tid, c, 128)) = A[tid * 128 + c];
        else *(float*)(sa + off(c, tid, 128)) = A[tid * 128 + c];
    }
    for (int c = 0; c < NB; ++c) {
        if (bmn) *(float*)(sbuf + off(tid, c, NB)) = B[tid * NB + c];
        else *(float*)(sbuf + off(c, tid, 128)) = B[tid * NB + c];      // [NB x 128] K-major
    }
    uint32_t bar = s32(&barmem);
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(&tptr)), "r"(128u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem = tptr;
    if (tid == 0) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((amn ? 1u : 0u) << 15) | ((bmn ? 1u : 0u) << 16) | ((uint32_t)(NB >> 3) << 17) | (8u << 24);
        for (int k = 0; k < 16; ++k) {
            uint64_t da, db;
            da = amn ? desc(s32(sa) + k * 4096, 4096, 128) : desc(s32(sa) + k * 256, 128, 4096);
            db = bmn ? desc(s32(sbuf) + k * NB * 32, NB * 32, 128) : desc(s32(sbuf) + k * 256, 128, 4096);
            uint32_t acc = k > 0;
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    }
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
        if (ok) break;
        if (it > (1u << 22)) __trap();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t u[16];
    for (int c0 = 0; c0 < NB; c0 += 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 16; ++c) D[tid * NB + c0 + c] = __uint_as_float(u[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u));
}

__device__ __forceinline__ uint32_t off16(int r, int c, int K) { return (r >> 3) * (K * 16) + (c >> 3) * 128 + (r & 7) * 16 + (c & 7) * 2; }
__global__ void probe16(const float* A, const float* B, float* D, int NB, int variant) {
    extern __shared__ __align__(1024) char smem[];
    __shared__ __align__(8) unsigned long long barmem;
    __shared__ uint32_t tptr;
    char* sa = smem; char* sbuf = smem + 128 * 128 * 2;
    int tid = threadIdx.x, warp = tid >> 5;
    const bool amn = variant & 1, bmn = variant & 2;
    for (int c = 0; c < 128; ++c) {
        __nv_bfloat16 v = __float2bfloat16(A[tid * 128 + c]);
        if (amn) *(__nv_bfloat16*)(sa + off16(tid, c, 128)) = v; else *(__nv_bfloat16*)(sa + off16(c, tid, 128)) = v;
    }
    for (int c = 0; c < NB; ++c) {
        __nv_bfloat16 v = __float2bfloat16(B[tid * NB + c]);
        if (bmn) *(__nv_bfloat16*)(sbuf + off16(tid, c, NB)) = v; else *(__nv_bfloat16*)(sbuf + off16(c, tid, 128)) = v;
    }
    uint32_t bar = s32(&barmem);
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(&tptr)), "r"(128u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem = tptr;
    if (tid == 0) {
        // kind::f16: a/b format 1 = BF16
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((amn ? 1u : 0u) << 15) | ((bmn ? 1u : 0u) << 16) | ((uint32_t)(NB >> 3) << 17) | (8u << 24);
        for (int k = 0; k < 8; ++k) {          // K = 16 per instruction
            uint64_t da, db;
            da = amn ? desc(s32(sa) + k * 2 * (128 * 16), 128 * 16, 128) : desc(s32(sa) + k * 256, 128, 128 * 16);
            db = bmn ? desc(s32(sbuf) + k * 2 * (NB * 16), NB * 16, 128) : desc(s32(sbuf) + k * 256, 128, 128 * 16);
            uint32_t acc = k > 0;
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    }
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
        if (ok) break;
        if (it > (1u << 22)) __trap();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t u[16];
    for (int c0 = 0; c0 < NB; c0 += 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 16; ++c) D[tid * NB + c0 + c] = __uint_as_float(u[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u));
}
int main() {
    for (int use16 = 1; use16 < 2; ++use16) for (int NB : {16, 64}) for (int variant = 0; variant < 4; ++variant) {
        std::vector<float> A(128 * 128), B(128 * NB), D(128 * NB), R(128 * NB, 0.f);
        for (int i = 0; i < 128 * 128; ++i) A[i] = (float)((i * 7 + 3) % 13 - 6) * 0.25f;
        for (int i = 0; i < 128 * NB; ++i) B[i] = (float)((i * 5 + 1) % 11 - 5) * 0.5f;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < NB; ++n) { double s = 0; for (int k = 0; k < 128; ++k) s += A[k * 128 + m] * B[k * NB + n]; R[m * NB + n] = (float)s; }
        float *dA, *dB, *dD; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, D.size() * 4);
        size_t sm = 128 * 128 * 4 + 128 * 128 * 4;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaFuncSetAttribute(probe16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (use16) probe16<<<1, 128, sm>>>(dA, dB, dD, NB, variant); else probe<<<1, 128, sm>>>(dA, dB, dD, NB, variant);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, nrm = 0; int zeros = 0;
        for (size_t i = 0; i < D.size(); ++i) { err += (D[i] - R[i]) * (D[i] - R[i]); nrm += R[i] * R[i]; zeros += D[i] == 0.f; }
        printf("bf16 %d NB %d variant %d: %s rel err %.3e zeros %d/%zu  D[0..3] %g %g %g %g  R %g %g %g %g\n", use16, NB, variant, cudaGetErrorString(e), sqrt(err / nrm), zeros, D.size(), D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
