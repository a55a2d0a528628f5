// umma_time3.cu -- cycles per tcgen05.mma with a clean issue path: warp-uniform branch, elect.sync, descriptors
// precomputed, issue sequence unrolled.  (umma_time.cu measured its own descriptor arithmetic: ~118 cycles / MMA.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_time3 umma_time3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint64_t mkdesc(uint32_t addr, uint32_t lt, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
template <int KIND, int TS>   // KIND 0 tf32, 1 bf16
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint32_t a_tmem, uint64_t b, uint32_t id) {
    if (KIND == 0) {
        if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" :: "r"(d), "r"(a_tmem), "l"(b), "r"(id) : "memory");
        else    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(d), "l"(a), "l"(b), "r"(id) : "memory");
    } else {
        if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" :: "r"(d), "r"(a_tmem), "l"(b), "r"(id) : "memory");
        else    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(d), "l"(a), "l"(b), "r"(id) : "memory");
    }
}
__device__ __forceinline__ void commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
struct TCfg { int M, N, a_lt, a_lbo, a_sbo, b_lt, b_lbo, b_sbo, tr, a_step, b_step, iters, count; };

template <int KIND, int TS>
__global__ void k_time(long long *out, TCfg c) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t *mbar = (uint64_t *)(sm + 131072);
    uint32_t *tm = (uint32_t *)(sm + 131072 + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32768; i += blockDim.x) ((float *)sm)[i] = 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tm)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(mbar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)tm;
    long long t0 = 0, t1 = 0, ti = 0;
    if (warp == 0) {
        uint32_t id = (1u << 4) | ((KIND ? 1u : 2u) << 7) | ((KIND ? 1u : 2u) << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
        if (c.tr) id |= (1u << 15) | (1u << 16);
        const uint32_t abase = s32(sm), bbase = s32(sm) + 65536;
        uint64_t a[4], b[4];
        uint32_t d[4], at[4];
        for (int i = 0; i < 4; ++i) {
            a[i] = mkdesc(abase + i * c.a_step, c.a_lt, c.a_lbo, c.a_sbo);
            b[i] = mkdesc(bbase + i * c.b_step, c.b_lt, c.b_lbo, c.b_sbo);
            d[i] = tmem + (uint32_t)((i * c.N) % 256);
            at[i] = tmem + 384 + i * 8;
        }
        if (elect_one()) {
            t0 = clock64();
            if (c.count > 0) {                 // queue-depth mode: exactly c.count MMAs, no loop-carried arithmetic
#pragma unroll
                for (int u = 0; u < 64; ++u) if (u < c.count) mma<KIND, TS>(d[u & 3], a[u & 3], at[u & 3], b[(u >> 2) & 3], id);
            } else
            for (int it = 0; it < c.iters; ++it) {
#pragma unroll
                for (int u = 0; u < 16; ++u) mma<KIND, TS>(d[u & 3], a[u & 3], at[u & 3], b[(u >> 2) & 3], id);
            }
            commit(s32(mbar));
            ti = clock64();
        }
        __syncwarp();
    }
    if (warp == 0) mbar_wait(s32(mbar), 0);
    t1 = clock64();
    __syncwarp();
    if (t0 != 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

template <int KIND, int TS>
void run(const char *name, TCfg c, long long *dt) {
    cudaFuncSetAttribute(k_time<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 64);
    long long r4[2], r68[2];
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 2; ++rep) {
        c.iters = 4;  k_time<KIND, TS><<<1, 128, 131072 + 64>>>(dt, c); cudaDeviceSynchronize(); cudaMemcpy(r4, dt, 16, cudaMemcpyDeviceToHost);
        c.iters = 68; k_time<KIND, TS><<<1, 128, 131072 + 64>>>(dt, c); e = cudaDeviceSynchronize(); cudaMemcpy(r68, dt, 16, cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); exit(1); }
    printf("%s: %7.2f cycles/MMA  (issue %6.2f/MMA; 64 MMAs: %lld cycles)\n", name, (double)(r68[0] - r4[0]) / 1024.0, (double)(r68[1] - r4[1]) / 1024.0, r4[0]);
}

template <int KIND, int TS>
void depth(const char *name, TCfg c, long long *dt) {
    cudaFuncSetAttribute(k_time<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 64);
    printf("%s: issue cycles / total cycles for R MMAs:", name);
    for (int R : {1, 2, 4, 6, 8, 12, 16, 24, 32, 48, 64}) {
        long long r[2];
        c.count = R; c.iters = 0;
        for (int rep = 0; rep < 2; ++rep) { k_time<KIND, TS><<<1, 128, 131072 + 64>>>(dt, c); cudaDeviceSynchronize(); }
        cudaMemcpy(r, dt, 16, cudaMemcpyDeviceToHost);
        printf("  R=%d %lld/%lld", R, r[1], r[0]);
    }
    printf("\n");
}

int main() {
    long long *dt; cudaMalloc(&dt, 64);
    depth<1, 0>("queue depth, bf16 SS MN/MN sw128 M64 N64 (32 cycles each)", {64, 64, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0, 0}, dt);
    depth<0, 1>("queue depth, tf32 TS M128 N32 (16 cycles each)          ", {128, 32, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    depth<0, 0>("queue depth, tf32 SS M128 N32 (40 cycles each)          ", {128, 32, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    depth<0, 0>("queue depth, tf32 SS M128 N256 (128 cycles each)        ", {128, 256, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    //                                        M    N  a_lt lbo  sbo  b_lt lbo  sbo  tr a_step b_step
    run<0, 0>("tf32 SS none  M128 N16 ", {128, 16, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<0, 0>("tf32 SS none  M128 N32 ", {128, 32, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<0, 0>("tf32 SS none  M128 N64 ", {128, 64, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<0, 0>("tf32 SS none  M128 N128", {128, 128, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<0, 0>("tf32 SS none  M128 N256", {128, 256, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<0, 0>("tf32 SS none  M64  N32 ", {64, 32, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<0, 0>("tf32 SS none  M64  N64 ", {64, 64, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<0, 0>("tf32 SS sw128 M128 N32 ", {128, 32, 2, 16, 1024, 2, 16, 1024, 0, 32, 32, 0, 0}, dt);
    run<0, 0>("tf32 SS sw128 M128 N64 ", {128, 64, 2, 16, 1024, 2, 16, 1024, 0, 32, 32, 0, 0}, dt);
    run<0, 0>("tf32 SS sw128 M128 N256", {128, 256, 2, 16, 1024, 2, 16, 1024, 0, 32, 32, 0, 0}, dt);
    run<0, 1>("tf32 TS       M128 N32 ", {128, 32, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<0, 1>("tf32 TS       M128 N64 ", {128, 64, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<0, 1>("tf32 TS       M128 N256", {128, 256, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<1, 0>("bf16 SS none  M128 N32 ", {128, 32, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<1, 0>("bf16 SS none  M128 N64 ", {128, 64, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<1, 0>("bf16 SS none  M128 N96 ", {128, 96, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<1, 0>("bf16 SS none  M128 N128", {128, 128, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<1, 0>("bf16 SS none  M128 N256", {128, 256, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<1, 0>("bf16 SS sw128 M128 N32 ", {128, 32, 2, 16, 1024, 2, 16, 1024, 0, 32, 32, 0, 0}, dt);
    run<1, 0>("bf16 SS sw128 M128 N256", {128, 256, 2, 16, 1024, 2, 16, 1024, 0, 32, 32, 0, 0}, dt);
    run<1, 1>("bf16 TS       M128 N32 ", {128, 32, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0, 0}, dt);
    run<1, 0>("bf16 SS MN/MN sw128 M128 N64 ", {128, 64, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0, 0}, dt);
    run<1, 0>("bf16 SS MN/MN sw128 M128 N96 ", {128, 96, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0, 0}, dt);
    run<1, 0>("bf16 SS MN/MN sw128 M128 N128", {128, 128, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0, 0}, dt);
    run<1, 0>("bf16 SS MN/MN sw128 M64  N64 ", {64, 64, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0, 0}, dt);
    run<1, 0>("bf16 SS MN/MN sw128 M64  N32 ", {64, 32, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0, 0}, dt);
    return 0;
}
