// umma_time.cu -- cycles per tcgen05.mma on the B200 for the shapes / operand sources the SIREN kernels can use.
// A whole warp enters the issue branch (lane 0 issues, lanes 1..31 wait at __syncwarp) so that no lane of the
// issuing warp spins on the mbarrier while lane 0 is still issuing (that pattern costs ~260 cycles per MMA).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_time umma_time.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint64_t mkdesc(uint32_t addr, uint32_t lt, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
}
#define MMA(kind, d, a, b, id) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::" kind " [%0], %1, %2, %3, p;\n\t}\n" :: "r"(d), "l"(a), "l"(b), "r"(id) : "memory")
#define MMA_TS(kind, d, a, b, id) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::" kind " [%0], [%1], %2, %3, p;\n\t}\n" :: "r"(d), "r"(a), "l"(b), "r"(id) : "memory")
__device__ __forceinline__ void commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
struct TCfg { int M, N, bf16, ts, ndst, a_lt, a_lbo, a_sbo, b_lt, b_lbo, b_sbo, tr, a_step, b_step, R, mode; };

__global__ void k_time(long long *out, TCfg c) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t *mbar = (uint64_t *)(sm + 131072);
    uint32_t *tm = (uint32_t *)(sm + 131072 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
        if (lane == 0) {
            uint32_t id = (1u << 4) | ((c.bf16 ? 1u : 2u) << 7) | ((c.bf16 ? 1u : 2u) << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
            if (c.tr) id |= (1u << 15) | (1u << 16);
            const uint32_t abase = s32(sm), bbase = s32(sm) + 65536;
            t0 = clock64();
            for (int r = 0; r < c.R; ++r) {
                const uint32_t d = tmem + (uint32_t)((r & (c.ndst - 1)) * c.N);
                const uint64_t a = mkdesc(abase + (r & 3) * c.a_step, c.a_lt, c.a_lbo, c.a_sbo);
                const uint64_t b = mkdesc(bbase + (r & 3) * c.b_step, c.b_lt, c.b_lbo, c.b_sbo);
                if (c.bf16) { if (c.ts) MMA_TS("f16", d, tmem + 384 + (r & 3) * 8, b, id); else MMA("f16", d, a, b, id); }
                else        { if (c.ts) MMA_TS("tf32", d, tmem + 384 + (r & 3) * 8, b, id); else MMA("tf32", d, a, b, id); }
            }
            commit(s32(mbar));
            ti = clock64();
        }
        __syncwarp();
    }
    if (c.mode == 0 || warp == 0) mbar_wait(s32(mbar), 0);          // mode 0: every warp polls; mode 1: only warp 0 polls
    t1 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

int main() {
    long long *dt; cudaMalloc(&dt, 64);
    cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 64);
    struct { const char *name; TCfg c; } tc[] = {
        //                                       M    N  bf ts nd  a_lt lbo  sbo   b_lt lbo  sbo  tr a_step b_step
        {"tf32 SS none  M128 N32        ", {128, 32, 0, 0, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 SS none  M128 N32 nd1    ", {128, 32, 0, 0, 1, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 SS none  M128 N64        ", {128, 64, 0, 0, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 SS none  M128 N96        ", {128, 96, 0, 0, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 SS none  M128 N128       ", {128, 128, 0, 0, 2, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 SS none  M128 N256       ", {128, 256, 0, 0, 1, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 SS none  M64  N32        ", {64, 32, 0, 0, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 SS none  M64  N64        ", {64, 64, 0, 0, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 SS sw128 M128 N32        ", {128, 32, 0, 0, 4, 2, 16, 1024, 2, 16, 1024, 0, 32, 32, 0}},
        {"tf32 SS sw128 M128 N64        ", {128, 64, 0, 0, 4, 2, 16, 1024, 2, 16, 1024, 0, 32, 32, 0}},
        {"tf32 SS sw128 M128 N128       ", {128, 128, 0, 0, 2, 2, 16, 1024, 2, 16, 1024, 0, 32, 32, 0}},
        {"tf32 TS       M128 N32        ", {128, 32, 0, 1, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 TS       M128 N64        ", {128, 64, 0, 1, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"tf32 TS sw128B M128 N32       ", {128, 32, 0, 1, 4, 0, 128, 1024, 2, 16, 1024, 0, 256, 32, 0}},
        {"bf16 SS none  M128 N32        ", {128, 32, 1, 0, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"bf16 SS none  M128 N64        ", {128, 64, 1, 0, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"bf16 SS none  M128 N96        ", {128, 96, 1, 0, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"bf16 SS none  M128 N128       ", {128, 128, 1, 0, 2, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"bf16 SS none  M128 N256       ", {128, 256, 1, 0, 1, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"bf16 SS sw128 M128 N32        ", {128, 32, 1, 0, 4, 2, 16, 1024, 2, 16, 1024, 0, 32, 32, 0}},
        {"bf16 TS       M128 N32        ", {128, 32, 1, 1, 4, 0, 128, 1024, 0, 128, 1024, 0, 256, 256, 0}},
        {"bf16 SS MN/MN sw128 M128 N64  ", {128, 64, 1, 0, 4, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0}},
        {"bf16 SS MN/MN sw128 M128 N96  ", {128, 96, 1, 0, 4, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0}},
        {"bf16 SS MN/MN sw128 M128 N128 ", {128, 128, 1, 0, 2, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0}},
        {"bf16 SS MN/MN sw128 M64  N64  ", {64, 64, 1, 0, 4, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0}},
        {"bf16 SS MN/MN sw128 M64  N32  ", {64, 32, 1, 0, 4, 2, 8192, 1024, 2, 8192, 1024, 1, 2048, 2048, 0}},
    };
    for (int mode = 0; mode < 3; ++mode)
    for (auto &t : tc) {
        t.c.mode = mode; const int nthr = mode == 2 ? 32 : 128;
        long long r1[2], r64[2], r1k[2];
        cudaError_t e = cudaSuccess;
        for (int rep = 0; rep < 2; ++rep) {
            t.c.R = 1;    k_time<<<1, nthr, 131072 + 64>>>(dt, t.c); cudaDeviceSynchronize(); cudaMemcpy(r1, dt, 16, cudaMemcpyDeviceToHost);
            t.c.R = 64;   k_time<<<1, nthr, 131072 + 64>>>(dt, t.c); cudaDeviceSynchronize(); cudaMemcpy(r64, dt, 16, cudaMemcpyDeviceToHost);
            t.c.R = 1088; k_time<<<1, nthr, 131072 + 64>>>(dt, t.c); e = cudaDeviceSynchronize(); cudaMemcpy(r1k, dt, 16, cudaMemcpyDeviceToHost);
        }
        if (e != cudaSuccess) { printf("%s: error %s\n", t.name, cudaGetErrorString(e)); return 1; }
        printf("mode %d %s: %7.2f cycles/MMA  (issue %6.2f/MMA;  R=1 latency %lld, R=64 %lld)\n", mode, t.name, (double)(r1k[0] - r64[0]) / 1024.0,
               (double)(r1k[1] - r64[1]) / 1024.0, r1[0], r64[0]);
    }
    return 0;
}
