// Probe how tcgen05.mma (kind::tf32, no swizzle) addresses an MN-major B operand:
// A (K-major, 128 x 8) selects kk = m % 8;  B's memory holds its own float index, so D[m][n] = index read for B[n][m%8].
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ uint64_t desc(uint32_t a, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__global__ void probe(float *out, uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int a_mn) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float *A = (float *)sm;                 // 16 KB region
    float *B = (float *)(sm + 16384);       // 16 KB region
    uint64_t *mbar = (uint64_t *)(sm + 32768);
    uint32_t *tm = (uint32_t *)(sm + 32768 + 16);
    int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += blockDim.x) { A[i] = 0.f; B[i] = (float)i; }
    __syncthreads();
    if (!a_mn) {   // A K-major: element (m, k) at (m/8)*1024 + (k/4)*128 + (m%8)*16 + (k%4)*4 ; A[m][m%8] = 1
        for (int m = tid; m < 128; m += blockDim.x) { int k = m % 8; A[((m >> 3) * 1024 + (k >> 2) * 128 + (m & 7) * 16 + (k & 3) * 4) / 4] = 1.f; }
    } else {       // probing A as MN-major: A memory holds its index, B K-major selects kk = n % 8
        for (int i = tid; i < 4096; i += blockDim.x) { A[i] = (float)i; B[i] = 0.f; }
        __syncthreads();
        for (int n = tid; n < 32; n += blockDim.x) { int k = n % 8; B[((n >> 3) * 1024 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4) / 4] = 1.f; }
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tm)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(mbar)) : "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem = *(volatile uint32_t *)tm;
    if (tid == 0) {
        uint64_t da = desc(s32(A), a_lbo, a_sbo), db = desc(s32(B), b_lbo, b_sbo);
        // baseline: both K-major -> D = known pattern; then the probed descriptor pair accumulates on top
        uint64_t ka = desc(s32(A), 128, 1024), kb = desc(s32(B), 128, 1024);
        const uint32_t base = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem), "l"(ka), "l"(kb), "r"(base), "r"(0) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s32(mbar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}\n" :: "r"(s32(mbar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int warp = tid >> 5, lane = tid & 31;
    for (int c = 0; c < 32; c += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(32 * warp) << 16) + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[(32 * warp + lane) * 32 + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(32) : "memory");
}
int main() {
    float *out; cudaMalloc(&out, 128 * 32 * 4);
    float h[128 * 32];
    const uint32_t base = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    struct { const char *name; uint32_t idesc, albo, asbo, blbo, bsbo; int a_mn; } cfg[] = {
        {"B MN-major lbo=1024 sbo=128", base | (1u << 16), 128, 1024, 1024, 128, 0},
        {"B MN-major lbo=128 sbo=1024", base | (1u << 16), 128, 1024, 128, 1024, 0},
        {"B K-major  lbo=128 sbo=1024 (reference)", base, 128, 1024, 128, 1024, 0},
        {"A MN-major lbo=1024 sbo=128", base | (1u << 15), 1024, 128, 128, 1024, 1},
        {"A MN-major lbo=128 sbo=1024", base | (1u << 15), 128, 1024, 128, 1024, 1},
    };
    for (auto &c : cfg) {
        cudaMemset(out, 0, sizeof(h));
        probe<<<1, 128, 40000>>>(out, c.idesc, c.albo, c.asbo, c.blbo, c.bsbo, c.a_mn);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("== %s (%s)\n", c.name, cudaGetErrorString(e));
        if (!c.a_mn) {   // rows m = kk (0..7), columns n: which float index was read for B[n][kk]
            for (int m = 0; m < 2; ++m) { printf(" kk=%d:", m); for (int n = 0; n < 12; ++n) printf(" %5.0f", h[m * 32 + n]); printf(" ... n=31: %5.0f\n", h[m * 32 + 31]); }
        } else {         // D[m][n] = A[m][kk = n%8]: index read for A[m][kk]
            for (int n = 0; n < 2; ++n) { printf(" kk=%d:", n); for (int m = 0; m < 12; ++m) printf(" %5.0f", h[m * 32 + n]); printf(" ... m=31: %5.0f m=127: %5.0f\n", h[31 * 32 + n], h[127 * 32 + n]); }
        }
    }
    return 0;
}
