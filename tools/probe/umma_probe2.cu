// umma_probe2.cu -- measured answers about tcgen05.mma on the B200 that the kernel design depends on:
//   G  : which shared-memory address does the MMA read for operand element (row, kk) under a given
//        descriptor (layout type / LBO / SBO / start offset) and major-ness?  (index-coded operand)
//   TS : A operand taken from TMEM (written with tcgen05.st), kind::tf32
//   T  : cycles per MMA for several (M, N) shapes, tf32 and bf16
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe2 umma_probe2.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint64_t mkdesc(uint32_t addr, uint32_t lt, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

struct GCfg { int probeA; uint32_t idesc; uint32_t lt, lbo, sbo, off; int code; int M; };

// shared: [0,16K) X = probed operand (index coded), [16K,32K) onehot operand (K-major no swizzle), [32K,48K) zeros
__global__ void k_probe(float *out, GCfg c) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float *X = (float *)sm, *Hot = (float *)(sm + 16384), *Z = (float *)(sm + 32768);
    uint64_t *mbar = (uint64_t *)(sm + 49152);
    uint32_t *tm = (uint32_t *)(sm + 49152 + 16);
    const int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += blockDim.x) { X[i] = c.code ? (float)(i & 3) : (float)(i >> 2); Hot[i] = 0.f; Z[i] = 0.f; }
    __syncthreads();
    // one-hot: element (r, k = r % 8) = 1, K-major no swizzle (LBO 128, SBO 1024), 128 rows
    for (int r = tid; r < 128; r += blockDim.x) { int k = r % 8; Hot[((r >> 3) * 1024 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4) / 4] = 1.f; }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tm)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(mbar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)tm;
    if (tid == 0) {
        const uint64_t hot = mkdesc(s32(Hot), 0, 128, 1024), zero = mkdesc(s32(Z), 0, 128, 1024);
        const uint64_t x = mkdesc(s32(X) + c.off, c.lt, c.lbo, c.sbo);
        const uint32_t base = idesc_tf32(128, 32);
        mma_tf32_ss(tmem, hot, zero, base, 0);                        // D = 0
        if (c.probeA) mma_tf32_ss(tmem, x, hot, c.idesc, 1);          // D[m][n] = X_A(m, n % 8)
        else          mma_tf32_ss(tmem, hot, x, c.idesc, 1);          // D[m][n] = X_B(n, m % 8)
        commit(s32(mbar));
    }
    mbar_wait(s32(mbar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int cc = 0; cc < 32; cc += 8) {
        uint32_t r[8];
        ld8(tmem + ((uint32_t)(32 * warp) << 16) + cc, r);
        for (int i = 0; i < 8; ++i) out[(32 * warp + lane) * 32 + cc + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64) : "memory");
}

// bf16 control: B MN-major no swizzle; A one-hot K-major (K = 16: element (m, k = m % 16))
__global__ void k_probe_bf16(float *out, uint32_t idesc, uint32_t lt, uint32_t lbo, uint32_t sbo) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint16_t *X = (uint16_t *)sm, *Hot = (uint16_t *)(sm + 16384), *Z = (uint16_t *)(sm + 32768);
    uint64_t *mbar = (uint64_t *)(sm + 49152);
    uint32_t *tm = (uint32_t *)(sm + 49152 + 16);
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += blockDim.x) {
        float v = (float)(((i >> 3) & 127) + 1);              // 16-byte chunk index mod 128, +1 (exact in bf16)
        X[i] = (uint16_t)(__float_as_uint(v) >> 16); Hot[i] = 0; Z[i] = 0;
    }
    __syncthreads();
    for (int r = tid; r < 128; r += blockDim.x) { int k = r % 16; Hot[((r >> 3) * 1024 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2) / 2] = 0x3F80; }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tm)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(mbar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)tm;
    if (tid == 0) {
        const uint64_t hot = mkdesc(s32(Hot), 0, 128, 1024), zero = mkdesc(s32(Z), 0, 128, 1024);
        mma_f16_ss(tmem, hot, zero, idesc_bf16(128, 32), 0);
        mma_f16_ss(tmem, hot, mkdesc(s32(X), lt, lbo, sbo), idesc, 1);
        commit(s32(mbar));
    }
    mbar_wait(s32(mbar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int cc = 0; cc < 32; cc += 8) {
        uint32_t r[8];
        ld8(tmem + ((uint32_t)(32 * warp) << 16) + cc, r);
        for (int i = 0; i < 8; ++i) out[(32 * warp + lane) * 32 + cc + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64) : "memory");
}

// TS: A[m][k] = (k == m % 8) + 0.5 (k == (m + 3) % 8) written to TMEM columns 32..39 by the owning threads; B K-major index coded
__global__ void k_probe_ts(float *out, int code) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float *X = (float *)sm;
    uint64_t *mbar = (uint64_t *)(sm + 49152);
    uint32_t *tm = (uint32_t *)(sm + 49152 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 4096; i += blockDim.x) X[i] = code ? (float)(i & 3) : (float)(i >> 2);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tm)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(mbar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)tm;
    {
        const int m = 32 * warp + lane;
        uint32_t r[8];
        for (int k = 0; k < 8; ++k) r[k] = __float_as_uint((k == m % 8) ? 1.f : 0.f);
        st8(tmem + ((uint32_t)(32 * warp) << 16) + 32, r);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mma_tf32_ts(tmem, tmem + 32, mkdesc(s32(X), 0, 128, 1024), idesc_tf32(128, 32), 0);
        commit(s32(mbar));
    }
    mbar_wait(s32(mbar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int cc = 0; cc < 32; cc += 8) {
        uint32_t r[8];
        ld8(tmem + ((uint32_t)(32 * warp) << 16) + cc, r);
        for (int i = 0; i < 8; ++i) out[(32 * warp + lane) * 32 + cc + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64) : "memory");
}

// timing: R back-to-back MMAs (same operands, accumulate) issued by one thread; cycles from first issue to completion
__global__ void k_time(long long *out, int M, int N, int bf16, int R, int ts, int ndst) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t *mbar = (uint64_t *)(sm + 65536);
    uint32_t *tm = (uint32_t *)(sm + 65536 + 16);
    const int tid = threadIdx.x;
    for (int i = tid; i < 16384; i += blockDim.x) ((float *)sm)[i] = 0.f;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tm)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(mbar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)tm;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint64_t a = mkdesc(s32(sm), 0, 128, 1024), b = mkdesc(s32(sm) + 32768, 0, 128, 1024);
        const uint32_t id = bf16 ? idesc_bf16(M, N) : idesc_tf32(M, N);
        t0 = clock64();
        for (int r = 0; r < R; ++r) {
            const uint32_t d = tmem + (uint32_t)((r % ndst) * N) % 256;
            if (bf16) mma_f16_ss(d, a, b, id, 1);
            else if (ts) mma_tf32_ts(d, tmem + 256, b, id, 1);
            else mma_tf32_ss(d, a, b, id, 1);
        }
        commit(s32(mbar));
    }
    mbar_wait(s32(mbar), 0);
    t1 = clock64();
    if (tid == 0) { out[0] = t1 - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

static float h0[128 * 32], h1[128 * 32];
static float *dout;

static void run_g(const char *name, GCfg c) {
    cudaError_t e = cudaSuccess;
    for (int code = 0; code < 2; ++code) {
        c.code = code;
        cudaMemset(dout, 0, sizeof(h0));
        k_probe<<<1, 128, 50000>>>(dout, c);
        e = cudaDeviceSynchronize();
        cudaMemcpy(code ? h1 : h0, dout, sizeof(h0), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) break;
    }
    printf("== %s : %s\n", name, cudaGetErrorString(e));
    if (e != cudaSuccess) { exit(1); }
    // byte address of element (row, kk): probeA: D[m][n<8] -> (row m, kk n); probeB: D[m<8][n] -> (row n, kk m)
    const int rows = c.probeA ? 128 : 32;
    for (int kk = 0; kk < 8; ++kk) {
        printf("  kk=%d:", kk);
        for (int r = 0; r < rows; ++r) {
            if (c.probeA && !(r < 12 || (r % 8 == 0))) continue;
            const int idx = c.probeA ? r * 32 + kk : kk * 32 + r;
            printf(" %5d", (int)(h0[idx] * 16 + h1[idx] * 4));
        }
        printf("\n");
    }
}

int main() {
    cudaMalloc(&dout, sizeof(h0));
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 50000);
    cudaFuncSetAttribute(k_probe_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, 50000);
    cudaFuncSetAttribute(k_probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 50000);
    cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    const uint32_t base = idesc_tf32(128, 32);
    const uint32_t TA = 1u << 15, TB = 1u << 16;
    run_g("B K-major none lbo128 sbo1024 (control)", {0, base, 0, 128, 1024, 0, 0, 128});
    run_g("B K-major SW128 sbo1024 off0", {0, base, 2, 16, 1024, 0, 0, 128});
    run_g("B K-major SW128 sbo1024 off32 (k-step 1)", {0, base, 2, 16, 1024, 32, 0, 128});
    run_g("B K-major SW128 sbo1024 off96 (k-step 3)", {0, base, 2, 16, 1024, 96, 0, 128});
    run_g("A K-major SW128 sbo1024 off64 (k-step 2)", {1, base, 2, 16, 1024, 64, 0, 128});
    run_g("B MN-major SW128 lbo1024 sbo1024", {0, base | TB, 2, 1024, 1024, 0, 0, 128});
    run_g("B MN-major SW128 lbo4096 sbo1024 off1024 (k-group 1)", {0, base | TB, 2, 4096, 1024, 1024, 0, 128});
    run_g("A MN-major SW128 lbo4096 sbo1024", {1, base | TA, 2, 4096, 1024, 0, 0, 128});
    run_g("A MN-major SW128 M=64 lbo4096 sbo1024", {1, idesc_tf32(64, 32) | TA, 2, 4096, 1024, 0, 0, 64});
    // K-major 64-byte swizzle (16 fp32 per row, 8-row atoms of 512 B): what 16-wide K slabs of the wide family would use to
    // halve its operand tiles (two CTAs per SM for S = 3) -- not yet run on hardware
    run_g("B K-major SW64 sbo512 off0", {0, base, 4, 16, 512, 0, 0, 128});
    run_g("B K-major SW64 sbo512 off32 (k-step 1)", {0, base, 4, 16, 512, 32, 0, 128});
    run_g("A K-major SW64 sbo512 off0", {1, base, 4, 16, 512, 0, 0, 128});
    run_g("B MN-major SW64 lbo1024 sbo512", {0, base | TB, 4, 1024, 512, 0, 0, 128});
    run_g("B MN-major SW32 lbo1024 sbo256", {0, base | TB, 6, 1024, 256, 0, 0, 128});
    run_g("B MN-major none lbo1024 sbo128", {0, base | TB, 0, 1024, 128, 0, 0, 128});
    run_g("B MN-major none lbo128 sbo1024", {0, base | TB, 0, 128, 1024, 0, 0, 128});
    run_g("A K-major none M=64 (control for the M=64 TMEM layout)", {1, idesc_tf32(64, 32), 0, 128, 1024, 0, 0, 64});

    {   // bf16 control
        struct { const char *n; uint32_t lt, lbo, sbo; } cf[] = {{"bf16 B MN-major none lbo1024 sbo128", 0, 1024, 128}, {"bf16 B MN-major none lbo128 sbo1024", 0, 128, 1024},
                                                                 {"bf16 B MN-major SW128 lbo1024 sbo1024", 2, 1024, 1024}, {"bf16 B K-major none (control)", 0, 128, 1024}};
        for (int i = 0; i < 4; ++i) {
            cudaMemset(dout, 0, sizeof(h0));
            k_probe_bf16<<<1, 128, 50000>>>(dout, idesc_bf16(128, 32) | (i < 3 ? TB : 0), cf[i].lt, cf[i].lbo, cf[i].sbo);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h0, dout, sizeof(h0), cudaMemcpyDeviceToHost);
            printf("== %s : %s\n", cf[i].n, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
            for (int kk = 0; kk < 16; kk += 5) { printf("  kk=%d:", kk); for (int n = 0; n < 32; ++n) printf(" %3.0f", h0[kk * 32 + n]); printf("\n"); }
        }
    }
    {   // TS
        cudaError_t e = cudaSuccess;
        for (int code = 0; code < 2; ++code) {
            cudaMemset(dout, 0, sizeof(h0));
            k_probe_ts<<<1, 128, 50000>>>(dout, code);
            e = cudaDeviceSynchronize();
            cudaMemcpy(code ? h1 : h0, dout, sizeof(h0), cudaMemcpyDeviceToHost);
        }
        printf("== TS (A from TMEM, lane = row, column = k), B K-major : %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        for (int m = 0; m < 128; m += 37) { printf("  m=%3d (kk=%d):", m, m % 8); for (int n = 0; n < 32; ++n) printf(" %5d", (int)(h0[m * 32 + n] * 16 + h1[m * 32 + n] * 4)); printf("\n"); }
    }
    {   // timing
        long long *dt; cudaMalloc(&dt, 64);
        struct { int M, N, bf16, ts, ndst; } tc[] = {{128, 32, 0, 0, 1}, {128, 32, 0, 0, 4}, {128, 16, 0, 0, 4}, {128, 64, 0, 0, 4}, {128, 128, 0, 0, 2}, {128, 256, 0, 0, 1},
                                                      {64, 32, 0, 0, 4}, {64, 64, 0, 0, 4}, {64, 128, 0, 0, 2}, {128, 32, 1, 0, 4}, {128, 64, 1, 0, 4}, {128, 32, 0, 1, 4}, {128, 64, 0, 1, 4}};
        for (auto &t : tc) {
            long long c1 = 0, c2 = 0;
            for (int rep = 0; rep < 2; ++rep) {
                k_time<<<1, 128, 70000>>>(dt, t.M, t.N, t.bf16, 64, t.ts, t.ndst); cudaDeviceSynchronize(); cudaMemcpy(&c1, dt, 8, cudaMemcpyDeviceToHost);
                k_time<<<1, 128, 70000>>>(dt, t.M, t.N, t.bf16, 1088, t.ts, t.ndst); cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&c2, dt, 8, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) { printf("timing error %s\n", cudaGetErrorString(e)); return 1; }
            }
            printf("time M=%3d N=%3d %s%s ndst=%d : %.2f cycles/MMA (64: %lld, 1088: %lld)\n", t.M, t.N, t.bf16 ? "bf16 K=16" : "tf32 K=8", t.ts ? " TS" : "", t.ndst,
                   (double)(c2 - c1) / 1024.0, c1, c2);
        }
    }
    return 0;
}
