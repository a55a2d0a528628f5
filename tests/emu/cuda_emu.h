// cuda_emu.h -- a tiny SIMT emulator so the *unchanged* CUDA kernel sources of
// insr_pde_b200/csrc can be compiled with g++ and executed on the build container, which
// has nvcc but no GPU.
//
// TEST / DEBUG HARNESS ONLY.  It exists to catch indexing, layout and synchronisation
// mistakes before GPU minutes are spent.  It is compiled only by tests/emu/build_emu.py
// (-DINSR_CPU_EMU) into tests/emu/_build/libinsr_emu.so, is never loaded by the product
// package, and is NOT a CPU fallback: insr_pde_b200 refuses to run without the CUDA library.
//
// Model: blocks run one after another; the threads of a block are real OS threads;
// __syncthreads / __syncwarp / warp shuffles are barriers.  Data races inside a warp that
// CUDA would also consider races are not detected.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace insr_emu {
struct BlockCtx {
    dim3 grid, block;
    std::unique_ptr<std::barrier<>> cta_barrier;
    std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
    std::vector<std::vector<uint32_t>> warp_xchg;   // [warp][lane]
    std::vector<unsigned char> dyn_smem;
};
// per OS thread (set by launch() in the launching thread and in every kernel thread): launches issued concurrently from several
// host threads -- the ranks of the peer-memory exchange test -- do not see each other's block context
inline BlockCtx *&ctx() { static thread_local BlockCtx *c = nullptr; return c; }
struct ThreadCtx { uint3 tid, bid; };
inline ThreadCtx &tctx() { static thread_local ThreadCtx t; return t; }

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                BlockCtx c;
                c.grid = grid; c.block = block;
                c.cta_barrier.reset(new std::barrier<>(nthreads));
                for (unsigned w = 0; w < nwarps; ++w) {
                    unsigned lanes = std::min(32u, nthreads - w * 32);
                    c.warp_barrier.emplace_back(new std::barrier<>(lanes));
                    c.warp_xchg.emplace_back(32, 0u);
                }
                c.dyn_smem.assign(smem + 128, 0);
                ctx() = &c;
                std::vector<std::thread> ts;
                ts.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; ++t) {
                    ts.emplace_back([&, t]() {
                        ctx() = &c;
                        ThreadCtx &tc = tctx();
                        tc.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                        tc.bid = uint3{bx, by, bz};
                        body();
                    });
                }
                for (auto &th : ts) th.join();
                ctx() = nullptr;
            }
}
inline unsigned linear_tid() {
    auto &t = tctx(); auto *c = ctx();
    return t.tid.x + c->block.x * (t.tid.y + c->block.y * t.tid.z);
}
inline void *dyn_smem() {
    auto p = reinterpret_cast<uintptr_t>(ctx()->dyn_smem.data());
    return reinterpret_cast<void *>((p + 127) & ~uintptr_t(127));
}
}  // namespace insr_emu

#define threadIdx (insr_emu::tctx().tid)
#define blockIdx (insr_emu::tctx().bid)
#define blockDim (insr_emu::ctx()->block)
#define gridDim (insr_emu::ctx()->grid)

static inline void __syncthreads() { insr_emu::ctx()->cta_barrier->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
    insr_emu::ctx()->warp_barrier[insr_emu::linear_tid() / 32]->arrive_and_wait();
}
template <typename T>
static inline T insr_emu_shfl(T v, int src_lane) {
    static_assert(sizeof(T) == 4, "4-byte shuffles only");
    auto *c = insr_emu::ctx();
    unsigned tid = insr_emu::linear_tid(), w = tid / 32, lane = tid % 32;
    uint32_t bits; std::memcpy(&bits, &v, 4);
    c->warp_xchg[w][lane] = bits;
    c->warp_barrier[w]->arrive_and_wait();
    uint32_t got = c->warp_xchg[w][(unsigned)src_lane % 32];
    c->warp_barrier[w]->arrive_and_wait();
    T out; std::memcpy(&out, &got, 4);
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane, int = 32) { return insr_emu_shfl(v, lane); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return insr_emu_shfl(v, (int)(insr_emu::linear_tid() % 32) ^ m);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = (int)(insr_emu::linear_tid() % 32);
    return insr_emu_shfl(v, lane + (int)d < 32 ? lane + (int)d : lane);
}

static inline float atomicAdd(float *addr, float v) {
    auto *a = reinterpret_cast<std::atomic<float> *>(addr);
    float old = a->load(std::memory_order_relaxed);
    while (!a->compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {}
    return old;
}
static inline unsigned int atomicAdd(unsigned int *addr, unsigned int v) {
    return reinterpret_cast<std::atomic<unsigned int> *>(addr)->fetch_add(v, std::memory_order_acq_rel);
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned int atomicExch(unsigned int *addr, unsigned int v) {
    return reinterpret_cast<std::atomic<unsigned int> *>(addr)->exchange(v, std::memory_order_acq_rel);
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
#define __sinf(x) sinf(x)   /* glibc declares __sinf/__cosf itself: use macros */
#define __cosf(x) cosf(x)
#define __logf(x) logf(x)
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float fminf_(float a, float b) { return a < b ? a : b; }

static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrComputeCapabilityMajor = 75,
                      cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, int attr, int) {
    if (attr == cudaDevAttrMultiProcessorCount) *v = 2;        // tiny "GPU": 2 SMs
    else if (attr == cudaDevAttrComputeCapabilityMajor) *v = 10;
    else if (attr == cudaDevAttrMaxSharedMemoryPerBlockOptin) *v = 232448;
    else *v = 0;
    return cudaSuccess;
}
