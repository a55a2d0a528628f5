// insr_platform.h -- build switch between the real CUDA toolchain (the product) and the
// host-side SIMT emulator used by tests/emu (debug harness, never shipped, never loaded by
// the package).  Kernel sources use only what both provide.
#pragma once

#ifdef INSR_CPU_EMU
#include "cuda_emu.h"
#define INSR_LAUNCH(kfn, grid, block, smem, stream, ...) \
    insr_emu::launch((grid), (block), (smem), [&]() { kfn(__VA_ARGS__); })
#define INSR_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(insr_emu::dyn_smem())
#define INSR_PRAGMA_UNROLL
#define INSR_PRAGMA_UNROLL_N(n)
#else
#include <cuda_runtime.h>
#define INSR_LAUNCH(kfn, grid, block, smem, stream, ...) \
    kfn<<<(grid), (block), (smem), reinterpret_cast<cudaStream_t>(stream)>>>(__VA_ARGS__)
#define INSR_DYN_SMEM(type, name) \
    extern __shared__ __align__(128) unsigned char insr_dyn_smem_raw[]; \
    type *name = reinterpret_cast<type *>(insr_dyn_smem_raw)
#define INSR_PRAGMA_UNROLL _Pragma("unroll")
#define INSR_PRAGMA_UNROLL_N(n) _Pragma(#n)
#endif

#include <stdint.h>
