// rdp_common.cuh -- shared helpers for the sm_100a kernels of librdp.
#pragma once

#ifdef RDP_EMU
#include "cuda_emu.h"          // tests/emu/ (the CPU test harness only; -I tests/emu on that build)
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#ifdef RDP_EMU
#define RDP_DYN_SMEM(name) unsigned char* name = emu::ctx()->dyn_smem.data() + ((64 - ((uintptr_t)emu::ctx()->dyn_smem.data() & 63)) & 63)
#define RDP_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=] { kernel(__VA_ARGS__); })
#define RDP_UNROLL
#else
#define RDP_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define RDP_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<dim3(grid), dim3(block), (size_t)(smem), (stream)>>>(__VA_ARGS__)
#define RDP_UNROLL _Pragma("unroll")
#endif

namespace rdp {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -j : (x + jy)(-j) = y - jx
__device__ __forceinline__ float2 cmul_mj(float2 a) { return make_float2(a.y, -a.x); }

// Group barrier: groups of <= 32 threads live inside one warp (warp-synchronous); larger
// groups use a named hardware barrier (ids 1..15, id 0 is __syncthreads).
template <int GROUP_THREADS>
__device__ __forceinline__ void group_sync(int group_id) {
    if (GROUP_THREADS <= 32) {
        __syncwarp();
    } else {
#ifdef RDP_EMU
        emu::named_barrier(1 + group_id, GROUP_THREADS)->wait();
#else
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group_id), "n"(GROUP_THREADS) : "memory");
#endif
    }
}

}  // namespace rdp
