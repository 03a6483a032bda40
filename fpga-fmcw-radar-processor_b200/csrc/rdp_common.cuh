// rdp_common.cuh -- shared helpers for the sm_100a kernels of librdp.
#pragma once

#ifdef RDP_EMU
#include "cuda_emu.h"          // tests/emu/ (the CPU test harness only; -I tests/emu on that build)
#else
#include <cuda_runtime.h>
#include <cuda.h>          // CUtensorMap (type only; nothing from libcuda is linked)
#endif
#include <stdint.h>

#ifdef RDP_EMU
#define RDP_DYN_SMEM(name) unsigned char* name = emu::ctx()->dyn_smem.data() + ((64 - ((uintptr_t)emu::ctx()->dyn_smem.data() & 63)) & 63)
#define RDP_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=] { kernel(__VA_ARGS__); })
#define RDP_UNROLL
#define RDP_ROLL_P0
#define RDP_ROLL_P1
#else
#define RDP_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define RDP_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<dim3(grid), dim3(block), (size_t)(smem), (stream)>>>(__VA_ARGS__)
#define RDP_UNROLL _Pragma("unroll")
// K3's column loops: rolled (`unroll 1`, -DRDP_K3_ROLL_P0=1 / _P1=1) or fully unrolled (default).  Measured on B200
// (profiles/r2_summary.md): rolling shrinks the kernel from 5.2 k to 1.8 k SASS instructions but is slower (K3 9.56 ->
// 10.27 ms with both rolled, 9.56 with phase 0 only): the instruction-fetch stalls it removes cost less than the
// scheduling freedom and the column-kind branch it adds.
#ifndef RDP_K3_ROLL_P0
#define RDP_K3_ROLL_P0 0
#endif
#ifndef RDP_K3_ROLL_P1
#define RDP_K3_ROLL_P1 0
#endif
#if RDP_K3_ROLL_P0
#define RDP_ROLL_P0 _Pragma("unroll 1")
#else
#define RDP_ROLL_P0 _Pragma("unroll")
#endif
#if RDP_K3_ROLL_P1
#define RDP_ROLL_P1 _Pragma("unroll 1")
#else
#define RDP_ROLL_P1 _Pragma("unroll")
#endif
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

// ---- TMA (cp.async.bulk.tensor) + mbarrier primitives, shared by K1 (tile store) and K3 (tile load) ----------------
#ifdef RDP_EMU
struct CUtensorMap_st { alignas(64) unsigned char opaque[128]; };
typedef CUtensorMap_st CUtensorMap;
#define RDP_GRID_CONSTANT
#else
#define RDP_GRID_CONSTANT __grid_constant__
// mbarrier + TMA (cp.async.bulk.tensor) primitives, sm_90+ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 3-D tiled TMA load global -> shared; out-of-bounds elements are written as 0 (== cfar_edges ZERO)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tmap, unsigned long long* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
// the same box pulled into L2 only (no shared-memory destination): issued a tile ahead of the load that needs it
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tmap, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(x), "r"(y), "r"(z) : "memory");
}
// 3-D tiled TMA store shared -> global (bulk async-group completion); the shared-memory box is in the tensor map's
// swizzled layout.  Writers make their generic-proxy stores visible with tma_store_fence() before the barrier that
// precedes the issue; the issuing thread commits a group and waits for its READS before the box is reused.
__device__ __forceinline__ void tma_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, const void* src, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
#endif


}  // namespace rdp
