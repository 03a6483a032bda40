// cuda_emu.h -- TEST-ONLY host emulation of the small CUDA subset the kernels use.
//
// The authoring container has no GPU.  Compiling the very same kernel sources with g++
// against this shim (-DRDP_EMU) gives `librdp_emu.so`, which the CPU test-suite drives to
// check kernel LOGIC (indexing, FFT passes, CFAR counting, compaction) before GPU time is
// spent.  One OS thread per CUDA thread, real barriers, blocks run one after another.
// It is never built into librdp.so, never loaded by the product package (which refuses a
// library exporting `rdp_emu_marker`), and is not a CPU fallback.
#pragma once
#ifndef RDP_EMU
#error "cuda_emu.h is for the -DRDP_EMU test build only"
#endif

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct short2 { short x, y; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __constant__ static

namespace emu {

class Barrier {
  public:
    explicit Barrier(int n) : n_(n), count_(0), gen_(0) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        int g = gen_;
        if (++count_ == n_) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return g != gen_; });
        }
    }
  private:
    int n_, count_, gen_;
    std::mutex m_;
    std::condition_variable cv_;
};

struct BlockCtx {
    int nthreads = 0;
    std::unique_ptr<Barrier> block_bar;
    std::vector<std::unique_ptr<Barrier>> warp_bar;
    std::vector<std::array<uint64_t, 32>> warp_buf;
    std::mutex named_m;
    std::map<std::pair<int, int>, std::unique_ptr<Barrier>> named;
    std::vector<unsigned char> dyn_smem;
};

inline BlockCtx*& ctx() { static BlockCtx* c = nullptr; return c; }
struct Tls { dim3 threadIdx, blockIdx, blockDim, gridDim; int lin = 0; };
inline Tls& tls() { static thread_local Tls t; return t; }

inline Barrier* named_barrier(int id, int count) {
    BlockCtx* c = ctx();
    std::lock_guard<std::mutex> lk(c->named_m);
    auto key = std::make_pair(id, count);
    auto it = c->named.find(key);
    if (it == c->named.end()) it = c->named.emplace(key, std::make_unique<Barrier>(count)).first;
    return it->second.get();
}

template <typename F>
void launch(dim3 grid, dim3 block, size_t smem, F body) {
    BlockCtx c;
    int nt = (int)(block.x * block.y * block.z);
    c.nthreads = nt;
    c.block_bar = std::make_unique<Barrier>(nt);
    int nwarps = (nt + 31) / 32;
    for (int w = 0; w < nwarps; ++w)
        c.warp_bar.emplace_back(std::make_unique<Barrier>(std::min(32, nt - 32 * w)));
    c.warp_buf.resize(nwarps);
    c.dyn_smem.assign(smem + 64, 0);
    ctx() = &c;
    size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t) {
        th.emplace_back([&, t] {
            Tls& s = tls();
            s.blockDim = block;
            s.gridDim = grid;
            s.lin = t;
            s.threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            for (size_t b = 0; b < nblocks; ++b) {
                s.blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y),
                                  (unsigned)(b / ((size_t)grid.x * grid.y)));
                body();
                c.block_bar->wait();
            }
        });
    }
    for (auto& t : th) t.join();
    ctx() = nullptr;
}

inline void syncwarp() { ctx()->warp_bar[tls().lin / 32]->wait(); }

template <typename T>
inline T warp_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "");
    BlockCtx* c = ctx();
    int w = tls().lin / 32, lane = tls().lin % 32;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    c->warp_buf[w][lane] = raw;
    syncwarp();
    int nl = std::min(32, c->nthreads - 32 * w);
    uint64_t got = (src_lane >= 0 && src_lane < nl) ? c->warp_buf[w][src_lane] : raw;
    syncwarp();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
}  // namespace emu

#define threadIdx (emu::tls().threadIdx)
#define blockIdx (emu::tls().blockIdx)
#define blockDim (emu::tls().blockDim)
#define gridDim (emu::tls().gridDim)

static inline void __syncthreads() { emu::ctx()->block_bar->wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu::warp_exchange(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::warp_exchange(v, (emu::tls().lin % 32) ^ m); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu::warp_exchange(v, (emu::tls().lin % 32) + d); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu::warp_exchange(v, (emu::tls().lin % 32) - d); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) {
        int p = emu::warp_exchange(pred ? 1 : 0, l);
        int nl = std::min(32, emu::ctx()->nthreads - 32 * (emu::tls().lin / 32));
        if (l < nl && p) r |= 1u << l;
    }
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __reduce_add_sync(unsigned m, int v) {
    for (int o = 16; o > 0; o >>= 1) {
        int nl = std::min(32, emu::ctx()->nthreads - 32 * (emu::tls().lin / 32));
        int lane = emu::tls().lin % 32;
        int other = emu::warp_exchange(v, lane ^ o);
        if ((lane ^ o) < nl) v += other;
    }
    (void)m;
    return v;
}
static inline int __reduce_max_sync(unsigned m, int v) {
    for (int o = 16; o > 0; o >>= 1) {
        int nl = std::min(32, emu::ctx()->nthreads - 32 * (emu::tls().lin / 32));
        int lane = emu::tls().lin % 32;
        int other = emu::warp_exchange(v, lane ^ o);
        if ((lane ^ o) < nl) v = std::max(v, other);
    }
    (void)m;
    return v;
}
static inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fmul_rz(float a, float b) {          // round toward zero (operands >= 0 here)
    double p = (double)a * (double)b; float r = (float)p;
    if ((double)r > p) r = std::nextafterf(r, 0.0f);
    return r;
}
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }

// ---- runtime API subset ------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
struct EmuEvent { std::chrono::steady_clock::time_point t; };
typedef EmuEvent* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDefault = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; size_t sharedMemPerBlockOptin; int major, minor; int l2CacheSize; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p));
    std::snprintf(p->name, sizeof(p->name), "rdp host emulator (not a GPU)");
    p->multiProcessorCount = 2; p->sharedMemPerBlockOptin = 227 * 1024; p->major = 10; p->minor = 0;
    p->l2CacheSize = 1 << 20;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, int, cudaStream_t) {
    for (size_t i = 0; i < h; ++i) std::memcpy((char*)d + i * dp, (const char*)s + i * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess;
}
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
