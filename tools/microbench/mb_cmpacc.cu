// mb_cmpacc.cu -- throughput of the compare-and-count idioms considered for the OS-CFAR
// phase-1 counter (K3).  Prints compares/s for each variant on the current GPU.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_cmpacc mb_cmpacc.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2000
#define NR 16   // refs held per thread
#define NC 8    // counters per thread

__global__ void __launch_bounds__(256) k_setp(const float* __restrict__ t, int* out, float c) {
    float v[NR], cut[NC]; int cnt[NC];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC; ++j) { cut[j] = c + j; cnt[j] = 0; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; ++i)
            #pragma unroll
            for (int j = 0; j < NC; ++j) cnt[j] += (v[i] < cut[j]) ? 1 : 0;
        #pragma unroll
        for (int j = 0; j < NC; ++j) cut[j] += 0.37f;
    }
    int s = 0; for (int j = 0; j < NC; ++j) s += cnt[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_sign(const float* __restrict__ t, int* out, float c) {
    float v[NR], cut[NC]; unsigned cnt[NC];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC; ++j) { cut[j] = c + j; cnt[j] = 0; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; ++i)
            #pragma unroll
            for (int j = 0; j < NC; ++j) cnt[j] += __float_as_uint(v[i] - cut[j]) >> 31;
        #pragma unroll
        for (int j = 0; j < NC; ++j) cut[j] += 0.37f;
    }
    int s = 0; for (int j = 0; j < NC; ++j) s += cnt[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_sat(const float* __restrict__ t, int* out, float c) {
    float v[NR], tb[NC]; unsigned cnt[NC];
    const float BIG = 1.8446744e19f;
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC; ++j) { tb[j] = (c + j) * BIG; cnt[j] = 0; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; i += 2)
            #pragma unroll
            for (int j = 0; j < NC; ++j) {
                float r0, r1;
                asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r0) : "f"(v[i]), "f"(-BIG), "f"(tb[j]));
                asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r1) : "f"(v[i + 1]), "f"(-BIG), "f"(tb[j]));
                cnt[j] += __float_as_uint(r0) + __float_as_uint(r1);
            }
        #pragma unroll
        for (int j = 0; j < NC; ++j) tb[j] += 0.37f * BIG;
    }
    int s = 0; for (int j = 0; j < NC; ++j) s += cnt[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
// bf16x2: two refs per register (or one ref duplicated) against a pair of cuts; mask result
__global__ void __launch_bounds__(256) k_bf16_mask(const uint32_t* __restrict__ t, int* out, uint32_t c) {
    uint32_t v[NR], cut[NC / 2]; unsigned cnt[NC / 2];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC / 2; ++j) { cut[j] = c + 0x00010001u * j; cnt[j] = 0; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; i += 2)
            #pragma unroll
            for (int j = 0; j < NC / 2; ++j) {
                uint32_t m0, m1;
                asm("set.lt.u32.bf16x2 %0, %1, %2;" : "=r"(m0) : "r"(v[i]), "r"(cut[j]));
                asm("set.lt.u32.bf16x2 %0, %1, %2;" : "=r"(m1) : "r"(v[i + 1]), "r"(cut[j]));
                cnt[j] -= m0 + m1;
            }
        #pragma unroll
        for (int j = 0; j < NC / 2; ++j) cut[j] += 0x00010001u;
    }
    int s = 0; for (int j = 0; j < NC / 2; ++j) s += cnt[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_bf16_add(const uint32_t* __restrict__ t, int* out, uint32_t c) {
    uint32_t v[NR], cut[NC / 2]; __nv_bfloat162 cnt[NC / 2];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC / 2; ++j) { cut[j] = c + 0x00010001u * j; cnt[j] = __float2bfloat162_rn(0.f); }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; ++i)
            #pragma unroll
            for (int j = 0; j < NC / 2; ++j) {
                uint32_t m0;
                asm("set.lt.bf16x2.bf16x2 %0, %1, %2;" : "=r"(m0) : "r"(v[i]), "r"(cut[j]));
                cnt[j] = __hadd2(cnt[j], *reinterpret_cast<__nv_bfloat162*>(&m0));
            }
        #pragma unroll
        for (int j = 0; j < NC / 2; ++j) cut[j] += 0x00010001u;
    }
    int s = 0; for (int j = 0; j < NC / 2; ++j) s += *reinterpret_cast<int*>(&cnt[j]);
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
// fp16x2 variant of the mask form
__global__ void __launch_bounds__(256) k_f16_mask(const uint32_t* __restrict__ t, int* out, uint32_t c) {
    uint32_t v[NR], cut[NC / 2]; unsigned cnt[NC / 2];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC / 2; ++j) { cut[j] = c + 0x00010001u * j; cnt[j] = 0; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; i += 2)
            #pragma unroll
            for (int j = 0; j < NC / 2; ++j) {
                uint32_t m0, m1;
                asm("set.lt.u32.f16x2 %0, %1, %2;" : "=r"(m0) : "r"(v[i]), "r"(cut[j]));
                asm("set.lt.u32.f16x2 %0, %1, %2;" : "=r"(m1) : "r"(v[i + 1]), "r"(cut[j]));
                cnt[j] -= m0 + m1;
            }
        #pragma unroll
        for (int j = 0; j < NC / 2; ++j) cut[j] += 0x00010001u;
    }
    int s = 0; for (int j = 0; j < NC / 2; ++j) s += cnt[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

// DPX: max(min(cut + (-ref), 1), 0) per signed 16-bit half = [ref < cut] on bf16 bit patterns (order-preserving for
// non-negative values); the 0/1 halves add up directly
__global__ void __launch_bounds__(256) k_dpx(const uint32_t* __restrict__ t, int* out, uint32_t c) {
    uint32_t v[NR], cut[NC / 2]; unsigned cnt[NC / 2];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC / 2; ++j) { cut[j] = c + 0x00010001u * j; cnt[j] = 0; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; i += 2)
            #pragma unroll
            for (int j = 0; j < NC / 2; ++j)
                cnt[j] += __viaddmin_s16x2_relu(cut[j], v[i], 0x00010001u) + __viaddmin_s16x2_relu(cut[j], v[i + 1], 0x00010001u);
        #pragma unroll
        for (int j = 0; j < NC / 2; ++j) cut[j] += 0x00010001u;
    }
    int s = 0; for (int j = 0; j < NC / 2; ++j) s += cnt[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
// SIMD-in-a-word compare intrinsic
__global__ void __launch_bounds__(256) k_vcmp(const uint32_t* __restrict__ t, int* out, uint32_t c) {
    uint32_t v[NR], cut[NC / 2]; unsigned cnt[NC / 2];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC / 2; ++j) { cut[j] = c + 0x00010001u * j; cnt[j] = 0; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; i += 2)
            #pragma unroll
            for (int j = 0; j < NC / 2; ++j) cnt[j] -= __vcmpltu2(v[i], cut[j]) + __vcmpltu2(v[i + 1], cut[j]);
        #pragma unroll
        for (int j = 0; j < NC / 2; ++j) cut[j] += 0x00010001u;
    }
    int s = 0; for (int j = 0; j < NC / 2; ++j) s += cnt[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
// saturating SIMD subtract: min(max(cut - ref, 0), ...) needs a second op; measure vsubus2 + vminu2
__global__ void __launch_bounds__(256) k_vsub(const uint32_t* __restrict__ t, int* out, uint32_t c) {
    uint32_t v[NR], cut[NC / 2]; unsigned cnt[NC / 2];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC / 2; ++j) { cut[j] = c + 0x00010001u * j; cnt[j] = 0; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; ++i)
            #pragma unroll
            for (int j = 0; j < NC / 2; ++j) cnt[j] += __vminu2(__vsubus2(cut[j], v[i]), 0x00010001u);
        #pragma unroll
        for (int j = 0; j < NC / 2; ++j) cut[j] += 0x00010001u;
    }
    int s = 0; for (int j = 0; j < NC / 2; ++j) s += cnt[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

// HSET2 masks accumulated on the FMA pipe (IMAD acc = m * -1 + acc) instead of IADD3 on the ALU pipe
__global__ void __launch_bounds__(256) k_bf16_imad(const uint32_t* __restrict__ t, int* out, uint32_t c) {
    uint32_t v[NR], cut[NC / 2]; int cnt[NC / 2];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC / 2; ++j) { cut[j] = c + 0x00010001u * j; cnt[j] = 0; }
    int minus1 = -(int)(c >> 31) - 1 + (int)(threadIdx.x >> 10);      // -1, opaque to the compiler
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; ++i)
            #pragma unroll
            for (int j = 0; j < NC / 2; ++j) {
                uint32_t m0;
                asm("set.lt.u32.bf16x2 %0, %1, %2;" : "=r"(m0) : "r"(v[i]), "r"(cut[j]));
                asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(cnt[j]) : "r"(m0), "r"(minus1));
            }
        #pragma unroll
        for (int j = 0; j < NC / 2; ++j) cut[j] += 0x00010001u;
    }
    int s = 0; for (int j = 0; j < NC / 2; ++j) s += cnt[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

// half of the references through HSET2 (fp16-class pipe), half through VIADDMNMX (integer pipe)
__global__ void __launch_bounds__(256) k_mix(const uint32_t* __restrict__ t, int* out, uint32_t c) {
    uint32_t v[NR], cut[NC / 2], ncut[NC / 2]; unsigned cnt[NC / 2], cge[NC / 2];
    for (int i = 0; i < NR; ++i) v[i] = t[i * 256 + threadIdx.x];
    for (int j = 0; j < NC / 2; ++j) { cut[j] = c + 0x00010001u * j; ncut[j] = 0x00010001u - cut[j]; cnt[j] = 0; cge[j] = 0; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NR; i += 4)
            #pragma unroll
            for (int j = 0; j < NC / 2; ++j) {
                uint32_t m0, m1;
                asm("set.lt.u32.bf16x2 %0, %1, %2;" : "=r"(m0) : "r"(v[i]), "r"(cut[j]));
                asm("set.lt.u32.bf16x2 %0, %1, %2;" : "=r"(m1) : "r"(v[i + 1]), "r"(cut[j]));
                cnt[j] -= m0 + m1;
                cge[j] += __viaddmin_s16x2_relu(v[i + 2], ncut[j], 0x00010001u) + __viaddmin_s16x2_relu(v[i + 3], ncut[j], 0x00010001u);
            }
        #pragma unroll
        for (int j = 0; j < NC / 2; ++j) { cut[j] += 0x00010001u; ncut[j] -= 0x00010001u; }
    }
    int s = 0; for (int j = 0; j < NC / 2; ++j) s += cnt[j] + cge[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <class K, class T, class A>
void run(const char* name, K k, const T* d_t, int* d_out, A arg, double compares_per_thread) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * 8;
    k<<<grid, 256>>>(d_t, d_out, arg);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); k<<<grid, 256>>>(d_t, d_out, arg); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    double total = compares_per_thread * 256.0 * grid;
    cudaError_t e = cudaGetLastError();
    printf("%-12s %8.3f ms  %8.2f Gcompare/s  (%.3f clk/warp-compare/SMSP at 1.9 GHz)  %s\n", name, best,
           total / best * 1e-6, 1.9e9 * 148 * 4 / (total / 32.0 / (best * 1e-3)), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    float* d_t; int* d_out;
    cudaMalloc(&d_t, NR * 256 * sizeof(float)); cudaMalloc(&d_out, 148 * 8 * 256 * sizeof(int));
    float h[NR * 256]; for (int i = 0; i < NR * 256; ++i) h[i] = (float)((i * 2654435761u) % 1000) * 0.01f + 1.0f;
    cudaMemcpy(d_t, h, sizeof(h), cudaMemcpyHostToDevice);
    double scalar = (double)ITERS * NR * NC, packed = (double)ITERS * NR * NC;   // packed: NR x NC/2 instr x 2 halves
    run("setp+add", k_setp, d_t, d_out, 3.0f, scalar);
    run("fadd+lea.hi", k_sign, d_t, d_out, 3.0f, scalar);
    run("ffma.sat", k_sat, d_t, d_out, 3.0f, scalar);
    run("bf16x2 mask", k_bf16_mask, (const uint32_t*)d_t, d_out, 0x40404040u, packed);
    run("bf16x2 add", k_bf16_add, (const uint32_t*)d_t, d_out, 0x40404040u, packed);
    run("f16x2 mask", k_f16_mask, (const uint32_t*)d_t, d_out, 0x40404040u, packed);
    run("dpx viaddmin", k_dpx, (const uint32_t*)d_t, d_out, 0x40404040u, packed);
    run("bf16x2+imad", k_bf16_imad, (const uint32_t*)d_t, d_out, 0x40404040u, packed);
    run("hset2+dpx mix", k_mix, (const uint32_t*)d_t, d_out, 0x40404040u, packed);
    run("vcmpltu2", k_vcmp, (const uint32_t*)d_t, d_out, 0x40404040u, packed);
    run("vsubus2+vminu2", k_vsub, (const uint32_t*)d_t, d_out, 0x40404040u, packed);
    return 0;
}
