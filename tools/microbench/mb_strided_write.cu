// mb_strided_write.cu -- HBM write bandwidth of the K1 corner-turn store pattern: SEG-byte segments, one per
// range line, lines 2 KiB apart (256 chirps x 8 B), against a fully contiguous write of the same bytes.
// Each CTA plays one K1 tile: 512 lines x SEG bytes; tiles walk across a line, then on to the next map.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <int SEG>
__global__ void __launch_bounds__(256) strided(float4* __restrict__ out, int n_tiles, int lines, int line_bytes) {
    constexpr int V = SEG / 16;                                    // float4 per segment
    const int tiles_per_map = line_bytes / SEG;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int mp = t / tiles_per_map, tc = t % tiles_per_map;
        char* base = reinterpret_cast<char*>(out) + (size_t)mp * lines * line_bytes + (size_t)tc * SEG;
        for (int i = threadIdx.x; i < lines * V; i += 256) {
            const int k = i / V, v = i % V;
            *reinterpret_cast<float4*>(base + (size_t)k * line_bytes + v * 16) = make_float4(1.f, 2.f, 3.f, (float)i);
        }
    }
}
__global__ void __launch_bounds__(256) contiguous(float4* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

template <int SEG>
float run_strided(float4* buf, size_t bytes, int grid) {
    const int lines = 512, line_bytes = 2048;
    const int n_tiles = (int)(bytes / ((size_t)lines * SEG));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    strided<SEG><<<grid, 256>>>(buf, n_tiles, lines, line_bytes);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) strided<SEG><<<grid, 256>>>(buf, n_tiles, lines, line_bytes);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return (float)(5.0 * bytes / (ms * 1e-3) / 1e9);
}

int main() {
    const size_t bytes = (size_t)4 << 30;                          // 4 GiB >> L2
    float4* buf;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    for (int per_sm = 2; per_sm <= 8; per_sm *= 2) {
        const int grid = p.multiProcessorCount * per_sm;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        contiguous<<<grid, 256>>>(buf, bytes / 16);
        cudaEventRecord(a);
        for (int r = 0; r < 5; ++r) contiguous<<<grid, 256>>>(buf, bytes / 16);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("CTAs/SM %d  contiguous %.0f GB/s | segments of 32 B %.0f  64 B %.0f  128 B %.0f  256 B %.0f  512 B %.0f GB/s\n", per_sm,
               5.0 * bytes / (ms * 1e-3) / 1e9, run_strided<32>(buf, bytes, grid), run_strided<64>(buf, bytes, grid),
               run_strided<128>(buf, bytes, grid), run_strided<256>(buf, bytes, grid), run_strided<512>(buf, bytes, grid));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
