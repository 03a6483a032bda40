// mb_bf16_subnormal.cu -- does the packed bfloat16 compare (set.lt.u32.bf16x2 -> HSET2.BF16_V2) honour subnormal
// operands, or flush them to zero?  K3's pre-filters compare bf16 truncations of magnitudes against bf16 thresholds;
// if subnormals were flushed, thresholds below 1.18e-38 would need the exact path.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_bf16_subnormal mb_bf16_subnormal.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void k(const uint32_t* a, const uint32_t* b, uint32_t* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint32_t r;
        asm("set.lt.u32.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b[i]));
        out[i] = r;
    }
}

int main() {
    // bf16 bit patterns: 0x0001 = smallest subnormal (9.2e-41), 0x007f = largest subnormal, 0x0080 = smallest normal
    const uint16_t av[] = {0x0000, 0x0001, 0x0001, 0x0002, 0x007f, 0x007f, 0x0080, 0x0000, 0x0040, 0x0041};
    const uint16_t bv[] = {0x0001, 0x0002, 0x0001, 0x0001, 0x0080, 0x007f, 0x007f, 0x0000, 0x0041, 0x0040};
    const int n = sizeof(av) / sizeof(av[0]);
    uint32_t ha[n], hb[n], ho[n], *da, *db, *dout;
    for (int i = 0; i < n; ++i) { ha[i] = av[i] | ((uint32_t)av[i] << 16); hb[i] = bv[i] | ((uint32_t)bv[i] << 16); }
    cudaMalloc(&da, sizeof(ha)); cudaMalloc(&db, sizeof(hb)); cudaMalloc(&dout, sizeof(ho));
    cudaMemcpy(da, ha, sizeof(ha), cudaMemcpyHostToDevice); cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice);
    k<<<1, 32>>>(da, db, dout, n);
    if (cudaMemcpy(ho, dout, sizeof(ho), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        const bool exp = av[i] < bv[i];                      // non-negative bf16: bit order == value order
        const bool got = (ho[i] & 0xffffu) != 0;
        printf("a=0x%04x b=0x%04x  a<b expected %d got %d (mask 0x%08x)%s\n", av[i], bv[i], exp, got, ho[i], exp == got ? "" : "  <-- MISMATCH");
        bad += exp != got;
    }
    printf("%s\n", bad ? "bf16x2 compare does NOT honour subnormals" : "bf16x2 compare honours subnormal operands exactly");
    return 0;
}
