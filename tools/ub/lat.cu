// Latency microbenchmarks used to budget the parse step (one warp, dependent chains).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__global__ void k(int which, uint32_t *out, const uint32_t *chase, int iters, uint32_t mask = 4095)
{
    uint32_t lane = threadIdx.x, x = lane * 0x9e3779b1u + 7, acc = 0;
    __shared__ uint32_t sm[4096];
    for (int i = lane; i < 4096; i += 32) sm[i] = (i * 7 + 3) & 4095;
    __syncwarp();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        switch (which) {
        case 0: x = __match_any_sync(FULL, x ^ lane * 0x10001u) + x * 3; break;          // all distinct (mostly)
        case 1: x = __match_any_sync(FULL, (x & 0) + (lane >> 1)) ^ (x >> 31); break;    // 16 pairs
        case 2: x = __match_any_sync(FULL, x & 0) ^ (x >> 31); break;                    // all same
        case 3: x = __shfl_sync(FULL, x, (x + 1) & 31) + 1; break;
        case 4: x = __ballot_sync(FULL, x & 1) + lane; break;
        case 5: x = sm[x & 4095]; break;
        case 6: x = chase[x & mask]; break;                                                        // global chase
        case 7: x = x * 0x1e35a7bdu + 1; break;
        case 8: x = __reduce_or_sync(FULL, x) + lane; break;
        case 9: x = __match_any_sync(FULL, (x & 0) + (lane & 7)) ^ (x >> 31); break;     // 8 groups of 4
        }
        acc += x;
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = (uint32_t)((t1 - t0) / iters); out[1] = acc; }
}
int main()
{
    uint32_t *out, *chase; 
    cudaMalloc(&out, 8);
    const char *names[] = {"match_any distinct", "match_any 16 pairs", "match_any all same", "shfl", "ballot", "lds", "ldg chase", "imad", "redux.or", "match_any 8 groups"};
    for (int sz : {1 << 12, 1 << 22, 1 << 27}) { // words: 16 KiB (L1), 16 MiB (L2), 512 MiB (DRAM)
        uint32_t *h = (uint32_t *)malloc((size_t)sz * 4);
        // random cyclic permutation with a big stride
        uint64_t stride = (uint64_t)sz / 2 + 12345 | 1;
        for (uint64_t i = 0; i < (uint64_t)sz; ++i) h[i] = (uint32_t)((i * 1 + stride * 97) % sz);
        for (uint64_t i = 0; i < (uint64_t)sz; ++i) h[i] = (uint32_t)((i + stride) % sz);
        cudaMalloc(&chase, (size_t)sz * 4);
        cudaMemcpy(chase, h, (size_t)sz * 4, cudaMemcpyHostToDevice);
        k<<<1, 32>>>(6, out, chase, 2000, sz - 1);
        k<<<1, 32>>>(6, out, chase, 20000, sz - 1);
        uint32_t r[2]; cudaError_t e = cudaMemcpy(r, out, 8, cudaMemcpyDeviceToHost); if (e) printf("err %s\n", cudaGetErrorString(e));
        printf("ldg chase over %d KiB: %u cycles\n", sz / 256, r[0]);
        cudaFree(chase); free(h);
    }
    cudaMalloc(&chase, 4096 * 4);
    cudaMemset(chase, 0, 4096 * 4);
    for (int w = 0; w < 10; ++w) {
        if (w == 6) continue;
        k<<<1, 32>>>(w, out, chase, 1000);
        k<<<1, 32>>>(w, out, chase, 10000);
        uint32_t r[2]; cudaError_t e = cudaMemcpy(r, out, 8, cudaMemcpyDeviceToHost); if (e) printf("err %s\n", cudaGetErrorString(e));
        printf("%-20s %u cycles\n", names[w], r[0]);
    }
    return 0;
}
