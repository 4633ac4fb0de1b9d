// Latency of the in-warp "who has my slot index" step: MATCH.ANY vs 12 ballots.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__device__ __forceinline__ unsigned eq_mask12(uint32_t idx)
{
    unsigned m = FULL;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const bool bit = (idx >> k) & 1u;
        const unsigned bal = __ballot_sync(FULL, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
}
template <int W> __global__ void k(uint32_t *out, int iters, int groups)
{
    uint32_t lane = threadIdx.x, x = lane, acc = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        // idx: 12-bit value; `groups` distinct values over the warp, changing every iteration
        const uint32_t idx = ((lane % groups) * 0x9e5u + x * 0x31u) & 4095u;
        unsigned g;
        if (W == 0) g = __match_any_sync(FULL, idx);
        else if (W == 1) g = eq_mask12(idx);
        else g = idx;
        x = (x + (g & 1u) + 1) & 0xffffu; // dependent
        acc += g;
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = (uint32_t)((t1 - t0) / iters); out[1] = acc; }
    if (W == 1) { // check
        const uint32_t idx = ((lane % groups) * 0x9e5u) & 4095u;
        if (__match_any_sync(FULL, idx) != eq_mask12(idx)) out[2] = 1;
    }
}
int main()
{
    uint32_t *out; cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
    for (int groups : {32, 17, 8, 1}) {
        uint32_t r[3][3];
        k<0><<<1, 32>>>(out, 10000, groups); cudaMemcpy(r[0], out, 12, cudaMemcpyDeviceToHost);
        k<1><<<1, 32>>>(out, 10000, groups); cudaMemcpy(r[1], out, 12, cudaMemcpyDeviceToHost);
        k<2><<<1, 32>>>(out, 10000, groups); cudaMemcpy(r[2], out, 12, cudaMemcpyDeviceToHost);
        printf("groups %2d: match_any %u, 12 ballots %u, empty loop %u cycles (mismatch flag %u) %s\n", groups, r[0][0], r[1][0], r[2][0], r[1][2], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
