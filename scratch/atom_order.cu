// Does a shared-memory atomicAdd hand out old values in lane order when several lanes of a warp hit the same address?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(unsigned long long *viol, unsigned long long *trials, int nbins, int iters, uint32_t seed)
{
    extern __shared__ uint32_t cnt[];      // [warps][nbins]
    const uint32_t w = threadIdx.x >> 5, l = threadIdx.x & 31;
    uint32_t *c = cnt + w * nbins;
    uint32_t s = seed ^ (blockIdx.x * 7919u + threadIdx.x * 104729u);
    unsigned long long v = 0, t = 0;
    for (int it = 0; it < iters; it++) {
        for (int i = l; i < nbins; i += 32) c[i] = 0;
        __syncwarp();
        uint32_t expect_base[8];
        for (int r = 0; r < 8; r++) {
            s = s * 1664525u + 1013904223u;
            uint32_t mode = (it >> 3) & 3;
            uint32_t d = mode == 0 ? (s >> 8) % nbins : mode == 1 ? (s >> 8) % 4 : mode == 2 ? ((s >> 8) % 2 ? 5 : (s >> 12) % nbins) : (l / 4 + (s >> 30));
            d %= nbins;
            unsigned peers = __match_any_sync(0xffffffffu, d);
            uint32_t before = c[d];
            __syncwarp();
            uint32_t old = atomicAdd(&c[d], 1u);
            __syncwarp();
            uint32_t want = before + __popc(peers & ((1u << l) - 1));
            if (old != want) v++;
            t++;
        }
    }
    for (int d = 16; d; d >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, d); t += __shfl_xor_sync(0xffffffffu, t, d); }
    if (l == 0) { atomicAdd(viol, v); atomicAdd(trials, t); }
}
int main()
{
    unsigned long long *d, h[2] = {0, 0};
    cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
    for (int nb : {512, 256, 64}) {
        k<<<148 * 8, 256, 8 * nb * 4>>>(d, d + 1, nb, 4000, 12345u + nb);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("bins %d: %llu violations in %llu trials (%s)\n", nb, h[0], h[1], cudaGetErrorString(e));
    }
    return 0;
}
