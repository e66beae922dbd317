// microbenchmark: cost of intra-warp digit matching on B200
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_match(const uint32_t *in, uint32_t *out, int ndist)
{
    uint32_t d = in[threadIdx.x + blockIdx.x * blockDim.x] % ndist, acc = 0;
    for (int i = 0; i < ITERS; i++) {
        unsigned peers = __match_any_sync(0xffffffffu, d);
        acc += __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
        d = (d * 1664525u + 1013904223u + acc) % ndist;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_ballot10(const uint32_t *in, uint32_t *out, int ndist)
{
    uint32_t d = in[threadIdx.x + blockIdx.x * blockDim.x] % ndist, acc = 0;
    for (int i = 0; i < ITERS; i++) {
        unsigned peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 10; b++) {
            bool bit = (d >> b) & 1;
            unsigned m = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? m : ~m;
        }
        acc += __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
        d = (d * 1664525u + 1013904223u + acc) % ndist;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_base(const uint32_t *in, uint32_t *out, int ndist)
{
    uint32_t d = in[threadIdx.x + blockIdx.x * blockDim.x] % ndist, acc = 0;
    for (int i = 0; i < ITERS; i++) {
        acc += __popc(d & ((1u << (threadIdx.x & 31)) - 1));
        d = (d * 1664525u + 1013904223u + acc) % ndist;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_atoms(const uint32_t *in, uint32_t *out, int ndist)
{
    __shared__ uint32_t cnt[8][1024];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) (&cnt[0][0])[i] = 0;
    __syncthreads();
    uint32_t d = in[threadIdx.x + blockIdx.x * blockDim.x] % ndist, acc = 0;
    uint32_t *c = cnt[threadIdx.x >> 5];
    for (int i = 0; i < ITERS; i++) {
        acc += atomicAdd(&c[d & 1023], 1u);
        d = (d * 1664525u + 1013904223u + acc) % ndist;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_rank_match(const uint32_t *in, uint32_t *out, int ndist)
{
    __shared__ uint16_t cnt[8][1024];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) (&cnt[0][0])[i] = 0;
    __syncthreads();
    uint32_t d = in[threadIdx.x + blockIdx.x * blockDim.x] % ndist, acc = 0, l = threadIdx.x & 31;
    uint16_t *mycnt = cnt[threadIdx.x >> 5];
    for (int i = 0; i < ITERS; i++) {
        unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned lt = peers & ((1u << l) - 1);
        uint16_t b = mycnt[d];
        __syncwarp();
        if (lt == 0) mycnt[d] = (uint16_t)(b + __popc(peers));
        __syncwarp();
        acc += b + __popc(lt);
        d = (d * 1664525u + 1013904223u + acc) % ndist;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_rank_or(const uint32_t *in, uint32_t *out, int ndist)
{
    __shared__ uint16_t cnt[8][1024];
    __shared__ uint32_t M[8][1024];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) { (&cnt[0][0])[i] = 0; (&M[0][0])[i] = 0; }
    __syncthreads();
    uint32_t d = in[threadIdx.x + blockIdx.x * blockDim.x] % ndist, acc = 0, l = threadIdx.x & 31;
    uint16_t *mycnt = cnt[threadIdx.x >> 5];
    uint32_t *m = M[threadIdx.x >> 5];
    for (int i = 0; i < ITERS; i++) {
        atomicOr(&m[d], 1u << l);
        __syncwarp();
        unsigned peers = m[d];
        uint16_t b = mycnt[d];
        __syncwarp();
        unsigned lt = peers & ((1u << l) - 1);
        if (lt == 0) { mycnt[d] = (uint16_t)(b + __popc(peers)); m[d] = 0; }
        __syncwarp();
        acc += b + __popc(lt);
        d = (d * 1664525u + 1013904223u + acc) % ndist;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
int main()
{
    const int CTAS = 148 * 4, TH = 256;
    uint32_t *in, *out;
    cudaMalloc(&in, CTAS * TH * 4); cudaMalloc(&out, CTAS * TH * 4);
    uint32_t *h = new uint32_t[CTAS * TH];
    for (int i = 0; i < CTAS * TH; i++) h[i] = (uint32_t)rand();
    cudaMemcpy(in, h, CTAS * TH * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int nds[] = {1, 4, 12, 32, 400, 1024};
    for (int nd : nds) {
        float ms[6];
        for (int which = 0; which < 6; which++) {
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (which == 0) k_base<<<CTAS, TH>>>(in, out, nd);
                if (which == 1) k_match<<<CTAS, TH>>>(in, out, nd);
                if (which == 2) k_ballot10<<<CTAS, TH>>>(in, out, nd);
                if (which == 3) k_atoms<<<CTAS, TH>>>(in, out, nd);
                if (which == 4) k_rank_match<<<CTAS, TH>>>(in, out, nd);
                if (which == 5) k_rank_or<<<CTAS, TH>>>(in, out, nd);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms[which], e0, e1);
            }
        }
        double warp_items = (double)CTAS * TH / 32 * ITERS;
        // cycles per warp-item per SM sub-partition (4 per SM), 1.965 GHz
        auto cyc = [&](float m) { return m * 1e-3 * 1.965e9 * 148 * 4 / warp_items; };
        printf("ndist %4d: base %.3f ms (%.1f cyc/warp-item/SMSP)  match %.3f (%.1f)  ballot10 %.3f (%.1f)  atoms %.3f (%.1f) rank_match %.3f (%.1f) rank_or %.3f (%.1f)\n", nd,
               ms[0], cyc(ms[0]), ms[1], cyc(ms[1]), ms[2], cyc(ms[2]), ms[3], cyc(ms[3]), ms[4], cyc(ms[4]), ms[5], cyc(ms[5]));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
