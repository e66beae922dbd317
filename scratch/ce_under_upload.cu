// What does a small copy-engine operation (cudaMemsetAsync, 8-byte device-to-device / device-to-host cudaMemcpyAsync) cost on
// one stream while a large host-to-device copy runs on another, against a kernel that does the same?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/ce_under_upload scratch/ce_under_upload.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>
__global__ void k_fill(unsigned long long *p, unsigned long long v, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
__global__ void k_copy(const unsigned long long *s, unsigned long long *d, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) d[i] = s[i]; }
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <class F> static double med(F f, int n = 120)
{
    std::vector<double> t;
    for (int i = 0; i < n; i++) { double a = now_us(); f(); t.push_back(now_us() - a); }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}
int main()
{
    const size_t N = 3ull << 30;
    unsigned char *h, *d;
    cudaMallocHost(&h, N); cudaMalloc(&d, N);
    unsigned long long *w, *hp;
    cudaMalloc(&w, 1 << 20); cudaMallocHost(&hp, 4096);
    cudaStream_t up, s;
    cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int pass = 0; pass < 2; pass++) {
        if (pass) cudaMemcpyAsync(d, h, N, cudaMemcpyHostToDevice, up);
        double a = med([&] { k_fill<<<1, 64, 0, s>>>(w, 1, 64); cudaStreamSynchronize(s); });
        double b = med([&] { cudaMemsetAsync(w, 0, 64, s); cudaStreamSynchronize(s); });
        double b2 = med([&] { cudaMemsetAsync(w, 0, 1 << 20, s); cudaStreamSynchronize(s); });
        double c = med([&] { cudaMemcpyAsync(w + 64, w, 8, cudaMemcpyDeviceToDevice, s); cudaStreamSynchronize(s); });
        double e = med([&] { cudaMemcpyAsync(hp, w, 8, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); });
        double f = med([&] { k_copy<<<1, 64, 0, s>>>(w, hp, 8); cudaStreamSynchronize(s); });
        double g = med([&] { k_fill<<<1, 64, 0, s>>>(w, 1, 64); cudaMemsetAsync(w + 128, 0, 64, s); k_fill<<<1, 64, 0, s>>>(w, 1, 64); cudaMemsetAsync(w + 128, 0, 64, s);
                             k_fill<<<1, 64, 0, s>>>(w, 1, 64); cudaMemcpyAsync(hp, w, 8, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); });
        double g2 = med([&] { k_fill<<<1, 64, 0, s>>>(w, 1, 64); k_fill<<<1, 64, 0, s>>>(w + 128, 0, 8); k_fill<<<1, 64, 0, s>>>(w, 1, 64); k_fill<<<1, 64, 0, s>>>(w + 128, 0, 8);
                              k_fill<<<1, 64, 0, s>>>(w, 1, 64); k_copy<<<1, 64, 0, s>>>(w, hp, 8); cudaStreamSynchronize(s); });
        bool running = pass && cudaStreamQuery(up) == cudaErrorNotReady;
        printf("%s: kernel+sync %.1f us | memset 64 B %.1f | memset 1 MiB %.1f | D2D 8 B %.1f | D2H 8 B %.1f | kernel store to host %.1f | 3 kernels + 2 memsets + D2H %.1f | the same as 6 kernels %.1f%s\n",
               pass ? "beside a 3 GiB upload" : "idle", a, b, b2, c, e, f, g, g2, pass ? (running ? " (upload still running)" : " (UPLOAD ENDED EARLY)") : "");
        cudaDeviceSynchronize();
    }
    return 0;
}
