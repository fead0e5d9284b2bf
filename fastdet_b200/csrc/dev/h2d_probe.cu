// dev tool: H2D rate of a cudaMallocHost buffer under the conditions the JPEG path creates
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
int main() {
    const size_t n = 33300000, cap = n + n / 4;
    char *h, *d;
    cudaStream_t cs;
    cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    cudaMallocHost(&h, cap);
    cudaMalloc(&d, cap);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto copy = [&](const char* tag) {
        cudaDeviceSynchronize();
        cudaEventRecord(a, cs);
        cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, cs);
        cudaEventRecord(b, cs);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-40s %.3f ms (%.1f GB/s)\n", tag, ms, n / ms / 1e6);
    };
    copy("untouched"); copy("untouched again");
    memset(h, 1, n); copy("after memset by main thread");
    for (int rep = 0; rep < 2; ++rep) {
        std::vector<std::thread> th;
        for (int t = 0; t < 16; ++t) th.emplace_back([&, t] { memset(h + n / 16 * t, t + rep, n / 16); });
        for (auto& x : th) x.join();
        copy("after memset by 16 threads");
    }
    for (int rep = 0; rep < 2; ++rep) {  // sparse int16 writes like the entropy decoder's
        std::vector<std::thread> th;
        for (int t = 0; t < 16; ++t) th.emplace_back([&, t] {
            short* p = reinterpret_cast<short*>(h + n / 16 * t);
            memset(p, 0, n / 16);
            for (size_t i = 0; i < n / 32; i += 7) p[i] = short(i + rep);
        });
        for (auto& x : th) x.join();
        copy("after zero + sparse int16 writes");
    }
    return 0;
}
