// e2e probe without torch: does intfft_exec_host overlap H2D / kernels / D2H with cudaMallocHost buffers?
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include "../include/intfft.h"
int main()
{
    intfft_generics g{12, 16, 16, 0, 0, 1, 1, 0};
    const long long batch = 65536;
    intfft_plan *p = nullptr;
    if (intfft_plan_create(&p, &g, batch, 0)) return 1;
    intfft_layout l; intfft_query(p, &l);
    void *hi, *ho;
    cudaMallocHost(&hi, l.in_bytes); cudaMallocHost(&ho, l.out_bytes);
    for (int r = 0; r < 4; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        int st = intfft_exec_host(p, hi, ho);
        auto t1 = std::chrono::steady_clock::now();
        std::printf("exec_host (cudaMallocHost): status %d, %.2f ms\n", st, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    return 0;
}
