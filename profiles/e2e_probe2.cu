#include <chrono>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../include/intfft.h"
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    intfft_generics g{12, 16, 16, 0, 0, 1, 1, 0};
    const long long batch = 65536, chunk = 2048, nch = batch / chunk;
    intfft_plan *p = nullptr;   // plan for ONE chunk
    if (intfft_plan_create(&p, &g, chunk, 0)) return 1;
    const size_t bytes = (size_t)batch * 4096 * 4, cb = (size_t)chunk * 4096 * 4;
    void *hi, *ho, *di, *dout;
    cudaMallocHost(&hi, bytes); cudaMallocHost(&ho, bytes); cudaMalloc(&di, bytes); cudaMalloc(&dout, bytes);
    cudaStream_t si, sk, so;
    cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking);
    std::vector<cudaEvent_t> ei(nch), ek(nch);
    for (auto &e : ei) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &e : ek) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaDeviceSynchronize();
            double t0 = now_ms();
            for (long long c = 0; c < nch; ++c) {
                cudaMemcpyAsync((char *)di + c * cb, (char *)hi + c * cb, cb, cudaMemcpyHostToDevice, si);
                cudaEventRecord(ei[c], si);
                cudaStreamWaitEvent(sk, ei[c], 0);
                if (mode >= 1) intfft_exec(p, (char *)di + c * cb, (char *)dout + c * cb, sk);
                cudaEventRecord(ek[c], sk);
                cudaStreamWaitEvent(so, ek[c], 0);
                cudaMemcpyAsync((char *)ho + c * cb, (char *)dout + c * cb, cb, cudaMemcpyDeviceToHost, so);
            }
            double t1 = now_ms();
            cudaDeviceSynchronize();
            double t2 = now_ms();
            std::printf("mode %d (%s): enqueue %.2f ms, total %.2f ms\n", mode, mode == 0 ? "copies only" : (mode == 1 ? "copies + exec" : "same again"), t1 - t0, t2 - t0);
        }
    }
    return 0;
}
