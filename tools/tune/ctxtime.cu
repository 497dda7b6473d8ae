// tools/tune/ctxtime.cu -- what a process pays for CUDA itself on this box before any of our code runs: cuInit + primary context
// (cudaFree(0)), then one stream.  Beside the CLI's start-up clock in tools/gpu/cli_startup.sh.
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    int n = 0;
    cudaGetDeviceCount(&n);
    const auto t1 = now();
    cudaSetDevice(0);
    cudaFree(0);
    const auto t2 = now();
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    const auto t3 = now();
    printf("{\"bare_cuda_process\": {\"devices\": %d, \"get_device_count_ms\": %.1f, \"set_device_and_context_ms\": %.1f, \"first_stream_ms\": %.1f}}\n", n, ms(t0, t1),
           ms(t1, t2), ms(t2, t3));
    return 0;
}
